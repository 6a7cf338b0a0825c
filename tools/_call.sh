set -x
python -m pytest tests/test_eval_gpu.py tests/test_sharded_gpu.py tests/test_bpr_eval_gpu.py tests/test_golden_gpu.py tests/test_cli_gpu.py -x -q > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest2.log
tail -30 gpurun_out/r02_pytest2.log
python - > gpurun_out/r02_eval_bench.log 2>&1 <<'PY'
import json, sys
sys.path.insert(0, '.')
import bench
for shape in ("c2", "large"):
    print(json.dumps(bench.run_eval_ours(shape)), flush=True)
PY
cat gpurun_out/r02_eval_bench.log
