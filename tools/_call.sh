set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log
tail -5 gpurun_out/r02_pytest1.log
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-bpr > gpurun_out/r02_plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wals_solve_kernel -s 2 -c 2 -o gpurun_out/r02_solve_base python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-bpr > gpurun_out/r02_ncu1.log 2>&1
tail -3 gpurun_out/r02_plain1.log
