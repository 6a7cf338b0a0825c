set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest4.log
tail -30 gpurun_out/r02_pytest4.log
