set -x
for v in pA pB; do
  echo "=== $v classic GRID_CAP=148" >> gpurun_out/r02_phases_1cta.log
  QMFB_GRID_CAP=148 QMFB_LIB=qmf_b200/variants/libqmf_b200_$v.so QMFB_SOLVE=classic timeout 100 python tools/exp_phases.py 23680,208,17800 >> gpurun_out/r02_phases_1cta.log 2>&1 || echo "variant $v: FAILED/TIMEOUT rc=$?" >> gpurun_out/r02_phases_1cta.log
done
