set -x
export QMFB_LIB=qmf_b200/libqmf_b200_prof.so
QMFB_SOLVE=classic timeout 300 python tools/exp_phases.py 11840,208,17800 > gpurun_out/r02_phases_classic_mma.log 2>&1
cat gpurun_out/r02_phases_classic_mma.log
