set -x
R=gpurun_out
S="python tools/ncu_summary.py"
NCU="ncu --set full --clock-control none --import-source on"
T=$R/traffic.json
cp profiles/traffic.json $T
timeout 200 python -m pytest tests/test_cli_gpu.py -x -q 2>&1 | tail -3
timeout 300 $NCU -k regex:bpr_epoch -s 2 -c 1 -f -o $R/r02_bpr_large_v2 python tools/run_section.py bpr_large > $R/ncu_bpr.log 2>&1
$S $R/r02_bpr_large_v2.ncu-rep $R/r02_bpr_large_v2_ncu.csv "ncu --set full, bpr_epoch_kernel<2> with the warp-wide membership test, 4M x 1M x k=64 shape" $T bpr_large
ncu -i $R/r02_bpr_large_v2.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src.csv 2>/dev/null && python tools/ncu_lines.py /tmp/src.csv 0 25 > $R/r02_bpr_large_v2_lines.txt 2>&1
rm -f $R/*.ncu-rep
