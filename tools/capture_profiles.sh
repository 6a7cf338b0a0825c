# Round-2 evidence capture (run under gpurun): plain bench lines first, then the ncu passes of the same commands.
# Outputs go to gpurun_out/r02_*; tools/ncu_summary.py turns the .ncu-rep files into profiles/*.csv here.
set -x
R=gpurun_out
Q="--no-e2e --no-bpr --no-cpu-baseline"
python bench.py --steps 5 --warmup 3 > $R/r02_bench_n1.json 2> $R/r02_bench_n1.err || exit 1
python bench.py --workload c1 --steps 20 --warmup 5 $Q > $R/r02_bench_c1.json 2> $R/r02_bench_c1.err || exit 1
python bench.py --workload c3 --steps 5 --warmup 3 $Q > $R/r02_bench_c3.json 2> $R/r02_bench_c3.err || exit 1
python bench.py --steps 2 --warmup 1 $Q > $R/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $R/r02_launches.csv python bench.py --steps 2 --warmup 1 $Q > $R/ncu_l.log 2>&1
# ncu --set full: user-rows and item-rows launch of C4 separately (one capture of both gave NaN counters for the second)
NCU="ncu --set full --clock-control none --import-source on"
python bench.py --steps 1 --warmup 1 $Q > $R/plain.log 2>&1 &&
$NCU -k regex:wals_solve -s 2 -c 1 -f -o $R/r02_solve_c4_user python bench.py --steps 1 --warmup 1 $Q > $R/ncu_u.log 2>&1
$NCU -k regex:wals_solve -s 3 -c 1 -f -o $R/r02_solve_c4_item python bench.py --steps 1 --warmup 1 $Q > $R/ncu_i.log 2>&1
python bench.py --workload c1 --steps 1 --warmup 1 $Q > $R/plain.log 2>&1 &&
$NCU -k regex:wals_solve -s 2 -c 2 -f -o $R/r02_solve_c1 python bench.py --workload c1 --steps 1 --warmup 1 $Q > $R/ncu_c1.log 2>&1
python bench.py --workload c3 --steps 1 --warmup 1 $Q > $R/plain.log 2>&1 &&
$NCU -k regex:wals_solve -s 2 -c 1 -f -o $R/r02_solve_c3_user python bench.py --workload c3 --steps 1 --warmup 1 $Q > $R/ncu_c3u.log 2>&1
$NCU -k regex:wals_solve -s 3 -c 1 -f -o $R/r02_solve_c3_item python bench.py --workload c3 --steps 1 --warmup 1 $Q > $R/ncu_c3i.log 2>&1
python tools/run_section.py bpr_large > $R/plain.log 2>&1 &&
$NCU -k regex:bpr_epoch -s 2 -c 1 -f -o $R/r02_bpr_large python tools/run_section.py bpr_large > $R/ncu_bpr.log 2>&1
python tools/run_section.py bpr_c2 > $R/plain.log 2>&1 &&
$NCU -k regex:bpr_epoch -s 2 -c 1 -f -o $R/r02_bpr_c2 python tools/run_section.py bpr_c2 > $R/ncu_bpr2.log 2>&1
python tools/run_section.py eval_large > $R/plain.log 2>&1 &&
$NCU -k regex:eval_score -s 1 -c 1 -f -o $R/r02_eval_large python tools/run_section.py eval_large > $R/ncu_ev.log 2>&1
python tools/run_section.py eval_c2 > $R/plain.log 2>&1 &&
$NCU -k regex:eval_score -s 1 -c 1 -f -o $R/r02_eval_c2 python tools/run_section.py eval_c2 > $R/ncu_ev2.log 2>&1
ls -la $R/r02_*.ncu-rep
