# Round-2 evidence capture (run under gpurun).  Plain (un-profiled) runs first, then the ncu passes of the same
# commands.  The .ncu-rep files are summarised ON THE BOX (tools/ncu_summary.py, tools/ncu_lines.py) into small CSV /
# text files and deleted: gpurun brings back at most 64 MiB.  Outputs: gpurun_out/r02_*; tools/summarise_profiles.sh
# copies them into profiles/.
set -x
R=gpurun_out
Q="--no-e2e --no-bpr --no-cpu-baseline"
S="python tools/ncu_summary.py"
NCU="ncu --set full --clock-control none --import-source on"
T=$R/traffic.json
cp profiles/traffic.json $T 2>/dev/null
summ() {  # summ <name> <comment> [traffic key]: summary csv (+ traffic) + hot source lines, then drop the report
  $S $R/$1.ncu-rep $R/$1_ncu.csv "$2" $T $3
  ncu -i $R/$1.ncu-rep --page source --csv --print-source cuda,sass > /tmp/$1_src.csv 2>/dev/null && python tools/ncu_lines.py /tmp/$1_src.csv 0 40 > $R/$1_lines.txt 2>&1
  rm -f $R/$1.ncu-rep
}
python bench.py --workload c1 --steps 20 --warmup 5 $Q > $R/r02_bench_c1.json 2> $R/r02_bench_c1.err
python bench.py --workload c3 --steps 5 --warmup 3 $Q > $R/r02_bench_c3.json 2> $R/r02_bench_c3.err
python bench.py --steps 2 --warmup 1 $Q > $R/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $R/r02_launches.csv python bench.py --steps 2 --warmup 1 $Q > $R/ncu_l.log 2>&1
# ncu --set full: user-rows and item-rows launch of C4 separately (one capture of both gave NaN counters for the second)
$NCU -k regex:wals_solve -s 2 -c 1 -f -o $R/r02_solve_c4_user python bench.py --steps 1 --warmup 1 $Q > $R/ncu_u.log 2>&1
summ r02_solve_c4_user "ncu --set full, wals_solve_kernel<16>, C4 user-rows launch (bench.py --steps 1 --warmup 1)" c4_user
$NCU -k regex:wals_solve -s 3 -c 1 -f -o $R/r02_solve_c4_item python bench.py --steps 1 --warmup 1 $Q > $R/ncu_i.log 2>&1
summ r02_solve_c4_item "ncu --set full, wals_solve_kernel<16>, C4 item-rows launch" c4_item
$NCU -k regex:wals_solve -s 2 -c 2 -f -o $R/r02_solve_c1 python bench.py --workload c1 --steps 1 --warmup 1 $Q > $R/ncu_c1.log 2>&1
summ r02_solve_c1 "ncu --set full, wals_solve_kernel<4>, C1 user-rows and item-rows launches" c1_both
$NCU -k regex:wals_solve -s 2 -c 1 -f -o $R/r02_solve_c3_user python bench.py --workload c3 --steps 1 --warmup 1 $Q > $R/ncu_c3u.log 2>&1
summ r02_solve_c3_user "ncu --set full, wals_solve_kernel<8>, C3 user-rows launch" c3_user
$NCU -k regex:wals_solve -s 3 -c 1 -f -o $R/r02_solve_c3_item python bench.py --workload c3 --steps 1 --warmup 1 $Q > $R/ncu_c3i.log 2>&1
summ r02_solve_c3_item "ncu --set full, wals_solve_kernel<8>, C3 item-rows launch" c3_item
python tools/run_section.py bpr_large > $R/r02_bpr_large.json 2> $R/plain.log &&
$NCU -k regex:bpr_epoch -s 2 -c 1 -f -o $R/r02_bpr_large python tools/run_section.py bpr_large > $R/ncu_bpr.log 2>&1
summ r02_bpr_large "ncu --set full, bpr_epoch_kernel, 4M x 1M x k=64 shape" bpr_large
$NCU -k regex:bpr_epoch -s 2 -c 1 -f -o $R/r02_bpr_c2 python tools/run_section.py bpr_c2 > $R/ncu_bpr2.log 2>&1
summ r02_bpr_c2 "ncu --set full, bpr_epoch_kernel, C2 (10k x 5k, k=30, biases)" bpr_c2
python tools/run_section.py eval_large > $R/r02_eval_large.json 2> $R/plain.log &&
$NCU -k regex:eval_score -s 1 -c 1 -f -o $R/r02_eval_large python tools/run_section.py eval_large > $R/ncu_ev.log 2>&1
summ r02_eval_large "ncu --set full, eval_score_kernel, 100k x 1M x k=128" eval_large
$NCU -k regex:eval_score -s 1 -c 1 -f -o $R/r02_eval_c2 python tools/run_section.py eval_c2 > $R/ncu_ev2.log 2>&1
summ r02_eval_c2 "ncu --set full, eval_score_kernel, C2 all users (10k x 5k x k=30)" eval_c2
timeout 300 python -m pytest tests/test_bpr_c2_gpu.py -q -s 2>&1 | grep "GPU BPR" > $R/r02_bpr_c2_means.txt
rm -f $R/*.ncu-rep; du -sh $R; ls -la $R
