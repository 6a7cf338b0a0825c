set -x
R=gpurun_out
Q="--no-e2e --no-bpr --no-cpu-baseline"
(time timeout 900 python -m pytest tests -m gpu -x -q) > $R/r02_gputests_final.log 2>&1; tail -6 $R/r02_gputests_final.log
timeout 200 python bench.py --workload c1 --steps 20 --warmup 5 $Q > $R/r02_bench_c1.json 2> $R/r02_bench_c1.err
timeout 200 python bench.py --workload c3 --steps 5 --warmup 3 $Q > $R/r02_bench_c3.json 2> $R/r02_bench_c3.err
timeout 600 python bench.py --steps 5 --warmup 3 > $R/r02_bench_n1.json 2> $R/r02_bench_n1.err; tail -3 $R/r02_bench_n1.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $R/r02_bench_ref_n1.json 2> $R/r02_bench_ref_n1.err; cat $R/r02_bench_ref_n1.json | cut -c1-600
python -c "import __graft_entry__ as g; g.smoke()"
