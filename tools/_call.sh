set -x
R=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 4 --steps 5 --warmup 3 > $R/r02_bench_n4.json 2> $R/r02_bench_n4.err
grep '^{"metric"' $R/r02_bench_n4.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['loss'], json.dumps(d['e2e'])[:900])"; tail -3 $R/r02_bench_n4.err
