set -x
R=gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 8 --steps 5 --warmup 3 > $R/r02_bench_n8.json 2> $R/r02_bench_n8.err
grep '^{"metric"' $R/r02_bench_n8.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['loss'], json.dumps(d['e2e']))"; tail -5 $R/r02_bench_n8.err
