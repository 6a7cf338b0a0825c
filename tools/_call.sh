set -x
R=gpurun_out
nvidia-smi -L | head -8
timeout 540 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --workload c5 --steps 3 --warmup 1 > $R/r02_bench_c5_n8.json 2> $R/r02_bench_c5_n8.err
cat $R/r02_bench_c5_n8.json; tail -15 $R/r02_bench_c5_n8.err
