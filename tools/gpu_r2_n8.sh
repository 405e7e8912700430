set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8_dealt.json 2> gpurun_out/r2_bench_n8_dealt.err; tail -2 gpurun_out/r2_bench_n8_dealt.err
