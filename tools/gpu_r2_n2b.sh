set -x
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 --no-ccsd > gpurun_out/r2_bench_n2_dealt.json 2> gpurun_out/r2_bench_n2_dealt.err; tail -2 gpurun_out/r2_bench_n2_dealt.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n2_dealt.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'])"
