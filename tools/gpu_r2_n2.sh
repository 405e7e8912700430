set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 --no-ccsd > gpurun_out/r2_bench_n2_dealt.json 2> gpurun_out/r2_bench_n2_dealt.err; tail -2 gpurun_out/r2_bench_n2_dealt.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n2_dealt.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['seconds'], d['e2e']['e_t_matches_n1'], d['e2e']['strong_scaling_efficiency'])"
python bench.py --steps 2 --warmup 3 --no-ccsd --no-cpu-baseline > gpurun_out/r2_bench_n1_c.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n1_c.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['seconds'], d['e2e']['e_t_matches_n1'])"
