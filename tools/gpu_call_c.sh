set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/e2e_breakdown.py 2>&1 | tail -12
python bench.py --steps 2 --warmup 3 --no-ccsd --no-cpu-baseline > gpurun_out/bench_stream_n1.json 2> gpurun_out/bench_stream_n1.err; tail -c 1500 gpurun_out/bench_stream_n1.json; tail -5 gpurun_out/bench_stream_n1.err
