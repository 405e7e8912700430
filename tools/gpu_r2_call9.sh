set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 900 python bench.py > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err; tail -3 gpurun_out/r2_bench_n1_b.err
