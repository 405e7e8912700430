set -x
timeout 1200 python tools/run_configs.py > gpurun_out/run_configs.log 2>&1; tail -8 gpurun_out/run_configs.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; cat gpurun_out/r2_bench_reference_arm.json | cut -c1-300
