set -x
timeout 600 python -m pytest tests/test_gpu_ws_gemm.py -x -q 2>&1 | tail -25
timeout 300 python tools/ws_gemm_bench.py > gpurun_out/r2_ws_gemm_bench.json 2> gpurun_out/r2_ws_gemm_bench.err; tail -3 gpurun_out/r2_ws_gemm_bench.err
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
