set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
QCP2_CCSD_ITERS=8 python tools/ccsd_iter.py 2>&1 | tail -1
QCP2_GEMM_TRACE=1 QCP2_CCSD_ITERS=1 python tools/ccsd_iter.py 2>&1 | grep "qcp2 gemm" | tail -45 > gpurun_out/gemm_trace.txt
