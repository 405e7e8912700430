set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
QCP2_CCSD_ITERS=8 timeout 300 python tools/ccsd_iter.py 2>&1 | tail -2
QCP2_GEMM_TRACE=1 QCP2_CCSD_ITERS=1 QCP2_CCSD_NO_GRAPH=1 timeout 300 python tools/ccsd_iter.py 2> gpurun_out/r2_ccsd_gemm_trace.txt | tail -1
QCP2_CCSD_ITERS=3 QCP2_CCSD_NO_GRAPH=1 timeout 300 python tools/ccsd_iter.py > gpurun_out/plain.log 2>&1 && QCP2_CCSD_ITERS=3 QCP2_CCSD_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_ccsd_v3.csv python tools/ccsd_iter.py > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
