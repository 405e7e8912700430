set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
QCP2_CH4_ITERS=8 timeout 300 python tools/ch4_iter.py 2>&1 | tail -2
QCP2_CCSD_NO_GRAPH=1 QCP2_CH4_ITERS=3 timeout 300 python tools/ch4_iter.py > gpurun_out/plain.log 2>&1 && QCP2_CCSD_NO_GRAPH=1 QCP2_CH4_ITERS=3 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_ch4.csv python tools/ch4_iter.py > gpurun_out/ncu.log 2>&1
QCP2_GEMM_TRACE=1 QCP2_CCSD_NO_GRAPH=1 QCP2_CH4_ITERS=1 timeout 300 python tools/ch4_iter.py 2> gpurun_out/r2_ch4_gemm_trace.txt | tail -1
