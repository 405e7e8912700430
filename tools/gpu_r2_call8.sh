set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
QCP2_CH4_ITERS=8 timeout 300 python tools/ch4_iter.py 2>&1 | tail -1
QCP2_CCSD_ITERS=8 timeout 300 python tools/ccsd_iter.py 2>&1 | tail -1
QCP2_CCSD_NO_GRAPH=1 QCP2_CH4_ITERS=3 timeout 300 python tools/ch4_iter.py > gpurun_out/plain.log 2>&1 && QCP2_CCSD_NO_GRAPH=1 QCP2_CH4_ITERS=3 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_ch4_v3.csv python tools/ch4_iter.py > gpurun_out/ncu.log 2>&1
QCP2_CCSD_ITERS=3 QCP2_CCSD_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_ccsd_v4.csv python tools/ccsd_iter.py > gpurun_out/ncu2.log 2>&1
