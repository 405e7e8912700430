set -x
QCP2_CCSD_ITERS=3 QCP2_CCSD_NO_GRAPH=1 python tools/ccsd_iter.py > gpurun_out/plain.log 2>&1 && QCP2_CCSD_ITERS=3 QCP2_CCSD_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_ccsd.csv python tools/ccsd_iter.py > gpurun_out/ncu.log 2>&1
QCP2_CCSD_NO_GRAPH=1 QCP2_CH4_ITERS=3 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_ch4.csv python tools/ch4_iter.py > gpurun_out/ncu2.log 2>&1
QCP2_CH4_ITERS=29 python tools/ch4_iter.py 2>&1 | tail -1
