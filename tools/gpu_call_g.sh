set -x
QCP2_CCSD_ITERS=3 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_ccsd_r1_v3.csv python tools/ccsd_iter.py > gpurun_out/ncu_ccsd_v3.log 2>&1; tail -1 gpurun_out/ncu_ccsd_v3.log
