set -x
python tools/ccsd_iter.py 2>&1 | tail -2
QCP2_CCSD_ITERS=2 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_ccsd_r1_sym.csv python tools/ccsd_iter.py > gpurun_out/ncu_ccsd_sym.log 2>&1; tail -2 gpurun_out/ncu_ccsd_sym.log
