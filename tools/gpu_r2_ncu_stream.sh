set -x
QCP2_CCSD_ITERS=2 QCP2_CCSD_NO_GRAPH=1 python tools/ccsd_iter.py > gpurun_out/plain.log 2>&1 && QCP2_CCSD_ITERS=2 QCP2_CCSD_NO_GRAPH=1 ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:\(int\)2, \(int\)4>' -s 6 -c 4 -o gpurun_out/r2_ccsd_stream python tools/ccsd_iter.py > gpurun_out/ncu_s.log 2>&1
QCP2_CCSD_ITERS=2 QCP2_CCSD_NO_GRAPH=1 ncu --set full --clock-control none -k 'regex:finalize_t2|tau_pack|ccsd_energy' -s 3 -c 4 -o gpurun_out/r2_ccsd_elementwise python tools/ccsd_iter.py > gpurun_out/ncu_e.log 2>&1
du -sh gpurun_out
