set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
QCP2_LARGE_N=150 timeout 900 python tools/post_hf_breakdown.py > gpurun_out/r2_post_hf_breakdown.txt 2>&1; tail -30 gpurun_out/r2_post_hf_breakdown.txt
QCP2_CCSD_ITERS=2 QCP2_CCSD_NO_GRAPH=1 timeout 300 python tools/ccsd_iter.py > gpurun_out/plain2.log 2>&1 && QCP2_CCSD_ITERS=2 QCP2_CCSD_NO_GRAPH=1 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:\(int\)6, \(int\)4>' -s 1 -c 1 -o gpurun_out/r2_ccsd_ladder python tools/ccsd_iter.py > gpurun_out/ncu2.log 2>&1
QCP2_CCSD_ITERS=2 QCP2_CCSD_NO_GRAPH=1 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:\(int\)5, \(int\)8>' -s 2 -c 2 -o gpurun_out/r2_ccsd_ring python tools/ccsd_iter.py > gpurun_out/ncu3.log 2>&1
tail -n 2 gpurun_out/ncu2.log gpurun_out/ncu3.log
du -sh gpurun_out
