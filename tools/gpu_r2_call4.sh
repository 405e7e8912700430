set -x
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -3 gpurun_out/r2_bench_n1.err
QCP2_CCSD_ITERS=3 QCP2_CCSD_NO_GRAPH=1 timeout 300 python tools/ccsd_iter.py > gpurun_out/plain.log 2>&1 && QCP2_CCSD_ITERS=3 QCP2_CCSD_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_ccsd.csv python tools/ccsd_iter.py > gpurun_out/ncu.log 2>&1
QCP2_CCSD_ITERS=2 QCP2_CCSD_NO_GRAPH=1 timeout 300 python tools/ccsd_iter.py > gpurun_out/plain2.log 2>&1 && QCP2_CCSD_ITERS=2 QCP2_CCSD_NO_GRAPH=1 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:WsShape<1, 8, 6, 4>' -s 1 -c 1 -o gpurun_out/r2_ccsd_ladder python tools/ccsd_iter.py > gpurun_out/ncu2.log 2>&1
QCP2_CCSD_ITERS=2 QCP2_CCSD_NO_GRAPH=1 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:WsShape<2, 4, 5, 8>' -s 2 -c 2 -o gpurun_out/r2_ccsd_ring python tools/ccsd_iter.py > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu2.log gpurun_out/ncu3.log
du -sh gpurun_out
