set -x
python bench.py > gpurun_out/bench_r1_final_n1.json 2> gpurun_out/bench_r1_final_n1.err; tail -c 600 gpurun_out/bench_r1_final_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_final_reference.json 2> gpurun_out/bench_r1_final_reference.err; tail -c 900 gpurun_out/bench_r1_final_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_final2.csv python bench.py --steps 2 --warmup 1 --no-ccsd --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench_final.log 2>&1; tail -2 gpurun_out/ncu_bench_final.log
QCP2_CCSD_ITERS=3 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_ccsd_r1_final2.csv python tools/ccsd_iter.py > gpurun_out/ncu_ccsd_final.log 2>&1; tail -1 gpurun_out/ncu_ccsd_final.log
QCP2_CCSD_ITERS=1 ncu --set full --clock-control none -k regex:dgemm_kernel --launch-skip 40 -c 45 -o gpurun_out/prof_ccsd_gemms_r1 -f python tools/ccsd_iter.py > gpurun_out/ncu_ccsd_gemms.log 2>&1; tail -2 gpurun_out/ncu_ccsd_gemms.log
python tools/bandwidth_kernels.py > gpurun_out/bw_r1d.log 2>&1; grep -E "transform_ao_mo|to_so|slice_ints vvvv" gpurun_out/bw_r1d.log
cp gpurun_out/bandwidth_kernels.json gpurun_out/bandwidth_kernels_r1d.json
