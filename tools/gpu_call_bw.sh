set -x
python -m pytest tests/test_gpu_molecules.py -x -q 2>&1 | tail -5
python tools/bandwidth_kernels.py > gpurun_out/bw_r1b.log 2>&1; tail -3 gpurun_out/bw_r1b.log
cp gpurun_out/bandwidth_kernels.json gpurun_out/bandwidth_kernels_r1b.json
QCP2_BW_ONCE=1 ncu --set full --clock-control none -k regex:'to_so_kernel|slice_kernel|permute_tiled_kernel|permute_generic_kernel|tau_kernel|mp2_kernel|ccsd_energy_kernel|multiply_ts_kernel|denom_from_F_kernel' -c 24 -o gpurun_out/prof_bw_r1 -f python tools/bandwidth_kernels.py > gpurun_out/ncu_bw.log 2>&1; tail -3 gpurun_out/ncu_bw.log
