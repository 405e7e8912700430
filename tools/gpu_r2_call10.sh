set -x
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-ccsd --no-cpu-baseline --no-peak --triples-per-rank 10"
$CMD > gpurun_out/plain_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_10triples.csv $CMD > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:t_gemm_tma -s 3 -c 1 -o gpurun_out/r2_t_gemm_tma $CMD > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none -k regex:t_energy_sym -s 3 -c 1 -o gpurun_out/r2_t_energy $CMD > gpurun_out/ncu_d.log 2>&1
tail -n 2 gpurun_out/ncu_c.log gpurun_out/ncu_d.log; du -sh gpurun_out
