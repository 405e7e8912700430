set -x
N=${1:-8}
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 2>&1 | tail -4
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_r1b_n$N.json 2> gpurun_out/bench_r1b_n$N.err; tail -c 2500 gpurun_out/bench_r1b_n$N.json; tail -3 gpurun_out/bench_r1b_n$N.err
cat > /tmp/in.json <<EOT
{"filename": "geom/water.xyz", "type": "CCSD(T)", "ref": "RHF", "basis": "cc-pvdz", "charge": 0, "multiplicity": 1, "maxiter": 200, "scf_conv": 1e-12, "cc_conv": 1e-12}
EOT
timeout 120 ./qc-p2_b200/bin/P2 --gpus $N /tmp/in.json 2>&1 | tail -2
