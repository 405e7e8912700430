set -x
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_stream_n$N.json 2> gpurun_out/bench_stream_n$N.err; tail -c 1800 gpurun_out/bench_stream_n$N.json; tail -5 gpurun_out/bench_stream_n$N.err
cat > /tmp/in.json <<EOT
{"filename": "geom/water.xyz", "type": "CCSD(T)", "ref": "RHF", "basis": "cc-pvdz", "charge": 0, "multiplicity": 1, "maxiter": 200, "scf_conv": 1e-12, "cc_conv": 1e-12}
EOT
./qc-p2_b200/bin/P2 --gpus $N /tmp/in.json 2>&1 | tail -4
./qc-p2_b200/bin/P2 /tmp/in.json 2>&1 | tail -2
