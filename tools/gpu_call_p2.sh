N=${1:-2}
cat > /tmp/in.json <<EOT
{"filename": "geom/methane.xyz", "type": "CCSD(T)", "ref": "RHF", "basis": "cc-pvtz", "charge": 0, "multiplicity": 1, "maxiter": 400, "scf_conv": 1e-10, "cc_conv": 1e-12}
EOT
timeout 200 ./qc-p2_b200/bin/P2 --fused --gpus $N /tmp/in.json 2>&1 | tail -3
timeout 200 ./qc-p2_b200/bin/P2 --fused /tmp/in.json 2>&1 | tail -1
