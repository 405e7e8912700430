set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/bandwidth_kernels.py > gpurun_out/bw_r1c.log 2>&1; grep -E "to_so|slice" gpurun_out/bw_r1c.log
cp gpurun_out/bandwidth_kernels.json gpurun_out/bandwidth_kernels_r1c.json
