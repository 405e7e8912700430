set -x
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 2>&1 | tail -3
