set -x
python -m pytest tests/test_gpu_parity.py -x -q -k mgpu 2>&1 | tail -40
python tools/e2e_breakdown.py 2>&1 | tail -30
