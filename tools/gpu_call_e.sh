set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
QCP2_CCSD_ITERS=8 python tools/ccsd_iter.py 2>&1 | tail -1
