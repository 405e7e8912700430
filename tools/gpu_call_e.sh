set -x
python -m pytest tests -m gpu -x -q -k "pipeline or ccsd or molecules or config or intermediates" 2>&1 | tail -6
QCP2_CCSD_ITERS=8 python tools/ccsd_iter.py 2>&1 | tail -1
