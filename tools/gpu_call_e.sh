set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
QCP2_CCSD_ITERS=8 python tools/ccsd_iter.py 2>&1 | tail -1
QCP2_CCSD_NO_GRAPH=1 QCP2_CCSD_ITERS=8 python tools/ccsd_iter.py 2>&1 | tail -1
python tools/run_configs.py --help 2>&1 | head -5
