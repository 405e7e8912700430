set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
QCP2_CH4_ITERS=8 timeout 300 python tools/ch4_iter.py 2>&1 | tail -1
QCP2_CCSD_ITERS=8 timeout 300 python tools/ccsd_iter.py 2>&1 | tail -1
timeout 600 python tools/post_hf_breakdown.py 2>&1 | grep "rep \|CCSD loop"
