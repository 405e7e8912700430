"""Per-kernel table of one CCSD iteration from an `ncu --metrics gpu__time_duration.sum --csv` launch list
(the last full iteration = the launches between the last two ccsd_energy_kernel launches).
usage: python tools/launch_table.py profiles/r2_launches_ccsd.csv [--seq]"""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            h, start = r, i + 1
            break
    ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
    seq = []
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("qcp2::", "").replace("<unnamed>::", "")
        seq.append((name[:86], float(r[vi].replace(",", "")) / 1000.0, r[gi]))
    return seq


def main():
    seq = load(sys.argv[1])
    idx = [i for i, (n, _, _) in enumerate(seq) if n.startswith("ccsd_energy")]
    it = seq[idx[-2]:idx[-1]] if len(idx) >= 2 else seq
    print(f"one iteration: {len(it)} launches, {sum(t for _, t, _ in it):.1f} us serialised")
    if "--seq" in sys.argv:
        for n, t, g in it:
            print(f"{t:9.1f}  {g:16s} {n}")
        return
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t, _ in it:
        agg[n][0] += 1
        agg[n][1] += t
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t:9.1f} us {c:3d} x  {n}")


if __name__ == "__main__":
    main()
