"""bench.py contract checks that need no GPU: the reference arm prints exactly ONE JSON line on stdout
with the keys the driver reads, and the host-side helpers (flop counts, configs) are consistent."""
import json
import math
import os
import subprocess
import sys

from helpers import ROOT


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--o", "6", "--v", "24"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ccsd_t_fp64_tflops" and d["unit"] == "TFLOP/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    own = d.get("reference_own_code")  # the reference's own CCSD_T (oracle/_ref), present where it was built
    if own and "error" not in own:
        assert own["kind"] == "reference" and own["abs_diff_vs_port"] < 1e-9 and own["value"] > 0


def test_algorithmic_flop_count_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    # C(o,3) * 2 v [v(v-1)/2] 3(v+o) = 8.326e14 at o=40, v=400
    per = bench.t_flops_per_triple(40, 400)
    assert per == 2.0 * 400 * (400 * 399 / 2) * 3 * 440
    assert abs(math.comb(40, 3) * per - 8.326e14) < 1e11
