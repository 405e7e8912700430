"""CPU tests of the boundary: the C-ABI library loads, exports every symbol the header
declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import qcp2_b200


def declared_symbols():
    text = open(qcp2_b200.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qcp2_[a-zA-Z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = qcp2_b200.load_library()
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/qcp2_b200.h but not exported"


def test_version_string():
    assert b"sm_100a" in qcp2_b200.load_library().qcp2_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is exercised on the CPU box")
    h = ctypes.c_void_p()
    rc = qcp2_b200.load_library().qcp2_create(0, ctypes.byref(h))
    assert rc != 0 and not h
    with pytest.raises(qcp2_b200.QCP2Error):
        qcp2_b200.Context(0)


def test_product_path_does_not_reference_oracle():
    root = os.path.dirname(qcp2_b200.HEADER_PATH)
    pkg = os.path.join(os.path.dirname(root), "qc-p2_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".py", "Makefile")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"


def test_p2_driver_refuses_to_run_without_gpus(tmp_path):
    """The C++ host mirror has no CPU path either: `P2` fails cleanly without a device, and `P2 --gpus N` checks
    the visible GPU count (in a throw-away child, so the parent can still fork its workers) before spawning anything."""
    import json
    import subprocess

    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is exercised on the CPU box")
    root = os.path.dirname(os.path.dirname(qcp2_b200.HEADER_PATH))
    exe = os.path.join(root, "qc-p2_b200", "bin", "P2")
    assert os.path.exists(exe), "build qc-p2_b200/host first (__graft_entry__.build())"
    cfg = json.load(open(os.path.join(root, "input.json")))
    inp = tmp_path / "input.json"
    inp.write_text(json.dumps(cfg))
    multi = subprocess.run([exe, "--gpus", "2", str(inp)], cwd=root, capture_output=True, text=True, timeout=120)
    assert multi.returncode != 0 and "GPU(s) visible" in multi.stderr
    single = subprocess.run([exe, str(inp)], cwd=root, capture_output=True, text=True, timeout=300)
    assert single.returncode != 0 and "no CPU fallback" in single.stderr
    usage = subprocess.run([exe], cwd=root, capture_output=True, text=True, timeout=30)
    assert usage.returncode == 2 and "usage" in usage.stderr
