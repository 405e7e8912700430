"""Import shim: the package directory is named `qc-p2_b200` (not a valid Python identifier),
so `import qcp2_b200` loads qc-p2_b200/qcp2.py and re-exports it."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(_here, "qc-p2_b200", rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_impl = _load("qcp2_b200_impl", "qcp2.py")
shard = _load("qcp2_b200_shard", "shard.py")
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
