"""Locate and import the UNMODIFIED reference package (TEST / BASELINE INFRASTRUCTURE ONLY).

Search order: `baseline/_ref/` (the pip-installed copy: `python -m pip install --no-index --no-build-isolation --no-deps
--target baseline/_ref <reference>`, git-ignored, travels to the GPU box with the gpurun snapshot), then /root/reference
(exists only in the build container).  Used to (a) generate the golden fixtures under tests/golden/ (oracle/gen_golden*.py),
(b) cross-check the numpy oracle and (c) time the reference's own CPU forward in `bench.py --impl reference` /
`cpu_baseline`.  `einx` is not installable here and is never called on the hot path: oracle/einx_shim stands in for it.
"""
import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "einx_shim")
CANDIDATES = [os.environ.get("VQB_REFERENCE_ROOT"), os.path.join(os.path.dirname(_HERE), "baseline", "_ref"), "/root/reference"]


def reference_root():
    for c in CANDIDATES:
        if c and os.path.isdir(os.path.join(c, "vector_quantize_pytorch")):
            return c
    return None


REFERENCE_ROOT = reference_root()


def reference_available() -> bool:
    return reference_root() is not None


def load_reference():
    """Returns the imported `vector_quantize_pytorch` reference module (einx shimmed)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not found (baseline/_ref, /root/reference)")
    try:
        importlib.import_module("einx")
    except ImportError:
        sys.path.insert(0, _SHIM)
    if root not in sys.path:
        sys.path.insert(0, root)
    return importlib.import_module("vector_quantize_pytorch")
