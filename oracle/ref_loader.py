"""Locate and import the UNMODIFIED reference package (TEST INFRASTRUCTURE ONLY).

Only used in the build container, where /root/reference exists, to (a) generate the golden
fixtures under tests/golden/ (oracle/gen_golden.py) and (b) cross-check the numpy oracle.  The GPU
box has no /root/reference: nothing under tests marked `gpu`, smoke() or bench.py calls this.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("VQB_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "einx_shim")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "vector_quantize_pytorch"))


def load_reference():
    """Returns the imported `vector_quantize_pytorch` reference module (einx shimmed)."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    try:
        importlib.import_module("einx")
    except ImportError:
        sys.path.insert(0, _SHIM)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module("vector_quantize_pytorch")
