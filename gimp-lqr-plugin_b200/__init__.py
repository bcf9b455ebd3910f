"""B200-native seam-carving engine behind the LqrCarver C API (drop-in for gimp-lqr-plugin's render path).

The product is two shared libraries built from ``csrc/``:
  * ``libb200carve.so`` -- the CUDA engine (sm_100a kernels + C ABI, include/b200carve.h)
  * ``liblqr-1.so``     -- the plain-C shim exporting include/lqr.h and dlopen()ing the engine
This Python package is only the test / bench harness around them (ctypes, no torch types at the boundary).
"""
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
SHIM_PATH = os.path.join(PKG_DIR, "liblqr-1.so")
ENGINE_PATH = os.path.join(PKG_DIR, "libb200carve.so")
ORACLE_PATH = os.path.join(REPO_DIR, "oracle", "liblqr_oracle.so")

from . import lqr, synth, render  # noqa: E402,F401


def load_product() -> "lqr.LqrLib":
    """The shipped path.  Fails loudly when the CUDA engine is missing -- there is no CPU fallback."""
    for p in (SHIM_PATH, ENGINE_PATH):
        if not os.path.exists(p):
            raise RuntimeError(f"{p} is not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
    return lqr.LqrLib(SHIM_PATH)


def load_oracle() -> "lqr.LqrLib":
    """CPU oracle: tests, smoke() and bench.py's cpu_baseline only."""
    return lqr.LqrLib(ORACLE_PATH)
