"""CPU checks of the drop-in boundary: every symbol that include/lqr.h and include/b200carve.h declare is
exported by the built libraries, the shim loads the engine, and without a GPU the product fails loudly
instead of falling back to a CPU path."""
import ctypes
import os
import re

import pytest

from cases import lqr, synth

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, prefix):
    src = open(os.path.join(REPO, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(%s\w+)\s*\(" % prefix, src)))


def test_lqr_header_symbols_exported_by_shim_and_oracle(pkg, oracle):
    names = [n for n in _declared("lqr.h", "lqr_")]
    assert len(names) >= 45
    shim = ctypes.CDLL(pkg.SHIM_PATH)
    orc = ctypes.CDLL(pkg.ORACLE_PATH)
    missing = [n for n in names if not hasattr(shim, n)]
    assert not missing, f"liblqr-1.so lacks {missing}"
    missing = [n for n in names if not hasattr(orc, n)]
    assert not missing, f"oracle lacks {missing}"
    # and the ctypes mirror binds exactly the declared API
    assert sorted(lqr.API) == names


def test_engine_header_symbols_exported(pkg):
    names = _declared("b200carve.h", "b200c_")
    names = [n for n in names if n != "b200c_progress_fn"]
    assert len(names) >= 25
    eng = ctypes.CDLL(pkg.ENGINE_PATH)
    missing = [n for n in names if not hasattr(eng, n)]
    assert not missing, f"libb200carve.so lacks {missing}"
    eng.b200c_abi_version.restype = ctypes.c_int
    assert eng.b200c_abi_version() == 2


def test_enum_values_match_the_pdb_contract():
    """batch-gimp-lqr.scm:51-52 / main.c:77-78 pass raw ints; render.c:772-773 needs LQR_OK == TRUE."""
    hdr = open(os.path.join(REPO, "include", "lqr.h")).read()
    for name, val in [("LQR_ERROR", 0), ("LQR_OK", 1), ("LQR_NOMEM", 2), ("LQR_USRCANCEL", 3),
                      ("LQR_RES_ORDER_HOR", 0), ("LQR_RES_ORDER_VERT", 1), ("LQR_EF_GRAD_XABS", 2),
                      ("LQR_EF_LUMA_GRAD_NORM", 3), ("LQR_EF_NULL", 6)]:
        assert re.search(r"\b%s\s*=\s*%d\b" % (name, val), hdr), name
    assert "#define __LQR_H__" in hdr  # io_functions.h:22-24


def test_product_has_no_cpu_fallback(pkg):
    """Without a CUDA device lqr_carver_new must fail (NULL), not silently compute on the host."""
    eng = ctypes.CDLL(pkg.ENGINE_PATH)
    eng.b200c_device_count.restype = ctypes.c_int
    if eng.b200c_device_count() > 0:
        pytest.skip("a GPU is present: the no-device failure mode cannot be observed")
    lib = pkg.load_product()
    with pytest.raises(lqr.LqrError):
        lib.carver(synth.flat(8, 8, 4))


def test_product_never_references_the_oracle(pkg):
    """The shipped libraries and host modules must not link, load or name anything under oracle/."""
    for path in (pkg.SHIM_PATH, pkg.ENGINE_PATH):
        blob = open(path, "rb").read()
        assert b"liblqr_oracle" not in blob and b"lqr_oracle" not in blob
    csrc = os.path.join(pkg.PKG_DIR, "csrc")
    for fn in os.listdir(csrc):
        text = open(os.path.join(csrc, fn)).read()
        assert "oracle" not in text.lower() or fn == "Makefile", fn
