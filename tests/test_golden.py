"""Golden vectors (tests/golden, generated from the oracle by make_golden.py): the oracle must keep reproducing
them (CPU), and the CUDA path must reproduce them without the oracle in the loop (GPU)."""
import importlib
import json
import os

import numpy as np
import pytest

import cases

HERE = os.path.dirname(os.path.abspath(__file__))


def _load():
    with open(os.path.join(HERE, "golden", "golden.json")) as f:
        gold = json.load(f)
    small = np.load(os.path.join(HERE, "golden", "golden_small.npz"))
    return gold, small


def _record(res):
    import hashlib

    def sha(a):
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()

    return {"info": res.info, "image_shape": list(res.image.shape), "image_sha256": sha(res.image),
            "aux_sha256": [sha(a) for a in res.aux],
            "vmaps": [{"depth": v.depth, "orientation": v.orientation, "shape": list(v.data.shape),
                       "sha256": sha(v.data.astype(np.int32))} for v in res.vmaps],
            "n_progress": len(res.progress)}


GOLD, SMALL = _load()
BY_NAME = {c["name"]: c for c in cases.CASES}


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_reproduces_golden(oracle, name):
    res = cases.run_case(oracle, BY_NAME[name])
    assert _record(res) == GOLD[name]
    if name + "__image" in SMALL:
        assert np.array_equal(res.image, SMALL[name + "__image"])
        for i, v in enumerate(res.vmaps):
            assert np.array_equal(v.data, SMALL[f"{name}__vmap{i}"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD))
def test_cuda_reproduces_golden(product, name):
    res = cases.run_case(product, BY_NAME[name])
    assert _record(res) == GOLD[name]
    if name + "__image" in SMALL:
        assert np.array_equal(res.image, SMALL[name + "__image"])
