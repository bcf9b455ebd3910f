"""the plug-in's own host loops next to the hot path -- write_vmap_to_layer's colouring
(reference src/io_functions.c:249-279) and guess_new_size (src/layers_combo.c:274-392).

CPU: the C restatement (oracle/plugin_oracle.c) against an independent numpy transcription of the reference
formulas, hand-computed values and the committed golden fixture.  GPU: the CUDA engine (include/b200carve.h)
against the C restatement, byte for byte."""
import hashlib
import importlib
import json
import os

import numpy as np
import pytest

import plugin_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "plugin_ops.json")

# colour pairs: the plug-in's defaults (yellow -> red: main.c col_vals) and arbitrary non-dyadic fractions
COLOURS = [((1.0, 1.0, 0.0), (1.0, 0.0, 0.0)), ((0.2, 0.7, 1 / 3), (0.9, 0.05, 0.6)), ((0.0, 0.0, 0.0), (1.0, 1.0, 1.0))]


def make_vmap(w, h, depth, seed):
    """A seam map as the engine dumps it: every row holds each of 1..depth once, 0 elsewhere."""
    rng = np.random.default_rng(seed)
    vm = np.zeros((h, w), dtype=np.int32)
    for y in range(h):
        cols = rng.choice(w, size=depth, replace=False)
        vm[y, cols] = rng.permutation(depth) + 1
    return vm


def numpy_vmap_colour(vm, depth, cs, ce):
    value = (depth + 1 - vm).astype(np.float64) / (depth + 1)
    out = np.zeros(vm.shape + (4,), dtype=np.uint8)
    for k in range(3):
        out[..., k] = (255 * (value * cs[k] + (1 - value) * ce[k])).astype(np.uint8)  # truncation, like the C cast
    out[..., 3] = (255 * (0.5 * (1 + value))).astype(np.uint8)
    out[vm == 0] = 0
    return out


def make_mask(w, h, bpp, seed):
    rng = np.random.default_rng(seed)
    m = rng.integers(0, 256, size=(h, w, bpp), dtype=np.uint8)
    m[rng.random((h, w)) < 0.5] = 0  # half the pixels empty
    return m


def numpy_guess(mask, has_alpha, x_off, y_off, ow, oh, direction):
    h, w, bpp = mask.shape
    c_bpp = bpp - (1 if has_alpha else 0)
    inten = mask[..., :c_bpp].astype(np.float64).sum(axis=2) / (255 * c_bpp)
    if has_alpha:
        inten = inten * (mask[..., bpp - 1].astype(np.float64) / 255)
    hit = inten >= 0.5 / c_bpp
    # the part of the mask over the layer
    x0, x1 = max(0, x_off), min(ow, w + x_off)
    y0, y1 = max(0, y_off), min(oh, h + y_off)
    old = ow if direction == 0 else oh
    if x1 <= x0 or y1 <= y0:
        return old
    sub = hit[y0 - y_off:y1 - y_off, x0 - x_off:x1 - x_off]
    return old - int(sub.sum(axis=1 if direction == 0 else 0).max())


GUESS_CASES = [  # (mask w, h, bpp, has_alpha, x_off, y_off, old_w, old_h)
    (64, 48, 4, True, 0, 0, 64, 48),
    (64, 48, 3, False, 0, 0, 64, 48),
    (50, 40, 2, True, 10, 5, 64, 48),      # inside the layer
    (80, 60, 4, True, -7, -9, 64, 48),     # hangs over the top-left
    (80, 60, 1, False, 20, 30, 64, 48),    # hangs over the bottom-right
    (30, 20, 4, True, 100, 0, 64, 48),     # no overlap
]


# ------------------------------------------------------------------------------------------------ CPU
def test_vmap_colour_hand_values():
    # depth 3: seam 1 -> value 3/4, seam 3 -> 1/4; yellow -> red
    vm = np.array([[0, 1, 3]], dtype=np.int32)
    out = plugin_oracle.vmap_colour(vm, 3, (1, 1, 0), (1, 0, 0))
    assert out[0, 0].tolist() == [0, 0, 0, 0]
    assert out[0, 1].tolist() == [255, int(255 * 0.75), 0, int(255 * 0.875)]
    assert out[0, 2].tolist() == [255, int(255 * 0.25), 0, int(255 * 0.625)]


@pytest.mark.parametrize("cs,ce", COLOURS)
def test_vmap_colour_oracle_vs_numpy(cs, ce):
    vm = make_vmap(97, 41, 23, seed=1)
    assert np.array_equal(plugin_oracle.vmap_colour(vm, 23, cs, ce), numpy_vmap_colour(vm, 23, cs, ce))


@pytest.mark.parametrize("case", GUESS_CASES)
@pytest.mark.parametrize("direction", [0, 1])
def test_guess_new_size_oracle_vs_numpy(case, direction):
    w, h, bpp, alpha, xo, yo, ow, oh = case
    mask = make_mask(w, h, bpp, seed=w * 1000 + h)
    assert plugin_oracle.guess_new_size(mask, alpha, xo, yo, ow, oh, direction) == numpy_guess(mask, alpha, xo, yo, ow, oh, direction)


def test_guess_new_size_hand_values():
    mask = np.zeros((4, 6, 4), dtype=np.uint8)
    mask[1, 1:5] = 255   # row 1: four opaque white pixels
    mask[2, 2] = 255     # row 2: one
    mask[3, :, :3] = 255  # row 3: white but fully transparent -> does not count
    assert plugin_oracle.guess_new_size(mask, True, 0, 0, 6, 4, 0) == 6 - 4
    assert plugin_oracle.guess_new_size(mask, True, 0, 0, 6, 4, 1) == 4 - 2  # column 2 holds rows 1 and 2


def golden_digests():
    out = {}
    for i, (cs, ce) in enumerate(COLOURS):
        vm = make_vmap(97, 41, 23, seed=1)
        out[f"vmap_colour_{i}"] = hashlib.sha256(plugin_oracle.vmap_colour(vm, 23, cs, ce).tobytes()).hexdigest()
    for i, case in enumerate(GUESS_CASES):
        w, h, bpp, alpha, xo, yo, ow, oh = case
        mask = make_mask(w, h, bpp, seed=w * 1000 + h)
        out[f"guess_{i}"] = [plugin_oracle.guess_new_size(mask, alpha, xo, yo, ow, oh, d) for d in (0, 1)]
    return out


def test_golden_fixture():
    """tests/golden/plugin_ops.json (written by `python tests/test_plugin_ops.py`) pins the restatement's outputs."""
    assert golden_digests() == json.load(open(GOLDEN))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def ops():
    return importlib.import_module("gimp-lqr-plugin_b200.plugin_ops")


@pytest.mark.gpu
@pytest.mark.parametrize("cs,ce", COLOURS)
@pytest.mark.parametrize("shape", [(97, 41, 23), (1, 1, 1), (640, 360, 200)])
def test_gpu_vmap_colour(ops, cs, ce, shape):
    w, h, depth = shape
    vm = make_vmap(w, h, min(depth, w), seed=w)
    assert np.array_equal(ops.vmap_colour(vm, depth, cs, ce), plugin_oracle.vmap_colour(vm, depth, cs, ce))


@pytest.mark.gpu
def test_gpu_vmap_colour_on_engine_vmap(product, oracle, ops):
    """config 5's output path end to end: a seam map dumped by the engine, coloured by the engine."""
    from cases import V, render, synth
    img = synth.smooth_noise(200, 120, 4)
    res = render.render_noninteractive(product, img, V(new_width=170, new_height=120, output_seams=True))
    vm = res.vmaps[0]
    got = ops.vmap_colour(vm.data, vm.depth, (1, 1, 0), (1, 0, 0))
    assert np.array_equal(got, plugin_oracle.vmap_colour(vm.data, vm.depth, (1, 1, 0), (1, 0, 0)))
    assert (got[..., 3] != 0).sum() == 30 * 120  # every seam pixel, and only those, is drawn


@pytest.mark.gpu
@pytest.mark.parametrize("case", GUESS_CASES + [(1920, 1080, 4, True, -3, 5, 1900, 1100)])
@pytest.mark.parametrize("direction", [0, 1])
def test_gpu_guess_new_size(ops, case, direction):
    w, h, bpp, alpha, xo, yo, ow, oh = case
    mask = make_mask(w, h, bpp, seed=w * 1000 + h)
    assert ops.guess_new_size(mask, alpha, xo, yo, ow, oh, direction) == \
        plugin_oracle.guess_new_size(mask, alpha, xo, yo, ow, oh, direction)


if __name__ == "__main__":
    json.dump(golden_digests(), open(GOLDEN, "w"), indent=1)
    print("wrote", GOLDEN)
