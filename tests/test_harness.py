"""The plug-in's call sequence replayed in C (tests/harness/plugin_sequence.c) against the LqrCarver API:
CPU: harness(oracle) == Python replay(oracle); GPU: harness(product) == harness(oracle)."""
import importlib

import numpy as np
import pytest

from cases import V, lqr, render, synth

harness = importlib.import_module("gimp-lqr-plugin_b200.harness")

CASES = [
    ("shrink", dict(new_width=100, new_height=90, output_seams=True), False),
    ("both_vert", dict(new_width=100, new_height=80, output_seams=True, res_order=lqr.LQR_RES_ORDER_VERT), False),
    ("enlarge", dict(new_width=140, new_height=95, output_seams=True), False),
    ("masks_rig", dict(new_width=104, new_height=90, delta_x=2, rigidity=4.0, output_seams=True), True),
    ("lqrback", dict(new_width=100, new_height=80, scaleback=True, output_seams=True), True),
]


def _inputs(with_masks):
    img = synth.smooth_noise(120, 90, 4, alpha="random")
    if not with_masks:
        return img, None, None, None
    return img, synth.ellipse_mask(120, 90), synth.iid(120, 90, 4, seed=7), synth.band_mask(120, 90)


@pytest.mark.parametrize("name,kw,masks", CASES, ids=[c[0] for c in CASES])
def test_c_harness_matches_python_replay(pkg, oracle, name, kw, masks):
    img, pres, disc, rig = _inputs(masks)
    vals = V(**kw)
    got, vm, res = harness.render(pkg.ORACLE_PATH, img, vals, pres, disc, rig)
    want = render.render_noninteractive(oracle, img, V(**kw), pres, disc, rig, log_progress=True)
    assert np.array_equal(got, want.image)
    assert res.n_vmaps == len(want.vmaps)
    assert np.array_equal(vm, want.vmaps[0].data)
    assert res.n_progress_updates == sum(1 for k, _ in want.progress if k == "update")


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw,masks", CASES, ids=[c[0] for c in CASES])
def test_c_harness_product_vs_oracle(pkg, product, name, kw, masks):
    img, pres, disc, rig = _inputs(masks)
    a, vma, ra = harness.render(pkg.SHIM_PATH, img, V(**kw), pres, disc, rig)
    b, vmb, rb = harness.render(pkg.ORACLE_PATH, img, V(**kw), pres, disc, rig)
    assert a.shape == b.shape and np.array_equal(a, b)
    assert ra.n_vmaps == rb.n_vmaps and np.array_equal(vma, vmb)
    assert ra.n_progress_updates == rb.n_progress_updates


def test_c_batch_driver_on_oracle(pkg, oracle):
    """harness_render_batch (C host threads, several images in flight) against the oracle library: the driver itself
    must give every image the result of a lone run.  (The oracle is a plain C library without shared state.)"""
    w, h, n = 64, 48, 6
    imgs = [synth.smooth_noise(w, h, 4, seed=900 + i) for i in range(10)]
    vals = V(new_width=w - n, new_height=h)
    outs = harness.render_batch(pkg.ORACLE_PATH, imgs, vals, in_flight=4, keep_outputs=True)["outputs"]
    for img, got in zip(imgs, outs):
        assert np.array_equal(got, render.render_noninteractive(oracle, img, vals).image)
