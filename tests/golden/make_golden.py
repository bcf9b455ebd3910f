"""Generates tests/golden/golden.json + golden_small.npz from the CPU ORACLE.

The reference tree ships no golden vectors for this path and liblqr cannot be built here, so
these vectors pin the oracle against ITSELF over time (regression pins) and give the GPU tests an input/output
set that does not need the oracle at run time.  They do NOT pin "oracle == liblqr" (parity unpinned).

    python tests/golden/make_golden.py
"""
import hashlib
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import cases  # noqa: E402

pkg = importlib.import_module("gimp-lqr-plugin_b200")

GOLDEN_CASES = ["rgba_shrink_w", "rgb_shrink_w", "graya_shrink_w", "flat_ties", "shrink_both_vert", "enlarge_multistep",
                "bidirectional_cfg5", "delta_x2", "rigmask", "pres_disc_offset", "cfg3_masks", "lqrback", "to_width_1",
                "energy_fn_0", "energy_fn_4", "energy_fn_6", "batch_scm_cfg1", "mid_dx3_rigmask", "tall"]
SMALL = ["flat_ties", "to_width_1", "tiny_3x3"]  # full arrays stored, not only digests


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def record(res):
    return {"info": res.info, "image_shape": list(res.image.shape), "image_sha256": sha(res.image),
            "aux_sha256": [sha(a) for a in res.aux],
            "vmaps": [{"depth": v.depth, "orientation": v.orientation, "shape": list(v.data.shape),
                       "sha256": sha(v.data.astype(np.int32))} for v in res.vmaps],
            "n_progress": len(res.progress)}


def main():
    oracle = pkg.load_oracle()
    by_name = {c["name"]: c for c in cases.CASES}
    out, arrays = {}, {}
    for name in GOLDEN_CASES + [n for n in SMALL if n not in GOLDEN_CASES]:
        res = cases.run_case(oracle, by_name[name])
        out[name] = record(res)
        if name in SMALL:
            arrays[name + "__image"] = res.image
            for i, v in enumerate(res.vmaps):
                arrays[f"{name}__vmap{i}"] = v.data.astype(np.int32)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **arrays)
    print(f"wrote {len(out)} cases, {len(arrays)} arrays")


if __name__ == "__main__":
    main()
