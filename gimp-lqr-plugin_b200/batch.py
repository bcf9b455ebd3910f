"""Batch front-end: independent images are partitioned across ranks (one process per GPU); there is no
collective inside the algorithm -- a seam depends on every earlier seam of the same image, and one DP row on
the whole previous row -- only an optional gather of the results / digests on rank 0.

Config 4 of BASELINE.json (256 x 1920x1080 RGBA, 100 seams each) runs through this module.
"""
from __future__ import annotations

import hashlib

import numpy as np

from . import render, synth


def shard_indices(n_images: int, world_size: int, rank: int) -> list[int]:
    """image i -> rank i mod world_size (round-robin keeps shards within one image of each other)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    return list(range(rank, n_images, world_size))


def batch_image(index: int, w: int, h: int, channels: int = 4) -> np.ndarray:
    """Deterministic synthetic image `index` of a batch (seed + index)."""
    return synth.smooth_noise(w, h, channels, seed=synth.SEED + index)


def digest(arr: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def carve_shard(lib, indices, w, h, new_w, new_h, channels=4, vals: render.PlugInVals | None = None, in_flight: int = 1):
    """Carves the images of one shard through the LqrCarver API; returns {index: (shape, sha256)}.

    in_flight > 1 keeps that many images in flight on the GPU: one host thread per image, and the engine gives every
    carver its own CUDA stream, so the row-serial DP chains of different images run on different SMs at the same
    time -- the only way this path approaches the bandwidth of the device.  The C calls release
    the GIL; each image's seams stay strictly sequential."""
    import dataclasses
    import threading

    base = vals or render.PlugInVals()
    todo = list(indices)[::-1]
    out, lock, errors = {}, threading.Lock(), []

    def worker():
        while True:
            with lock:
                if not todo or errors:
                    return
                i = todo.pop()
            try:
                v = dataclasses.replace(base, new_width=new_w, new_height=new_h)
                res = render.render_noninteractive(lib, batch_image(i, w, h, channels), v)
                with lock:
                    out[i] = (tuple(res.image.shape), digest(res.image))
            except Exception as e:  # noqa: BLE001 -- reported to the caller below
                with lock:
                    errors.append(e)

    n = max(1, min(int(in_flight), len(todo) or 1))
    if n == 1:
        worker()
    else:
        threads = [threading.Thread(target=worker) for _ in range(n)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if errors:
        raise errors[0]
    return out


def gather_results(local: dict, dist=None, dst: int = 0):
    """Reassembles the per-image results on rank `dst` (torch.distributed gather_object; NCCL or gloo)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(local)
    world = dist.get_world_size()
    gathered = [None] * world if dist.get_rank() == dst else None
    dist.gather_object(local, gathered, dst=dst)
    if dist.get_rank() != dst:
        return None
    merged = {}
    for part in gathered:
        merged.update(part)
    return merged
