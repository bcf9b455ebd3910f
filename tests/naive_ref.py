"""First-principles numpy restatements used to pin the ORACLE (they share no code with it).

They cover the parts of the algorithm that can be stated without liblqr's incremental machinery:
the energy formula, one full DP + backtrack (A.5, A.6), and -- for energies that are
exact small integers attached to pixels, where the 1e-5 keep-old rule of the incremental update cannot
matter -- the whole shrink loop by brute force (full DP before every seam).
"""
import numpy as np


def brightness(img: np.ndarray, luma: bool = False) -> np.ndarray:
    img = img.astype(np.float64) / 255.0
    c = img.shape[2]
    if c <= 2:
        v = img[:, :, 0]
    elif luma:
        v = 0.2126 * img[:, :, 0] + 0.7152 * img[:, :, 1] + 0.0722 * img[:, :, 2]
    else:
        v = (img[:, :, 0] + img[:, :, 1] + img[:, :, 2]) / 3
    if c in (2, 4):
        v = v * img[:, :, c - 1]
    return v


def energy(img: np.ndarray, ef: int) -> np.ndarray:
    """ef: 0 norm, 1 sumabs, 2 xabs, 3-5 same on luma, 6 null.  float32 (h, w)."""
    h, w = img.shape[:2]
    if ef == 6:
        return np.zeros((h, w), np.float32)
    b = brightness(img, luma=ef >= 3)
    gx = np.zeros_like(b)
    gy = np.zeros_like(b)
    if w > 1:
        gx[:, 0] = b[:, 1] - b[:, 0]
        gx[:, -1] = b[:, -1] - b[:, -2]
        gx[:, 1:-1] = (b[:, 2:] - b[:, :-2]) / 2
    else:
        gx[:, 0] = 0 - b[:, 0]
    if h > 1:
        gy[0, :] = b[1, :] - b[0, :]
        gy[-1, :] = b[-1, :] - b[-2, :]
        gy[1:-1, :] = (b[2:, :] - b[:-2, :]) / 2
    else:
        gy[0, :] = 0 - b[0, :]
    k = ef % 3
    if k == 0:
        e = np.sqrt(gx * gx + gy * gy)
    elif k == 1:
        e = (np.abs(gx) + np.abs(gy)) / 2
    else:
        e = np.abs(gx)
    return e.astype(np.float32)


def dp_seam(en: np.ndarray, delta_x: int = 1, leftright: int = 0, rig=None) -> np.ndarray:
    """One full float32 DP + backtrack.  Returns x[y].  rig: optional (rigmap[2*dx+1], rfact (h,w))."""
    h, w = en.shape
    m = np.zeros((h, w), np.float32)
    parent = np.zeros((h, w), np.int64)
    m[0] = en[0]
    for y in range(1, h):
        for x in range(w):
            best = None
            for dx in range(max(-x, -delta_x), min(w - 1 - x, delta_x) + 1):
                cand = m[y - 1, x + dx]
                if rig is not None:
                    cand = np.float32(cand + np.float32(rig[1][y, x] * rig[0][dx + delta_x]))
                if best is None or cand < best or (cand == best and leftright == 1):
                    best, bx = cand, x + dx
            m[y, x] = np.float32(en[y, x] + best)
            parent[y, x] = bx
    best, bx = np.float32(2 ** 29), 0
    for x in range(w):
        v = m[h - 1, x]
        if v < best or (v == best and leftright == 1):
            best, bx = v, x
    xs = np.zeros(h, np.int64)
    for y in range(h - 1, -1, -1):
        xs[y] = bx
        bx = parent[y, bx]
    return xs


def dp_seam_fast(en: np.ndarray, leftright: int = 0) -> np.ndarray:
    """Vectorised delta_x = 1 version of dp_seam (no rigidity)."""
    h, w = en.shape
    m = np.zeros((h, w), np.float32)
    par = np.zeros((h, w), np.int64)
    m[0] = en[0]
    inf = np.float32(np.inf)
    xs_idx = np.arange(w)
    for y in range(1, h):
        up = m[y - 1]
        left = np.concatenate(([inf], up[:-1]))
        right = np.concatenate((up[1:], [inf]))
        cand = np.stack([left, up, right])  # scan order: x-1, x, x+1
        if leftright == 0:
            k = np.argmin(cand, axis=0)  # first minimum
        else:
            k = 2 - np.argmin(cand[::-1], axis=0)  # last minimum
        best = cand[k, xs_idx]
        m[y] = (en[y] + best).astype(np.float32)
        par[y] = xs_idx + k - 1
    last = m[h - 1]
    mn = last.min()
    cands = np.nonzero(last == mn)[0]
    bx = cands[0] if leftright == 0 else cands[-1]
    xs = np.zeros(h, np.int64)
    for y in range(h - 1, -1, -1):
        xs[y] = bx
        bx = par[y, bx]
    return xs


def brute_force_vmap(E: np.ndarray, n_seams: int, switch_frequency: int = 0) -> np.ndarray:
    """Shrink by n_seams with per-pixel energies E (exact small integers, they travel with the pixel).
    Returns the seam-order map at the original size: 0 = untouched, k = k-th seam removed."""
    h, w = E.shape
    cur = E.astype(np.float32).copy()
    ids = np.tile(np.arange(w), (h, 1))
    vmap = np.zeros((h, w), np.int64)
    leftright = 0
    interval = ((n_seams + 1) - 1 - 1) // switch_frequency + 1 if switch_frequency else 0
    for k in range(1, n_seams + 1):
        xs = dp_seam_fast(cur, leftright)
        for y in range(h):
            vmap[y, ids[y, xs[y]]] = k
        keep = np.ones(cur.shape, bool)
        keep[np.arange(h), xs] = False
        cur = cur[keep].reshape(h, -1)
        ids = ids[keep].reshape(h, -1)
        if switch_frequency and ((k - 1 + interval // 2) % interval) == 0:
            leftright ^= 1
    return vmap
