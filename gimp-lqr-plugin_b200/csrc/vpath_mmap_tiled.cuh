// vpath_mmap_tiled.cuh -- K3 (seam backtrack, liblqr lqr_carver_build_vpath) and K2 (full m-map
// DP, lqr_carver_build_mmap, A.5) restructured for the B200: both are row-serial chains whose generic
// versions pay dependent L2 round trips on every row; here the chain only ever touches shared memory.
#pragma once
#include "carver_kernels.cuh"
#include "mmap_update_fast.cuh"

namespace b200c {

// ------------------------------------------------------------------------------------------------ K3
// The seam moves at most delta_x columns per row, so R rows of it stay inside a window of half-width R*delta_x
// around its position at the chunk's top row.  Warps 1.. gather the NEXT chunk's window (ids through the raw
// table, then parents through `least`, then resolve every cell's parent column) into the spare buffer -- its
// centre is the seam position at the current chunk's top, hence the doubled half-width 2*R*delta_x -- while
// lane 0 of warp 0 walks the CURRENT chunk: one dependent shared-memory load per row instead of three
// dependent global loads.
#define VP_THREADS 512
#define VP_CAP 6656 // entries per buffer: (R+1) * (4*R*delta_x + 1) <= VP_CAP

struct VpGeom {
    int R, HW, WC;
};

__host__ __device__ inline VpGeom vp_geometry(int delta_x)
{
    VpGeom g;
    int R = 64;
    while (R > 1 && (R + 1) * (4 * R * delta_x + 1) > VP_CAP) --R;
    g.R = R;
    g.HW = 2 * R * delta_x;
    g.WC = 2 * g.HW + 1;
    return g;
}

__device__ __forceinline__ void vp_load_chunk(const DevP &p, const VpGeom g, int top, int wl, int *zwin, int *par,
                                              short *nxt, int lt, int nl)
{
    const int rows_z = min(g.R + 1, top + 1);
    const int n = rows_z * g.WC;
    // pass 1: pixel ids of the window (rows top, top-1, ..., top-R)
#pragma unroll 4
    for (int idx = lt; idx < n; idx += nl) {
        const int ri = idx / g.WC, ci = idx - ri * g.WC;
        const int x = wl + ci, y = top - ri;
        zwin[idx] = (x >= 0 && x < p.w) ? p.raw[(size_t) y * p.raw_stride + x] : -1;
    }
    // pass 2: parent id of every cell that has a row above it inside the chunk (own entries only)
#pragma unroll 4
    for (int idx = lt; idx < n; idx += nl) {
        const int ri = idx / g.WC;
        const int z = zwin[idx];
        par[idx] = (z >= 0 && ri < g.R && top - ri > 0) ? p.least[z] : -1;
    }
    asm volatile("bar.sync 1, %0;" ::"r"(nl) : "memory"); // loader warps only: zwin complete
    // pass 3: column of the parent in the row above (search x-delta_x..x+delta_x, first match), else stay
    for (int idx = lt; idx < n; idx += nl) {
        const int ri = idx / g.WC, ci = idx - ri * g.WC;
        int res = ci;
        const int pz = par[idx];
        if (pz >= 0 && ri + 1 < rows_z) {
            const int x = wl + ci;
            const int x_lo = max(x - p.delta_x, 0), x_hi = min(x + p.delta_x, p.w - 1);
            const int *up = zwin + (ri + 1) * g.WC;
            for (int xx = x_lo; xx <= x_hi; ++xx) {
                const int cc = xx - wl;
                if (cc >= 0 && cc < g.WC && up[cc] == pz) {
                    res = cc;
                    break;
                }
            }
        }
        nxt[idx] = (short) res;
    }
}

__global__ void __launch_bounds__(VP_THREADS, 1) k_vpath_fast(DevP p)
{
    extern __shared__ __align__(16) unsigned char vp_smem[];
    int *zwin_b = reinterpret_cast<int *>(vp_smem);            // [2][VP_CAP]
    int *par_b = zwin_b + 2 * VP_CAP;                          // [2][VP_CAP]
    short *nxt_b = reinterpret_cast<short *>(par_b + 2 * VP_CAP); // [2][VP_CAP]
    __shared__ float s_v[32];
    __shared__ int s_x[32];
    __shared__ int s_cx[2], s_last[2], s_wl[2]; // double buffered by chunk parity

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = VP_THREADS / 32;
    const VpGeom g = vp_geometry(p.delta_x);
    const int h = p.h;

    // ---- arg-min over the last row (A.6 tie rule)
    const int *row = p.raw + (size_t) (h - 1) * p.raw_stride;
    float best = 536870912.f;
    int bx = -1;
    for (int x = tid; x < p.w; x += VP_THREADS) {
        const float v = p.m[row[x]];
        if (v < best || (v == best && p.leftright == 1)) {
            best = v;
            bx = x;
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_down_sync(0xffffffffu, best, off);
        const int ox = __shfl_down_sync(0xffffffffu, bx, off);
        if (seam_better(ov, ox, best, bx, p.leftright)) {
            best = ov;
            bx = ox;
        }
    }
    if (lane == 0) {
        s_v[warp] = best;
        s_x[warp] = bx;
    }
    __syncthreads();
    if (tid == 0) {
        for (int i = 1; i < nwarp; ++i)
            if (seam_better(s_v[i], s_x[i], best, bx, p.leftright)) {
                best = s_v[i];
                bx = s_x[i];
            }
        const int last_x = bx < 0 ? 0 : bx;
        s_cx[0] = last_x;
        s_last[0] = row[last_x];
        s_wl[0] = last_x - g.HW;
    }
    __syncthreads();

    // ---- chunk 0 is gathered by the loader warps alone, then the pipeline starts
    const int nl = VP_THREADS - 32, lt = tid - 32;
    if (warp != 0) vp_load_chunk(p, g, h - 1, s_wl[0], zwin_b, par_b, nxt_b, lt, nl);
    __syncthreads();

    int top = h - 1;
    for (int c = 0; top >= 0; ++c, top -= g.R) {
        const int b = c & 1;
        const int cx = s_cx[b]; // seam column at row `top`
        if (warp == 0) {
            if (lane == 0) {
                const int *par = par_b + b * VP_CAP;
                const short *nxt = nxt_b + b * VP_CAP;
                const int wl = s_wl[b];
                int ci = cx - wl, last = s_last[b];
                const int steps = min(g.R, top + 1);
                for (int ri = 0; ri < steps; ++ri) {
                    const int y = top - ri;
                    p.vpath[y] = last;
                    p.vpath_x[y] = wl + ci;
                    if (y > 0) {
                        if (ci < 0 || ci >= g.WC) { // cannot happen: the window is twice the seam's reach
                            atomicOr(p.err, 2);
                            ci = min(max(ci, 0), g.WC - 1);
                        }
                        const int idx = ri * g.WC + ci;
                        last = par[idx];
                        ci = nxt[idx];
                    }
                }
                s_cx[b ^ 1] = wl + ci; // position at row top - R (the next chunk's top)
                s_last[b ^ 1] = last;
            }
        } else if (top - g.R >= 0) {
            const int wl_next = cx - g.HW;
            if (tid == 32) s_wl[b ^ 1] = wl_next;
            vp_load_chunk(p, g, top - g.R, wl_next, zwin_b + (b ^ 1) * VP_CAP, par_b + (b ^ 1) * VP_CAP,
                          nxt_b + (b ^ 1) * VP_CAP, lt, nl);
        }
        __syncthreads();
    }
}

static inline size_t vp_smem_bytes() { return (size_t) VP_CAP * 2 * (4 + 4 + 2); }

// ------------------------------------------------------------------------------------------------ K2
// Full DP, one launch per block of K rows (the launch boundary is the grid-wide barrier between dependent row
// blocks).  CTA b owns S columns and redundantly recomputes a halo of K*delta_x columns on each side (the
// cone of dependence of its strip over K rows), so inside a launch no CTA waits on another.  The block's
// pixel ids / energies are staged in shared memory with two bulk cp.async passes before the row loop, which
// then runs from shared memory with one barrier per row.
__global__ void __launch_bounds__(1024, 1) k_mmap_full_tile(DevP p, int y0, int K, int S)
{
    extern __shared__ __align__(16) unsigned char mf_smem[];
    const int T = blockDim.x, t = threadIdx.x, D = p.delta_x, w = p.w, h = p.h;
    const bool has_rigmask = p.use_rig && p.rigmask != nullptr;
    int *zst = reinterpret_cast<int *>(mf_smem);                       // [K][T]
    float *est = reinterpret_cast<float *>(zst + K * T);               // [K][T]
    float *rst = est + K * T;                                          // [K][T] when rigmask
    float *mrow = rst + (has_rigmask ? K * T : 0);                     // [2][T]
    int *zrow = reinterpret_cast<int *>(mrow + 2 * T);                 // [2][T]
    float *rigsm = reinterpret_cast<float *>(zrow + 2 * T);            // [2*D+1]

    const int c0 = blockIdx.x * S, c1 = min(c0 + S, w);
    const int x = c0 - K * D + t;
    const bool inimg = x >= 0 && x < w;
    const int rows = min(K, h - y0);
    const int lr = p.leftright;

    if (p.use_rig)
        for (int i = t; i <= 2 * D; i += T) rigsm[i] = p.rigmap[i - D];
    const float *rigc = rigsm + D;

    if (inimg) {
        const int *src = p.raw + (size_t) y0 * p.raw_stride + x;
        for (int r = 0; r < rows; ++r) cp_async4(&zst[r * T + t], src + (size_t) r * p.raw_stride);
    }
    cp_async_commit();
    float m_in = 0.f;
    int z_in = -1;
    if (inimg && y0 > 0) {
        z_in = p.raw[(size_t) (y0 - 1) * p.raw_stride + x];
        m_in = p.m[z_in];
    }
    cp_async_wait<0>();
    if (inimg) {
        for (int r = 0; r < rows; ++r) {
            const int z = zst[r * T + t];
            cp_async4(&est[r * T + t], p.en + z);
            if (has_rigmask) cp_async4(&rst[r * T + t], p.rigmask + z);
        }
    }
    cp_async_commit();
    mrow[T + t] = m_in; // buffer 1 is "previous" for r == 0
    zrow[T + t] = z_in;
    cp_async_wait<0>();
    __syncthreads();

    for (int r = 0; r < rows; ++r) {
        const int y = y0 + r;
        const int cur = (r & 1) * T, prev = ((r & 1) ^ 1) * T;
        float val = 0.f;
        int z = -1;
        // columns whose whole cone of dependence (back to the block's input row) lies inside this CTA's window
        const bool valid = inimg && t >= (r + 1) * D && t <= T - 1 - (r + 1) * D;
        if (inimg) z = zst[r * T + t];
        if (y == 0) {
            if (inimg) {
                val = est[r * T + t];
                if (x >= c0 && x < c1) p.m[z] = val;
            }
        } else if (valid) {
            const int dlo = max(-x, -D), dhi = min(w - 1 - x, D);
            int bdx = dlo;
            float best;
            if (p.use_rig) {
                const float rf = has_rigmask ? rst[r * T + t] : 1.f;
                best = __fadd_rn(mrow[prev + t + dlo], __fmul_rn(rf, rigc[dlo]));
                for (int dx = dlo + 1; dx <= dhi; ++dx) {
                    const float cand = __fadd_rn(mrow[prev + t + dx], __fmul_rn(rf, rigc[dx]));
                    if (cand < best || (cand == best && lr == 1)) {
                        best = cand;
                        bdx = dx;
                    }
                }
            } else {
                best = mrow[prev + t + dlo];
                for (int dx = dlo + 1; dx <= dhi; ++dx) {
                    const float cand = mrow[prev + t + dx];
                    if (cand < best || (cand == best && lr == 1)) {
                        best = cand;
                        bdx = dx;
                    }
                }
            }
            val = __fadd_rn(est[r * T + t], best);
            if (x >= c0 && x < c1) {
                p.m[z] = val;
                p.least[z] = zrow[prev + t + bdx];
            }
        }
        mrow[cur + t] = val;
        zrow[cur + t] = z;
        __syncthreads();
    }
}

struct MfGeom {
    int K, S, T;
    size_t smem;
};

static inline MfGeom mf_geometry(int delta_x, int w, bool has_rigmask)
{
    MfGeom g;
    g.S = 32;
    g.K = delta_x <= 1 ? 64 : (64 / delta_x < 4 ? 4 : 64 / delta_x);
    g.T = g.S + 2 * g.K * delta_x;
    g.T = (g.T + 31) / 32 * 32;
    (void) w;
    g.smem = (size_t) g.K * g.T * 4 * (has_rigmask ? 3 : 2) + (size_t) 4 * g.T * 4 + (size_t) (2 * delta_x + 1) * 4 + 64;
    return g;
}

} // namespace b200c
