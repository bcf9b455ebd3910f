// mmap_update_fast.cuh -- K2b, the incremental m-map DP after one carve (liblqr lqr_carver_update_mmap,
// SURVEY.md A.8), written for the B200 memory system.
//
// The algorithm is a row-serial chain (row y needs row y-1's values AND the band limits that row y-1's
// keep/replace decisions produced), so one CTA walks the rows.  What makes the generic version slow is that
// every row pays several dependent L2 round trips: raw[y][x] -> en/m/least[z] -> raw[y-1][x+dx] -> m[zd].
// Here nothing on the chain touches global memory:
//   * the previous row's values and pixel ids live in shared memory (mrow/zrow, double buffered);
//   * everything a row needs that does NOT depend on the chain (its pixel ids, energy, old m, old parent,
//     rigidity factor) is gathered 2*KP / KP rows ahead with cp.async (LDGSTS) into shared-memory rings --
//     a two-level gather: stage A fetches the row's pixel ids through the raw index table, stage B uses
//     those ids to fetch en/m/least.  The columns fetched for a future row are a provable superset of the
//     band that row can have: the band grows by at most delta_x per row beyond the energy bands, whose
//     sliding extremes (pre_lo/pre_hi) the band-energy kernel computes in parallel beforehand.
// Results are bit-identical to the generic kernel: same scan order, same tie rule, same keep-old rule, same
// band trimming.  If a predicted window does not fit the staging capacity the kernel finishes the remaining
// rows with the generic row loop (same CTA, no relaunch).
#pragma once
#include "carver_kernels.cuh"

namespace b200c {

#define UF_THREADS 256
#define UF_CPT 4
#define UF_WIN (UF_THREADS * UF_CPT) // staged columns per row
#define UF_RW 2048                   // ring width (columns) of the previous-row buffers
#define UF_RWM (UF_RW - 1)
#define UF_MAX_DELTA 32

__device__ __forceinline__ void cp_async4(void *smem, const void *gmem)
{
    const unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// generic continuation: rows y_start..h-1 with all operands in global memory (the v1 row loop).
// (x_min, x_max) are the band limits left by row y_start-1 (or anything when y_start == 0).
__device__ __forceinline__ void update_rows_generic(const DevP &p, int y_start, int x_min, int x_max, int *s_red)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    if (y_start == 0) {
        x_min = max(p.nrg_xmin[0], 0);
        x_max = min(p.nrg_xmax[0], p.w - 1);
        for (int x = x_min + tid; x <= x_max; x += blockDim.x) {
            const int z = p.raw[x];
            p.m[z] = p.en[z];
        }
        __syncthreads();
        y_start = 1;
    }
    for (int y = y_start; y < p.h; ++y) {
        x_min = min(x_min, p.nrg_xmin[y]);
        x_max = max(x_max, p.nrg_xmax[y]);
        x_min = max(x_min - p.delta_x, 0);
        x_max = min(x_max + p.delta_x, p.w - 1);
        const int *row = p.raw + (size_t) y * p.raw_stride;
        int first = INT_MAX, last = INT_MIN;
        for (int x = x_min + tid; x <= x_max; x += blockDim.x) {
            const int z = row[x];
            int parent;
            const float new_m = __fadd_rn(p.en[z], best_parent(p, x, y, z, parent));
            const bool keep = (p.least[z] == parent) && ((double) fabsf(__fsub_rn(p.m[z], new_m)) < 1e-5);
            if (!keep) {
                p.m[z] = new_m;
                first = min(first, x);
                last = max(last, x);
            }
            p.least[z] = parent;
        }
        first = __reduce_min_sync(0xffffffffu, first);
        last = __reduce_max_sync(0xffffffffu, last);
        int *red = s_red + (y & 1) * 64;
        if (lane == 0) {
            red[warp] = first;
            red[32 + warp] = last;
        }
        __syncthreads();
        int F = INT_MAX, L = INT_MIN;
        for (int i = 0; i < nwarp; ++i) {
            F = min(F, red[i]);
            L = max(L, red[32 + i]);
        }
        if (x_max >= x_min) {
            const int nx_min = (F != INT_MAX) ? F : x_max + 1;
            const int nx_max = (L != INT_MIN) ? (L == x_max ? x_max : L + 1) : x_min;
            x_min = nx_min;
            x_max = nx_max;
        }
    }
}

template <int KP, bool RIG>
struct UfLayout {
    static constexpr int NZ = 2 * KP + 1; // rows of pixel ids in flight
    static constexpr int ND = KP + 1;     // rows of en/m/least in flight
    static constexpr size_t bytes =
        sizeof(int) * ((size_t) NZ * UF_WIN + (size_t) (3 + (RIG ? 1 : 0)) * ND * UF_WIN + 4 * UF_RW + 2 * NZ + 128 +
                       2 * UF_MAX_DELTA + 1);
};

// pre_lo[y] / pre_hi[y]: min of nrg_xmin / max of nrg_xmax over rows [y-2*KP, y+1] (clamped), written by
// k_energy_band_pre.
template <int KP, bool RIG>
__global__ void __launch_bounds__(UF_THREADS, 1)
k_mmap_update_fast(DevP p, const int *__restrict__ pre_lo, const int *__restrict__ pre_hi)
{
    using LY = UfLayout<KP, RIG>;
    constexpr int NZ = LY::NZ, ND = LY::ND;
    extern __shared__ __align__(16) unsigned char uf_smem[];
    int *zs = reinterpret_cast<int *>(uf_smem);                // [NZ][WIN]
    float *es = reinterpret_cast<float *>(zs + NZ * UF_WIN);   // [ND][WIN]
    float *ms = es + ND * UF_WIN;                              // [ND][WIN]
    int *ls = reinterpret_cast<int *>(ms + ND * UF_WIN);       // [ND][WIN]
    float *rs = reinterpret_cast<float *>(ls + ND * UF_WIN);   // [ND][WIN] when RIG
    float *mrow = rs + (RIG ? ND * UF_WIN : 0);                // [2][RW]
    int *zrow = reinterpret_cast<int *>(mrow + 2 * UF_RW);     // [2][RW]
    int *wlo = zrow + 2 * UF_RW;                               // [NZ]
    int *wn = wlo + NZ;                                        // [NZ]
    int *s_red = wn + NZ;                                      // [2][64]
    float *rigsm = reinterpret_cast<float *>(s_red + 128);     // [2*delta_x+1], centred at [delta_x]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = p.delta_x, w = p.w, h = p.h;
    const int M = (2 * KP + 3) * D; // how far a band can outgrow the prediction inputs, plus the parent halo
    const bool has_rigmask = RIG && p.rigmask != nullptr;
    if (RIG) {
        for (int i = tid; i <= 2 * D; i += UF_THREADS) rigsm[i] = p.rigmap[i - D];
    }
    const float *rigc = rigsm + D;
    const int lr = p.leftright;

    int x_min = max(p.nrg_xmin[0], 0);
    int x_max = min(p.nrg_xmax[0], w - 1);
    int fail_row = INT_MAX;
    int it = -2 * KP;
    unsigned long long cells = 0;
    __syncthreads();

    for (; it < h; ++it) {
        // ---- band limits left by the previous row (its keep/replace reduction)
        if (it >= 2) {
            const int *red = s_red + ((it - 1) & 1) * 64;
            int F = INT_MAX, L = INT_MIN;
#pragma unroll
            for (int i = 0; i < UF_THREADS / 32; ++i) {
                F = min(F, red[i]);
                L = max(L, red[32 + i]);
            }
            // limits row it-1 was processed with
            const int pmin = max(min(x_min, p.nrg_xmin[it - 1]) - D, 0);
            const int pmax = min(max(x_max, p.nrg_xmax[it - 1]) + D, w - 1);
            x_min = pmin;
            x_max = pmax;
            if (pmax >= pmin) {
                x_min = (F != INT_MAX) ? F : pmax + 1;
                x_max = (L != INT_MIN) ? (L == pmax ? pmax : L + 1) : pmin;
            }
        }
        // ---- (A) pixel ids of row it + 2*KP
        const int ya = it + 2 * KP;
        if (ya < h) {
            const int lo = max(0, min(x_min, pre_lo[ya]) - M);
            const int hi = min(w - 1, max(x_max, pre_hi[ya]) + M);
            int n = hi - lo + 1;
            if (n > UF_WIN) {
                fail_row = min(fail_row, ya);
                n = 0;
            }
            if (n < 0) n = 0;
            const int ring = ya % NZ;
            if (tid == 0) {
                wlo[ring] = lo;
                wn[ring] = n;
            }
            const int *src = p.raw + (size_t) ya * p.raw_stride + lo;
            for (int slot = tid; slot < n; slot += UF_THREADS) cp_async4(&zs[ring * UF_WIN + slot], src + slot);
        }
        cp_async_commit();
        // ---- (B) en / m / least (/ rigidity factor) of row it + KP through the ids fetched KP rows ago
        const int yb = it + KP;
        if (yb >= 0 && yb < h) {
            cp_async_wait<2 * KP>();
            const int n = wn[yb % NZ];
            const int *zsrc = zs + (yb % NZ) * UF_WIN;
            const int dring = (yb % ND) * UF_WIN;
            for (int slot = tid; slot < n; slot += UF_THREADS) {
                const int z = zsrc[slot];
                cp_async4(&es[dring + slot], p.en + z);
                cp_async4(&ms[dring + slot], p.m + z);
                cp_async4(&ls[dring + slot], p.least + z);
                if (has_rigmask) cp_async4(&rs[dring + slot], p.rigmask + z);
            }
        }
        cp_async_commit();
        // ---- (C) row `it`
        if (it >= 0) {
            if (it >= fail_row) break;
            cp_async_wait<2 * KP>();
            const int y = it;
            const int lo = wlo[y % NZ], n = wn[y % NZ];
            const int *zsrc = zs + (y % NZ) * UF_WIN;
            const int dring = (y % ND) * UF_WIN;
            int bmin = x_min, bmax = x_max;
            if (y > 0) {
                bmin = max(min(x_min, p.nrg_xmin[y]) - D, 0);
                bmax = min(max(x_max, p.nrg_xmax[y]) + D, w - 1);
            }
            const int cur = (y & 1) * UF_RW, prev = ((y & 1) ^ 1) * UF_RW;
            // invariant of the window prediction: this row's band lies inside its staged window and its
            // parents lie inside the previous row's window.  Cheap to check, fatal if ever violated.
            if (tid == 0 && bmax >= bmin) {
                cells += (unsigned long long) (bmax - bmin + 1);
                bool ok = bmin >= lo && bmax <= lo + n - 1;
                if (y > 0) {
                    const int plo = wlo[(y - 1) % NZ], pn = wn[(y - 1) % NZ];
                    ok = ok && max(bmin - D, 0) >= plo && min(bmax + D, w - 1) <= plo + pn - 1;
                }
                if (!ok) atomicOr(p.err, 1);
            }
            int first = INT_MAX, last = INT_MIN;
            for (int slot = tid; slot < n; slot += UF_THREADS) {
                const int x = lo + slot;
                const int z = zsrc[slot];
                const float mo = ms[dring + slot];
                float val = mo;
                if (x >= bmin && x <= bmax) {
                    const float e = es[dring + slot];
                    if (y == 0) {
                        val = e;
                        p.m[z] = e;
                    } else {
                        const int dlo = max(-x, -D), dhi = min(w - 1 - x, D);
                        int bdx = dlo;
                        float best;
                        if (RIG) {
                            const float rf = has_rigmask ? rs[dring + slot] : 1.f;
                            best = __fadd_rn(mrow[prev + ((x + dlo) & UF_RWM)], __fmul_rn(rf, rigc[dlo]));
                            for (int dx = dlo + 1; dx <= dhi; ++dx) {
                                const float cand =
                                    __fadd_rn(mrow[prev + ((x + dx) & UF_RWM)], __fmul_rn(rf, rigc[dx]));
                                if (cand < best || (cand == best && lr == 1)) {
                                    best = cand;
                                    bdx = dx;
                                }
                            }
                        } else {
                            best = mrow[prev + ((x + dlo) & UF_RWM)];
                            for (int dx = dlo + 1; dx <= dhi; ++dx) {
                                const float cand = mrow[prev + ((x + dx) & UF_RWM)];
                                if (cand < best || (cand == best && lr == 1)) {
                                    best = cand;
                                    bdx = dx;
                                }
                            }
                        }
                        const int parent = zrow[prev + ((x + bdx) & UF_RWM)];
                        const int lold = ls[dring + slot];
                        const float new_m = __fadd_rn(e, best);
                        const bool keep = (lold == parent) && ((double) fabsf(__fsub_rn(mo, new_m)) < 1e-5);
                        if (!keep) {
                            p.m[z] = new_m;
                            val = new_m;
                            first = min(first, x);
                            last = max(last, x);
                        }
                        if (lold != parent) p.least[z] = parent;
                    }
                }
                mrow[cur + (x & UF_RWM)] = val;
                zrow[cur + (x & UF_RWM)] = z;
            }
            if (y > 0) {
                first = __reduce_min_sync(0xffffffffu, first);
                last = __reduce_max_sync(0xffffffffu, last);
                if (lane == 0) {
                    int *red = s_red + (y & 1) * 64;
                    red[warp] = first;
                    red[32 + warp] = last;
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0 && p.cells) atomicAdd(p.cells, cells);
    if (it < h) {
        // a predicted window exceeded the staging capacity: finish with the generic row loop
        cp_async_wait<0>();
        __syncthreads();
        update_rows_generic(p, it, x_min, x_max, s_red);
    }
}

// K1b + prediction inputs: like k_energy_band, and additionally writes for every row y
//   * pre_lo / pre_hi: the extremes of the energy bands of rows [y - span, y + 1] (staged kernel, span = 2*KP);
//   * ctab[2*y + side]: {n[y-1], ext(n[y], n[y+1]), ext(n[y..y+2]), 0} of the energy-band limits n, side 0 = minima,
//     side 1 = maxima NEGATED -- the three values the control warp of the speculative kernel needs for row y
//     (rows clamped to the image).
__global__ void __launch_bounds__(256) k_energy_band_pre(DevP p, int span, int *pre_lo, int *pre_hi, int4 *ctab)
{
    const int lane = threadIdx.x & 31;
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (y >= p.h) return;
    const int r = p.nrg_radius;
    // lane l looks at row y - span + l (l <= span + 2 <= 31), clamped to the image: its energy band
    const int yy = min(max(y - span + lane, 0), p.h - 1);
    int bmin, bmax;
    {
        const int own = p.vpath_x[yy];
        int xmin = own, xmax = own - 1;
        for (int y1 = max(yy - r, 0); y1 <= min(yy + r, p.h - 1); ++y1) {
            const int x = p.vpath_x[y1];
            xmin = min(xmin, x - r);
            xmax = max(xmax, x + r - 1);
        }
        bmin = max(0, xmin);
        bmax = min(p.w - 1, xmax);
    }
    // lane `span` holds row y itself
    const int xmin = __shfl_sync(0xffffffffu, bmin, span);
    const int xmax = __shfl_sync(0xffffffffu, bmax, span);
    const bool in_pre = lane <= span + 1 && y - span + lane >= 0 && y - span + lane < p.h;
    const int lo = __reduce_min_sync(0xffffffffu, in_pre ? bmin : INT_MAX);
    const int hi = __reduce_max_sync(0xffffffffu, in_pre ? bmax : INT_MIN);
    const int nm1 = __shfl_sync(0xffffffffu, bmin, span - 1), xm1 = __shfl_sync(0xffffffffu, bmax, span - 1);
    const int np1 = __shfl_sync(0xffffffffu, bmin, span + 1), xp1 = __shfl_sync(0xffffffffu, bmax, span + 1);
    const int np2 = __shfl_sync(0xffffffffu, bmin, span + 2), xp2 = __shfl_sync(0xffffffffu, bmax, span + 2);
    if (lane == 0) {
        p.nrg_xmin[y] = xmin;
        p.nrg_xmax[y] = xmax;
        pre_lo[y] = lo;
        pre_hi[y] = hi;
        if (ctab) {
            ctab[2 * y + 0] = make_int4(nm1, min(xmin, np1), min(xmin, min(np1, np2)), 0);
            ctab[2 * y + 1] = make_int4(-xm1, -max(xmax, xp1), -max(xmax, max(xp1, xp2)), 0);
        }
    }
    for (int x = xmin + lane; x <= xmax; x += 32) {
        const int z = p.raw[(size_t) y * p.raw_stride + x];
        p.en[z] = energy_at(p, x, y);
    }
}

} // namespace b200c
