// mmap_update_spec.cuh -- K2b v3: incremental m-map DP (liblqr lqr_carver_update_mmap) as a
// warp-specialised, verified-speculative kernel.  One CTA of 13 warps:
//
//   * 8 COMPUTE warps walk the rows.  Per row a thread does one cell (up to 4 for very wide bands) entirely
//     from shared memory: parents from the previous row's ring (mrow/zrow), the cell's own id / energy /
//     old m / old parent from a tile the producers staged.  They do NOT wait for the exact band limits of
//     the row: they process a slightly wider ACTIVE range (the limits verified two rows earlier, grown by
//     2*delta_x and the energy bands in between) and apply the keep-old rule to every cell in it.
//   * 1 CONTROL warp runs one row behind.  From the ballot words of "value changed" cells it recomputes
//     liblqr's exact band limits (leading kept run advances x_min, trailing kept run pulls x_max back) and
//     VERIFIES the speculation: every changed cell must lie inside the exact band.  It publishes the active
//     range two rows ahead.  If the algorithm's band logic is sound -- a cell outside the band has unchanged
//     parents, so recomputing it reproduces the stored value within the keep tolerance -- the check never
//     fires.  If it ever does, nothing wrong has been written: results are COMMITTED to HBM by the compute
//     threads two rows late, only for verified rows, and the kernel finishes the remaining rows with the
//     exact generic row loop from the control warp's exact limits.  Results are therefore bit-identical to
//     liblqr's band algorithm in every case.
//   * 4 PRODUCER warps gather whole chunks of rows (8 rows, or 4 when the window is wider than 512 columns)
//     two chunks ahead with cp.async: pixel ids through the raw index table first, then en / m / least
//     through those ids.  A chunk's column window is a provable superset of every band (and active range,
//     and parent halo) its rows can have, computed from the limits verified at planning time.
//
// Dependent chain per row on the compute warps: LDS parents -> compare/select -> FADD -> keep test -> STS
// -> named barrier (288 threads).  No global memory, no band bookkeeping, no window arithmetic on the chain.
#pragma once
#include "carver_kernels.cuh"
#include "mmap_update_fast.cuh"

namespace b200c {

#define US_NCW 8
#define US_CT (US_NCW * 32)
#define US_NPW 4
#define US_PT (US_NPW * 32)
#define US_THREADS (US_CT + 32 + US_PT)
#define US_TILE 4096
#define US_RW 2048
#define US_RWM (US_RW - 1)
#define US_NSLOT 4
#define US_MAXCW (US_CT * US_NSLOT)

static constexpr size_t us_smem_bytes()
{
    return sizeof(int) * ((size_t) 3 * US_TILE + 6 * US_TILE + 4 * US_RW + 64 + 8 + 16 + 8 + 8 + 128);
}

__device__ __forceinline__ void bar_rows() { asm volatile("bar.sync 1, %0;" ::"n"(US_CT + 32) : "memory"); }

template <bool D1>
__global__ void __launch_bounds__(US_THREADS, 1) k_mmap_update_spec(DevP p)
{
    extern __shared__ __align__(16) unsigned char us_smem[];
    int *ztile = reinterpret_cast<int *>(us_smem);                    // [3][TILE] pixel ids
    float *etile = reinterpret_cast<float *>(ztile + 3 * US_TILE);    // [2][TILE] energy
    float *mtile = etile + 2 * US_TILE;                               // [2][TILE] old m
    int *ltile = reinterpret_cast<int *>(mtile + 2 * US_TILE);        // [2][TILE] old parent
    float *mrow = reinterpret_cast<float *>(ltile + 2 * US_TILE);     // [2][RW] previous / current row values
    int *zrow = reinterpret_cast<int *>(mrow + 2 * US_RW);            // [2][RW] previous / current row ids
    unsigned *nk = reinterpret_cast<unsigned *>(zrow + 2 * US_RW);    // [2][32] "changed" ballot words
    int *pub = reinterpret_cast<int *>(nk + 64);                      // [2][4] act_lo, act_hi, fail_row
    int *cdesc = pub + 8;                                             // [4][4] y0, rows, clo, cw
    int *clim = cdesc + 16;                                           // [2][4] x_min, x_max, y_v at chunk starts
    int *misc = clim + 8;                                             // [8] 0 stop, 1 fb_row, 2 fb_xmin, 3 fb_xmax
    int *s_red = misc + 8;                                            // [128] generic fallback scratch

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = p.delta_x, w = p.w, h = p.h, lr = p.leftright;
    const bool is_compute = tid < US_CT, is_control = warp == US_NCW;

    if (tid == 0) {
        misc[0] = 0;
        misc[1] = h;
    }

    if (is_compute) {
        // =============================================================================== COMPUTE
        int hz1[US_NSLOT], hp1[US_NSLOT], hz2[US_NSLOT], hp2[US_NSLOT];
        float hm1[US_NSLOT], hm2[US_NSLOT];
#pragma unroll
        for (int j = 0; j < US_NSLOT; ++j) hz1[j] = hz2[j] = -1, hp1[j] = hp2[j] = -2, hm1[j] = hm2[j] = 0.f;
        __syncthreads(); // start: prologue chunks staged, pub[0] published
        int y = 0;
        bool failed = false;
        for (int k = 0;; ++k) {
            const int *dsc = cdesc + (k & 3) * 4;
            const int rows = dsc[1], clo = dsc[2], cw = dsc[3];
            if (rows == 0) break;
            const int nslot = (cw + US_CT - 1) / US_CT;
            const int *zt = ztile + (k % 3) * US_TILE;
            const float *et = etile + (k & 1) * US_TILE;
            const float *mt = mtile + (k & 1) * US_TILE;
            const int *lt = ltile + (k & 1) * US_TILE;
            for (int r = 0; r < rows; ++r, ++y) {
                const int par = y & 1;
                const int act_lo = pub[par * 4 + 0], act_hi = pub[par * 4 + 1], fail_row = pub[par * 4 + 2];
                if (fail_row <= y - 2) {
                    failed = true;
                    break;
                }
                // commit the results of row y-2 (verified by now), then age the history
#pragma unroll
                for (int j = 0; j < US_NSLOT; ++j) {
                    if (hz2[j] >= 0) {
                        p.m[hz2[j]] = hm2[j];
                        if (hp2[j] != -2) p.least[hz2[j]] = hp2[j];
                    }
                    hz2[j] = hz1[j];
                    hm2[j] = hm1[j];
                    hp2[j] = hp1[j];
                    hz1[j] = -1;
                }
                const int cur = par * US_RW, prev = (par ^ 1) * US_RW;
                const int rbase = r * cw;
#pragma unroll
                for (int j = 0; j < US_NSLOT; ++j) {
                    if (j < nslot) {
                        const int c = j * US_CT + tid;
                        bool changed = false;
                        if (c < cw) {
                            const int x = clo + c;
                            const int rb = x & US_RWM;
                            const int z = zt[rbase + c];
                            const float mo = mt[rbase + c];
                            float val = mo;
                            if (x >= act_lo && x <= act_hi) {
                                const float e = et[rbase + c];
                                if (y == 0) {
                                    val = e;
                                    hz1[j] = z;
                                    hm1[j] = e;
                                    hp1[j] = -2;
                                } else {
                                    float best;
                                    int parent;
                                    if (D1) {
                                        const float m0 = mrow[prev + rb];
                                        const int z0 = zrow[prev + rb];
                                        const float ml = mrow[prev + ((rb - 1) & US_RWM)];
                                        const int zl = zrow[prev + ((rb - 1) & US_RWM)];
                                        const float mr = mrow[prev + ((rb + 1) & US_RWM)];
                                        const int zr = zrow[prev + ((rb + 1) & US_RWM)];
                                        best = m0;
                                        parent = z0;
                                        if (x > 0) {
                                            best = ml;
                                            parent = zl;
                                            if (m0 < best || (m0 == best && lr == 1)) {
                                                best = m0;
                                                parent = z0;
                                            }
                                        }
                                        if (x < w - 1 && (mr < best || (mr == best && lr == 1))) {
                                            best = mr;
                                            parent = zr;
                                        }
                                    } else {
                                        const int dlo = max(-x, -D), dhi = min(w - 1 - x, D);
                                        int bdx = dlo;
                                        best = mrow[prev + ((x + dlo) & US_RWM)];
                                        for (int dx = dlo + 1; dx <= dhi; ++dx) {
                                            const float cand = mrow[prev + ((x + dx) & US_RWM)];
                                            if (cand < best || (cand == best && lr == 1)) {
                                                best = cand;
                                                bdx = dx;
                                            }
                                        }
                                        parent = zrow[prev + ((x + bdx) & US_RWM)];
                                    }
                                    const float new_m = __fadd_rn(e, best);
                                    const bool keep =
                                        (lt[rbase + c] == parent) && ((double) fabsf(__fsub_rn(mo, new_m)) < 1e-5);
                                    if (!keep) {
                                        val = new_m;
                                        changed = true;
                                        hz1[j] = z;
                                        hm1[j] = new_m;
                                        hp1[j] = parent;
                                    }
                                }
                            }
                            mrow[cur + rb] = val;
                            zrow[cur + rb] = z;
                        }
                        const unsigned word = __ballot_sync(0xffffffffu, changed);
                        if (lane == 0) nk[par * 32 + j * US_NCW + warp] = word;
                    }
                }
                bar_rows();
            }
            if (failed) break;
            __syncthreads(); // chunk end: next chunk's tiles are complete
        }
        if (!failed) {
            // drain: rows y_end-2 and y_end-1 still wait for verification
            for (int d = 0; d < 2; ++d, ++y) {
                const int par = y & 1;
                const int fail_row = pub[par * 4 + 2];
                if (fail_row <= y - 2) break;
#pragma unroll
                for (int j = 0; j < US_NSLOT; ++j) {
                    if (hz2[j] >= 0) {
                        p.m[hz2[j]] = hm2[j];
                        if (hp2[j] != -2) p.least[hz2[j]] = hp2[j];
                    }
                    hz2[j] = hz1[j];
                    hm2[j] = hm1[j];
                    hp2[j] = hp1[j];
                    hz1[j] = -1;
                }
                bar_rows();
            }
        }
    } else if (is_control) {
        // =============================================================================== CONTROL
        int x_min = max(p.nrg_xmin[0], 0), x_max = min(p.nrg_xmax[0], w - 1);
        int fail_row = INT_MAX;
        unsigned long long cells = 0;
        if (lane == 0) {
            pub[0] = x_min; // row 0: the active range is the exact band (m = en there)
            pub[1] = x_max;
            pub[2] = INT_MAX;
            clim[0] = x_min;
            clim[1] = x_max;
            clim[2] = 0;
        }
        // rolling window of the energy-band limits: a*[0] = row y-1, [1] = row y, [2] = row y+1
        int an0 = 0, an1 = p.nrg_xmin[0], an2 = p.nrg_xmin[min(1, h - 1)];
        int ax0 = 0, ax1 = p.nrg_xmax[0], ax2 = p.nrg_xmax[min(1, h - 1)];
        __syncthreads(); // start
        int y = 0;
        int clo_prev = 0, cw_prev = 0;
        bool stop = false;
        int fb_xmin = x_min, fb_xmax = x_max;
        auto iteration = [&](int clo_v, int cw_v, int y_lim) {
            // runs while the compute warps process row y: verify row y-1, publish the active range of row y+1
            const int an3 = p.nrg_xmin[min(y + 2, h - 1)], ax3 = p.nrg_xmax[min(y + 2, h - 1)];
            const int yv = y - 1;
            if (yv >= 1 && yv < h && yv < y_lim && fail_row == INT_MAX) {
                const int bmin = max(min(x_min, an0) - D, 0);
                const int bmax = min(max(x_max, ax0) + D, w - 1);
                const int nwords = (cw_v + 31) >> 5;
                const unsigned wv = lane < nwords ? nk[(yv & 1) * 32 + lane] : 0u;
                const unsigned any = __ballot_sync(0xffffffffu, wv != 0u);
                int F = INT_MAX, L = INT_MIN;
                if (any) {
                    const int lf = __ffs(any) - 1, ll = 31 - __clz(any);
                    const unsigned wf = __shfl_sync(0xffffffffu, wv, lf), wl = __shfl_sync(0xffffffffu, wv, ll);
                    F = clo_v + 32 * lf + (__ffs(wf) - 1);
                    L = clo_v + 32 * ll + (31 - __clz(wl));
                }
                fb_xmin = x_min;
                fb_xmax = x_max;
                bool violation;
                if (bmax >= bmin) {
                    cells += (unsigned long long) (bmax - bmin + 1);
                    violation = any && (F < bmin || L > bmax);
                    x_min = any ? F : bmax + 1;
                    x_max = any ? (L == bmax ? bmax : L + 1) : bmin;
                } else {
                    violation = any != 0u;
                    x_min = bmin;
                    x_max = bmax;
                }
                if (violation) {
                    fail_row = yv;
                    if (lane == 0) {
                        misc[0] = 1;
                        misc[1] = yv;
                        misc[2] = fb_xmin;
                        misc[3] = fb_xmax;
                    }
                }
            }
            if (lane == 0) {
                // active range of row y+1 from the limits after row y-1: two rows of growth
                int *pb = pub + ((y + 1) & 1) * 4;
                pb[0] = max(0, min(x_min, min(an1, an2)) - 2 * D);
                pb[1] = min(w - 1, max(x_max, max(ax1, ax2)) + 2 * D);
                pb[2] = fail_row;
            }
            an0 = an1, an1 = an2, an2 = an3;
            ax0 = ax1, ax1 = ax2, ax2 = ax3;
        };
        // The loop mirrors the compute warps' exactly (same stop predicate at the top of every row), so both
        // roles execute the same sequence of row and chunk barriers.
        int y_end = INT_MAX;
        for (int k = 0;; ++k) {
            const int *dsc = cdesc + (k & 3) * 4;
            const int rows = dsc[1], clo = dsc[2], cw = dsc[3];
            if (rows == 0) break;
            for (int r = 0; r < rows; ++r, ++y) {
                if (fail_row <= y - 2) {
                    stop = true;
                    break;
                }
                if (r == 0)
                    iteration(clo_prev, cw_prev, y_end);
                else
                    iteration(clo, cw, y_end);
                if (r == rows - 1 && lane == 0) {
                    // limits the producers plan chunk k+3 from (they read them after the chunk barrier)
                    int *cl = clim + ((k + 1) & 1) * 4;
                    cl[0] = x_min;
                    cl[1] = x_max;
                    cl[2] = max(y - 1, 0);
                }
                bar_rows();
            }
            if (stop) break;
            clo_prev = clo;
            cw_prev = cw;
            __syncthreads(); // chunk end
        }
        if (!stop) {
            y_end = y;
            for (int d = 0; d < 2; ++d, ++y) {
                if (fail_row <= y - 2) {
                    stop = true;
                    break;
                }
                iteration(clo_prev, cw_prev, y_end);
                bar_rows();
            }
            if (fail_row == INT_MAX && lane == 0 && y_end < h) {
                // capacity stop: rows below y_end are exact and committed; hand over the exact limits
                misc[1] = y_end;
                misc[2] = x_min;
                misc[3] = x_max;
            }
        }
        if (lane == 0 && p.cells) atomicAdd(p.cells, cells);
    } else {
        // =============================================================================== PRODUCERS
        const int pt = tid - (US_CT + 32);
        // plan the chunk starting at row ya from the limits (xv_min, xv_max) valid after row yv.
        // rows == 0: nothing to stage (image done, or the window exceeds the staging capacity -> the fast
        // region ends before this chunk).  Every producer lane computes the same values.
        auto plan_regs = [&](int ya, int xv_min, int xv_max, int yv, int &rows, int &lo, int &cw) {
            rows = 0, lo = 0, cw = 0;
            if (ya >= h) return;
            for (int attempt = 0; attempt < 2; ++attempt) {
                const int want = attempt == 0 ? 8 : 4;
                const int yb = min(ya + want, h) - 1;
                // energy-band extremes over rows [yv+1, yb+1] (one row per lane)
                const int j = yv + 1 + lane;
                const bool in = j <= min(yb + 1, h - 1);
                int nlo = in ? p.nrg_xmin[j] : INT_MAX, nhi = in ? p.nrg_xmax[j] : INT_MIN;
                nlo = __reduce_min_sync(0xffffffffu, nlo);
                nhi = __reduce_max_sync(0xffffffffu, nhi);
                // a band grows by at most delta_x per row beyond the energy bands; +1 row for the parent halo
                const int dist = (yb + 2 - yv) * D;
                lo = max(0, min(xv_min, nlo) - dist);
                const int hi = min(w - 1, max(xv_max, nhi) + dist);
                cw = max(hi - lo + 1, 0);
                rows = yb - ya + 1;
                if (cw <= (attempt == 0 ? 512 : US_MAXCW) && yb - yv <= 31) return;
                rows = 0;
            }
        };
        auto put_desc = [&](int slot, int ya, int rows, int lo, int cw) {
            if (pt == 0) {
                int *d = cdesc + slot * 4;
                d[0] = ya;
                d[1] = rows;
                d[2] = lo;
                d[3] = cw;
            }
        };
        auto stage_a = [&](int ztile_idx, int ya, int rows, int lo, int cw) {
            int *zt = ztile + ztile_idx * US_TILE;
            for (int r = 0; r < rows; ++r) {
                const int *src = p.raw + (size_t) (ya + r) * p.raw_stride + lo;
                for (int c = pt; c < cw; c += US_PT) cp_async4(&zt[r * cw + c], src + c);
            }
        };
        auto stage_b = [&](int ztile_idx, int dtile_idx, int rows, int cw) {
            const int *zt = ztile + ztile_idx * US_TILE;
            float *et = etile + dtile_idx * US_TILE, *mt = mtile + dtile_idx * US_TILE;
            int *lt = ltile + dtile_idx * US_TILE;
            for (int r = 0; r < rows; ++r)
                for (int c = pt; c < cw; c += US_PT) {
                    const int i = r * cw + c;
                    const int z = zt[i]; // staged by this same thread, complete after cp.async.wait_group
                    cp_async4(&et[i], p.en + z);
                    cp_async4(&mt[i], p.m + z);
                    cp_async4(&lt[i], p.least + z);
                }
        };
        const int b0min = max(p.nrg_xmin[0], 0), b0max = min(p.nrg_xmax[0], w - 1);
        int ya0, rows0, lo0, cw0, ya1, rows1, lo1, cw1;
        // ---- prologue
        ya0 = 0;
        plan_regs(ya0, b0min, b0max, 0, rows0, lo0, cw0);
        put_desc(0, ya0, rows0, lo0, cw0);
        stage_a(0, ya0, rows0, lo0, cw0);
        cp_async_commit();
        cp_async_wait<0>();
        stage_b(0, 0, rows0, cw0);
        cp_async_commit();
        ya1 = ya0 + rows0;
        rows1 = 0, lo1 = 0, cw1 = 0;
        if (rows0 > 0) plan_regs(ya1, b0min, b0max, 0, rows1, lo1, cw1);
        put_desc(1, ya1, rows1, lo1, cw1);
        stage_a(1, ya1, rows1, lo1, cw1);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads(); // start
        // ---- steady state: during chunk k, stage en/m/least of chunk k+1 and the ids of chunk k+2
        int rows_k = rows0; // chunk k
        for (int k = 0;; ++k) {
            if (rows_k == 0) {
                __syncthreads(); // matches the exit barrier of the other roles
                break;
            }
            // chunk k+1: ids are complete (waited before the last barrier)
            stage_b((k + 1) % 3, (k + 1) & 1, rows1, cw1);
            cp_async_commit();
            // chunk k+2: plan from the limits published at the end of chunk k-1 (or the initial band)
            const int *cl = clim + (k & 1) * 4;
            const int ya2 = ya1 + rows1;
            int rows2 = 0, lo2 = 0, cw2 = 0;
            if (rows1 > 0) plan_regs(ya2, cl[0], cl[1], cl[2], rows2, lo2, cw2);
            stage_a((k + 2) % 3, ya2, rows2, lo2, cw2);
            cp_async_commit();
            put_desc((k + 2) & 3, ya2, rows2, lo2, cw2);
            cp_async_wait<0>();
            __syncthreads(); // chunk k end (or the exit barrier of a failed speculation)
            if (misc[0]) break;
            rows_k = rows1;
            ya1 = ya2, rows1 = rows2, lo1 = lo2, cw1 = cw2;
        }
        cp_async_wait<0>();
    }

    // =================================================================================== exit / fallback
    if (!(tid >= US_CT + 32)) __syncthreads(); // compute + control: exit barrier (producers already passed theirs)
    const int fb_row = misc[1];
    if (fb_row < h) update_rows_generic(p, fb_row, misc[2], misc[3], s_red);
}

} // namespace b200c
