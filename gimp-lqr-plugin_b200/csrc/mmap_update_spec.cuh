// mmap_update_spec.cuh -- K2b: incremental m-map DP (liblqr lqr_carver_update_mmap) as a
// warp-specialised, verified-speculative kernel.  One CTA of 13 warps:
//
//   * 8 COMPUTE warps walk the rows.  Per row a thread does one cell (2 or 4 for very wide bands) entirely
//     from shared memory: parents from the previous row's ring (mrow/zrow), the cell's own id / energy /
//     old m / old parent from a tile the producers staged.  They do NOT wait for the exact band limits of
//     the row: they recompute a slightly wider ACTIVE range (the limits verified two rows earlier, grown by
//     2*delta_x and the energy bands in between), apply the keep-old rule to every cell in it and leave the
//     result in place in the tile, plus one ballot word per warp marking the cells whose value changed.
//   * 1 CONTROL warp runs one row behind.  From the ballot words it recomputes liblqr's exact band limits
//     (the leading kept run advances x_min, a trailing kept run pulls x_max back) and VERIFIES the
//     speculation: every changed cell must lie inside the exact band.  It publishes the active / guard
//     ranges two rows ahead.  If the algorithm's band logic is sound -- a cell outside the band has unchanged
//     parents, so recomputing it reproduces the stored value within the keep tolerance -- the check never
//     fires.  If it ever does, nothing wrong has reached HBM: rows are COMMITTED only after verification,
//     and the kernel finishes the remaining rows with the exact generic row loop from the control warp's
//     exact limits.  Results are bit-identical to liblqr's band algorithm in every case.
//   * 4 PRODUCER warps work around the chain: two chunks AHEAD they gather whole chunks of rows (8 rows, or
//     4 when the window is wider than 512 columns) with cp.async -- pixel ids through the raw index table
//     first, then en / m / least through those ids -- and one chunk BEHIND they commit the verified results
//     (m, least of the changed cells) from the tile to HBM.  A chunk's column window is a provable superset
//     of every band, active range, guard range and parent halo its rows can have.
//
// Dependent chain per row on the compute warps: LDS parents -> min/select -> FADD -> keep test -> STS ->
// named barrier (288 threads).  No global memory, no band bookkeeping, no commit, no window arithmetic.
#pragma once
#include "carver_kernels.cuh"
#include "mmap_update_fast.cuh"

namespace b200c {

#define US_NCW 8
#define US_CT (US_NCW * 32)
#define US_NPW 4
#define US_PT (US_NPW * 32)
#define US_THREADS (US_CT + 32 + US_PT)
#define US_TILE 3072
#define US_RW 2048
#define US_RWM (US_RW - 1)
#define US_MAXROWS 8
#define US_MAXCW (US_CT * 4)

// shared-memory words after the tiles and row rings
#define US_NKW (2 * US_MAXROWS * 32) // ballot words  [chunk parity][row][32]
#define US_RIW (2 * US_MAXROWS * 4)  // row info      [chunk parity][row]{guard base, nslots, -, -}
static constexpr size_t us_smem_bytes()
{
    return sizeof(int) * ((size_t) 4 * US_TILE + 9 * US_TILE + 4 * US_RW + US_NKW + US_RIW + 16 + 16 + 8 + 8 + 128);
}

__device__ __forceinline__ void bar_rows() { asm volatile("bar.sync 1, %0;" ::"n"(US_CT + 32) : "memory"); }

// One row of the compute warps.  Thread t handles the columns gr_lo + t + 256*j (j < NS) of the row's guard
// range [gr_lo, gr_hi]: cells inside the active range are recomputed (parents from the previous row's ring),
// the others only forward their old value / id to the ring for the next row's parents.  The new value and
// parent are left in place in the tile (mt / lt) for the commit; `changed` bits go to nkrow.
template <int NS, bool D1>
__device__ __forceinline__ void us_row(const DevP &p, int y, int rbm, const int *__restrict__ zt,
                                       const float *__restrict__ et, float *__restrict__ mt, int *__restrict__ lt,
                                       float *mrow, int *zrow, int cur, int prev, int act_lo, int act_hi, int gr_lo,
                                       int gr_hi, unsigned *nkrow, int tid, int lane, int warp)
{
    const int w = p.w;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        const int x = gr_lo + j * US_CT + tid;
        bool changed = false;
        if (x <= gr_hi) {
            const int idx = rbm + x;
            const int rb = x & US_RWM;
            const int z = zt[idx];
            const float mo = mt[idx];
            float val = mo;
            if (x >= act_lo && x <= act_hi) {
                const float e = et[idx];
                if (y == 0) {
                    val = e; // row 0: m = en over the (exact) band
                    changed = true;
                    mt[idx] = e;
                } else {
                    float best;
                    int parent;
                    if (D1) {
                        const float inf = __int_as_float(0x7f800000);
                        const float m0 = mrow[prev + rb];
                        const int z0 = zrow[prev + rb];
                        float ml = mrow[prev + ((rb - 1) & US_RWM)];
                        const int zl = zrow[prev + ((rb - 1) & US_RWM)];
                        float mr = mrow[prev + ((rb + 1) & US_RWM)];
                        const int zr = zrow[prev + ((rb + 1) & US_RWM)];
                        // left-to-right scan with strict '<' == leftmost minimum; ties go right when leftright == 1.
                        // (all m are finite: an out-of-image neighbour is replaced by +inf and can never win)
                        ml = x > 0 ? ml : inf;
                        mr = x < w - 1 ? mr : inf;
                        best = fminf(fminf(ml, m0), mr);
                        if (p.leftright)
                            parent = mr == best ? zr : (m0 == best ? z0 : zl);
                        else
                            parent = ml == best ? zl : (m0 == best ? z0 : zr);
                    } else {
                        const int D = p.delta_x;
                        const int dlo = max(-x, -D), dhi = min(w - 1 - x, D);
                        int bdx = dlo;
                        best = mrow[prev + ((x + dlo) & US_RWM)];
                        for (int dx = dlo + 1; dx <= dhi; ++dx) {
                            const float cand = mrow[prev + ((x + dx) & US_RWM)];
                            if (cand < best || (cand == best && p.leftright == 1)) {
                                best = cand;
                                bdx = dx;
                            }
                        }
                        parent = zrow[prev + ((x + bdx) & US_RWM)];
                    }
                    const float new_m = __fadd_rn(e, best);
                    // (double) |d| < 1e-5  <=>  |d| <= 0x3727C5AC: that float is the largest one below the double 1e-5
                    const bool keep = (lt[idx] == parent) && (fabsf(__fsub_rn(mo, new_m)) <= __int_as_float(0x3727C5AC));
                    if (!keep) {
                        val = new_m;
                        changed = true;
                        mt[idx] = new_m;
                        lt[idx] = parent;
                    }
                }
            }
            mrow[cur + rb] = val;
            zrow[cur + rb] = z;
        }
        const unsigned word = __ballot_sync(0xffffffffu, changed);
        if (lane == 0) nkrow[j * US_NCW + warp] = word;
    }
}

// commit rows [r_begin, r_end) of a chunk: m (and least, below row 0) of every changed cell, tile -> HBM
__device__ __forceinline__ void us_commit(const DevP &p, const int *dsc, const int *zt, const float *mt, const int *lt,
                                          const unsigned *nkc, const int *ric, int r_begin, int r_end, int t, int nt)
{
    const int y0 = dsc[0], clo = dsc[2], cw = dsc[3];
    for (int r = r_begin; r < r_end; ++r) {
        const int gb = ric[r * 4 + 0], ns = ric[r * 4 + 1];
        const int ncols = ns * US_CT;
        const int rbm = r * cw - clo;
        for (int c = t; c < ncols; c += nt) {
            if ((nkc[r * 32 + (c >> 5)] >> (c & 31)) & 1u) {
                const int idx = rbm + gb + c;
                const int z = zt[idx];
                p.m[z] = mt[idx];
                if (y0 + r > 0) p.least[z] = lt[idx];
            }
        }
    }
}

template <bool D1>
__global__ void __launch_bounds__(US_THREADS, 1) k_mmap_update_spec(DevP p)
{
    extern __shared__ __align__(16) unsigned char us_smem[];
    int *ztile = reinterpret_cast<int *>(us_smem);                    // [4][TILE] pixel ids
    float *etile = reinterpret_cast<float *>(ztile + 4 * US_TILE);    // [3][TILE] energy
    float *mtile = etile + 3 * US_TILE;                               // [3][TILE] old m -> new m of changed cells
    int *ltile = reinterpret_cast<int *>(mtile + 3 * US_TILE);        // [3][TILE] old parent -> new parent
    float *mrow = reinterpret_cast<float *>(ltile + 3 * US_TILE);     // [2][RW] previous / current row values
    int *zrow = reinterpret_cast<int *>(mrow + 2 * US_RW);            // [2][RW] previous / current row ids
    unsigned *nk = reinterpret_cast<unsigned *>(zrow + 2 * US_RW);    // [2][8][32] "changed" ballot words
    int *rinfo = reinterpret_cast<int *>(nk + US_NKW);                // [2][8][4] guard base, slots
    int *pub = rinfo + US_RIW;                                        // [2][8] act_lo, act_hi, fail_row, slots, gr_lo, gr_hi
    int *cdesc = pub + 16;                                            // [4][4] y0, rows, clo, cw
    int *clim = cdesc + 16;                                           // [2][4] x_min, x_max, y_v at chunk starts
    volatile int *misc = clim + 8;                                    // [8] 0 stop, 1 fb_row, 2 fb_xmin, 3 fb_xmax,
                                                                      //     4 last chunk entered, 5 rows verified
    int *s_red = const_cast<int *>(misc) + 8;                         // [128] generic fallback scratch

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = p.delta_x, w = p.w, h = p.h;
    const bool is_compute = tid < US_CT, is_control = warp == US_NCW;

    if (tid == 0) {
        misc[0] = 0;
        misc[1] = h;
        misc[4] = -1;
        misc[5] = 0;
    }

    if (is_compute) {
        // =============================================================================== COMPUTE
        __syncthreads(); // start 1: prologue chunks staged and described
        __syncthreads(); // start 2: ranges of row 0 published
        int y = 0;
        bool failed = false;
        for (int k = 0;; ++k) {
            const int *dsc = cdesc + (k & 3) * 4;
            const int rows = dsc[1], clo = dsc[2], cw = dsc[3];
            if (rows == 0) break;
            if (tid == 0) misc[4] = k;
            const int *zt = ztile + (k & 3) * US_TILE;
            const float *et = etile + (k % 3) * US_TILE;
            float *mt = mtile + (k % 3) * US_TILE;
            int *lt = ltile + (k % 3) * US_TILE;
            unsigned *nkc = nk + (k & 1) * US_MAXROWS * 32;
            int *ric = rinfo + (k & 1) * US_MAXROWS * 4;
            for (int r = 0; r < rows; ++r, ++y) {
                const int par = y & 1;
                // ranges of this row, already clamped to the chunk window by the control warp
                const int4 pa = *reinterpret_cast<const int4 *>(pub + par * 8);
                const int2 pg = *reinterpret_cast<const int2 *>(pub + par * 8 + 4);
                if (pa.z <= y - 2) {
                    failed = true;
                    break;
                }
                const int act_lo = pa.x, act_hi = pa.y, ns = pa.w, gr_lo = pg.x, gr_hi = pg.y;
                const int cur = par * US_RW, prev = (par ^ 1) * US_RW;
                const int rbm = r * cw - clo;
                unsigned *nkrow = nkc + r * 32;
                if (tid == 0) {
                    ric[r * 4 + 0] = gr_lo;
                    ric[r * 4 + 1] = ns;
                }
                if (ns == 1)
                    us_row<1, D1>(p, y, rbm, zt, et, mt, lt, mrow, zrow, cur, prev, act_lo, act_hi, gr_lo, gr_hi, nkrow, tid,
                                  lane, warp);
                else if (ns == 2)
                    us_row<2, D1>(p, y, rbm, zt, et, mt, lt, mrow, zrow, cur, prev, act_lo, act_hi, gr_lo, gr_hi, nkrow, tid,
                                  lane, warp);
                else
                    us_row<4, D1>(p, y, rbm, zt, et, mt, lt, mrow, zrow, cur, prev, act_lo, act_hi, gr_lo, gr_hi, nkrow, tid,
                                  lane, warp);
                bar_rows();
            }
            if (failed) break;
            __syncthreads(); // chunk end: next chunk's tiles are complete
        }
        if (!failed) {
            // drain: the last two rows still wait for verification
            for (int d = 0; d < 2; ++d, ++y) {
                if (pub[(y & 1) * 8 + 2] <= y - 2) break;
                bar_rows();
            }
        }
    } else if (is_control) {
        // =============================================================================== CONTROL
        int x_min = max(p.nrg_xmin[0], 0), x_max = min(p.nrg_xmax[0], w - 1);
        int fail_row = INT_MAX;
        unsigned long long cells = 0;
        // guard range of a row: every column the NEXT row's active range (+- delta_x parents) can need
        auto guard_lo = [&](int xm, int a, int b, int c) { return max(0, min(xm, min(a, min(b, c))) - 4 * D); };
        auto guard_hi = [&](int xm, int a, int b, int c) { return min(w - 1, max(xm, max(a, max(b, c))) + 4 * D); };
        const int n0 = p.nrg_xmin[0], n1 = p.nrg_xmin[min(1, h - 1)], n2 = p.nrg_xmin[min(2, h - 1)];
        const int m0 = p.nrg_xmax[0], m1 = p.nrg_xmax[min(1, h - 1)], m2 = p.nrg_xmax[min(2, h - 1)];
        // publish the ranges of row `yr` (window clo/cw of its chunk): active range from limits + two rows of
        // growth, guard range from four; both clamped to the staged window; slots per thread from the guard width
        auto publish = [&](int yr, int a_lo, int a_hi, int g_lo, int g_hi, int clo_r, int cw_r, int fail) {
            const int gl = max(g_lo, clo_r), gh = min(g_hi, clo_r + cw_r - 1);
            const int gw = gh - gl + 1;
            int *pb = pub + (yr & 1) * 8;
            pb[0] = max(a_lo, gl);
            pb[1] = min(a_hi, gh);
            pb[2] = fail;
            pb[3] = gw <= US_CT ? 1 : (gw <= 2 * US_CT ? 2 : 4);
            pb[4] = gl;
            pb[5] = gh;
        };
        if (lane == 0) {
            clim[0] = x_min;
            clim[1] = x_max;
            clim[2] = 0;
        }
        // rolling window of the energy-band limits: a*0 = row y-1, a*1 = row y, a*2 = row y+1, a*3 = row y+2
        int an0 = 0, an1 = n0, an2 = n1, an3 = n2;
        int ax0 = 0, ax1 = m0, ax2 = m1, ax3 = m2;
        __syncthreads(); // start 1: the producers described chunks 0..2
        if (lane == 0) // row 0: the active range is the exact band (m = en there)
            publish(0, x_min, x_max, guard_lo(x_min, n0, n0, n1), guard_hi(x_max, m0, m0, m1), cdesc[2], cdesc[3], INT_MAX);
        __syncthreads(); // start 2
        int y = 0;
        bool stop = false;
        // nkv / riv: ballot words and row info of the row being verified (row y-1); clo_n / cw_n: window of the
        // chunk that holds row y+1
        auto iteration = [&](const unsigned *nkv, const int *riv, int y_lim, int clo_n, int cw_n) {
            // runs while the compute warps process row y: verify row y-1, publish the ranges of row y+1
            const int an4 = p.nrg_xmin[min(y + 3, h - 1)], ax4 = p.nrg_xmax[min(y + 3, h - 1)];
            const int yv = y - 1;
            if (yv >= 1 && yv < h && yv < y_lim && fail_row == INT_MAX) {
                const int bmin = max(min(x_min, an0) - D, 0);
                const int bmax = min(max(x_max, ax0) + D, w - 1);
                const int gb = riv[0], nwords = riv[1] * US_NCW;
                const unsigned wv = lane < nwords ? nkv[lane] : 0u;
                const unsigned any = __ballot_sync(0xffffffffu, wv != 0u);
                int F = INT_MAX, L = INT_MIN;
                if (any) {
                    const int lf = __ffs(any) - 1, ll = 31 - __clz(any);
                    const unsigned wf = __shfl_sync(0xffffffffu, wv, lf), wl = __shfl_sync(0xffffffffu, wv, ll);
                    F = gb + 32 * lf + (__ffs(wf) - 1);
                    L = gb + 32 * ll + (31 - __clz(wl));
                }
                const int old_min = x_min, old_max = x_max;
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
                        misc[1] = yv;
                        misc[2] = old_min;
                        misc[3] = old_max;
                        __threadfence_block();
                        misc[0] = 1;
                    }
                }
            }
            // active range of row y+1 from the limits after row y-1 (two rows of growth) and its guard range
            if (lane == 0) {
                publish(y + 1, max(0, min(x_min, min(an1, an2)) - 2 * D), min(w - 1, max(x_max, max(ax1, ax2)) + 2 * D),
                        guard_lo(x_min, an1, an2, an3), guard_hi(x_max, ax1, ax2, ax3), clo_n, cw_n, fail_row);
                if (fail_row == INT_MAX) misc[5] = min(y, y_lim); // rows [0, y) are verified (producers commit them)
                __threadfence_block();
            }
            an0 = an1, an1 = an2, an2 = an3, an3 = an4;
            ax0 = ax1, ax1 = ax2, ax2 = ax3, ax3 = ax4;
        };
        // The loop mirrors the compute warps' exactly (same stop predicate at the top of every row), so both
        // roles execute the same sequence of row and chunk barriers.
        const unsigned *nk_last = nk;
        const int *ri_last = rinfo;
        for (int k = 0;; ++k) {
            const int *dsc = cdesc + (k & 3) * 4;
            const int rows = dsc[1], clo = dsc[2], cw = dsc[3];
            if (rows == 0) break;
            const int *dn = cdesc + ((k + 1) & 3) * 4; // next chunk (described at least one chunk ago)
            const int clo_nx = dn[2], cw_nx = dn[3];
            const unsigned *nkc = nk + (k & 1) * US_MAXROWS * 32;
            const int *ric = rinfo + (k & 1) * US_MAXROWS * 4;
            for (int r = 0; r < rows; ++r, ++y) {
                if (fail_row <= y - 2) {
                    stop = true;
                    break;
                }
                const bool nx = r + 1 >= rows; // row y+1 opens the next chunk
                if (r == 0) // row y-1 is the last row of the previous chunk
                    iteration(nk_last, ri_last, INT_MAX, nx ? clo_nx : clo, nx ? cw_nx : cw);
                else
                    iteration(nkc + (r - 1) * 32, ric + (r - 1) * 4, INT_MAX, nx ? clo_nx : clo, nx ? cw_nx : cw);
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
            nk_last = nkc + (rows - 1) * 32;
            ri_last = ric + (rows - 1) * 4;
            __syncthreads(); // chunk end
        }
        if (!stop) {
            const int y_end = y;
            for (int d = 0; d < 2; ++d, ++y) {
                if (fail_row <= y - 2) {
                    stop = true;
                    break;
                }
                iteration(nk_last, ri_last, y_end, 0, 0); // d == 0 verifies row y_end-1; d == 1 has nothing to verify
                bar_rows();
            }
            if (fail_row == INT_MAX && lane == 0 && y_end < h) {
                // capacity stop: rows below y_end are exact; hand the exact limits to the generic loop
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
            for (int want = US_MAXROWS; want >= 2; want >>= 1) {
                const int yb = min(ya + want, h) - 1;
                // energy-band extremes over rows [yv+1, yb+2] (two rows per lane: up to 64 rows)
                const int last = min(yb + 2, h - 1);
                const int j0 = yv + 1 + lane, j1 = j0 + 32;
                int nlo = j0 <= last ? p.nrg_xmin[j0] : INT_MAX, nhi = j0 <= last ? p.nrg_xmax[j0] : INT_MIN;
                if (j1 <= last) {
                    nlo = min(nlo, p.nrg_xmin[j1]);
                    nhi = max(nhi, p.nrg_xmax[j1]);
                }
                nlo = __reduce_min_sync(0xffffffffu, nlo);
                nhi = __reduce_max_sync(0xffffffffu, nhi);
                // a band grows by at most delta_x per row beyond the energy bands; the guard range of the last
                // row reaches 4 rows of growth past the limits verified two rows before it
                const int dist = (yb + 3 - yv) * D;
                lo = max(0, min(xv_min, nlo) - dist);
                const int hi = min(w - 1, max(xv_max, nhi) + dist);
                cw = max(hi - lo + 1, 0);
                rows = yb - ya + 1;
                if (rows * cw <= US_TILE && cw <= US_MAXCW && yb + 1 - yv <= 63) return;
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
        auto stage_a = [&](int kk, int ya, int rows, int lo, int cw) {
            int *zt = ztile + (kk & 3) * US_TILE;
            for (int r = 0; r < rows; ++r) {
                const int *src = p.raw + (size_t) (ya + r) * p.raw_stride + lo;
                for (int c = pt; c < cw; c += US_PT) cp_async4(&zt[r * cw + c], src + c);
            }
        };
        auto stage_b = [&](int kk, int rows, int cw) {
            const int *zt = ztile + (kk & 3) * US_TILE;
            float *et = etile + (kk % 3) * US_TILE, *mt = mtile + (kk % 3) * US_TILE;
            int *lt = ltile + (kk % 3) * US_TILE;
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
        // geometry of the chunks in flight: c1 = chunk k+1 (data being staged), c2 = chunk k+2 (ids staged)
        int ya1, rows1, lo1, cw1, ya2, rows2, lo2, cw2;
        {
            // ---- prologue: chunk 0 and 1 fully staged, ids of chunk 2 staged
            int rows0, lo0, cw0;
            plan_regs(0, b0min, b0max, 0, rows0, lo0, cw0);
            put_desc(0, 0, rows0, lo0, cw0);
            stage_a(0, 0, rows0, lo0, cw0);
            cp_async_commit();
            ya1 = rows0, rows1 = 0, lo1 = 0, cw1 = 0;
            if (rows0 > 0) plan_regs(ya1, b0min, b0max, 0, rows1, lo1, cw1);
            put_desc(1, ya1, rows1, lo1, cw1);
            stage_a(1, ya1, rows1, lo1, cw1);
            cp_async_commit();
            ya2 = ya1 + rows1, rows2 = 0, lo2 = 0, cw2 = 0;
            if (rows1 > 0) plan_regs(ya2, b0min, b0max, 0, rows2, lo2, cw2);
            put_desc(2, ya2, rows2, lo2, cw2);
            stage_a(2, ya2, rows2, lo2, cw2);
            cp_async_commit();
            cp_async_wait<0>();
            stage_b(0, rows0, cw0);
            stage_b(1, rows1, cw1);
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads(); // start 1
            __syncthreads(); // start 2
            // ---- steady state.  During chunk k: commit chunk k-1 (verified), then stage en/m/least of chunk k+2 and
            // the ids of chunk k+3.  The loads waited for were issued a whole chunk earlier.
            int rows_k = rows0;
            for (int k = 0;; ++k) {
                if (rows_k == 0) {
                    __syncthreads(); // matches the exit barrier of the other roles
                    break;
                }
                if (k > 0) {
                    // chunk k-1: its tiles are about to be recycled (data tile by stage_b(k+2), id tile by
                    // stage_a(k+3)), so its verified rows go to HBM first.  Its last row is verified during row 0
                    // of chunk k.
                    const int *dp = cdesc + ((k - 1) & 3) * 4;
                    const int yp0 = dp[0], prow = dp[1];
                    int done;
                    while ((done = misc[5]) < yp0 + prow && misc[0] == 0) __nanosleep(32);
                    const int r_end = min(prow, max(misc[0] ? min(done, (int) misc[1]) - yp0 : prow, 0));
                    us_commit(p, dp, ztile + ((k - 1) & 3) * US_TILE, mtile + ((k - 1) % 3) * US_TILE,
                              ltile + ((k - 1) % 3) * US_TILE, nk + ((k - 1) & 1) * US_MAXROWS * 32,
                              rinfo + ((k - 1) & 1) * US_MAXROWS * 4, 0, r_end, pt, US_PT);
                    asm volatile("bar.sync 2, %0;" ::"n"(US_PT) : "memory"); // every producer is done with the old tiles
                }
                cp_async_wait<0>(); // ids of chunk k+2 and data of chunk k+1, issued during chunk k-1
                stage_b(k + 2, rows2, cw2);
                cp_async_commit();
                // chunk k+3: plan from the limits published at the end of chunk k-1 (or the initial band)
                const int *cl = clim + (k & 1) * 4;
                const int ya3 = ya2 + rows2;
                int rows3 = 0, lo3 = 0, cw3 = 0;
                if (rows2 > 0) plan_regs(ya3, cl[0], cl[1], cl[2], rows3, lo3, cw3);
                stage_a(k + 3, ya3, rows3, lo3, cw3);
                cp_async_commit();
                put_desc((k + 3) & 3, ya3, rows3, lo3, cw3);
                __syncthreads(); // chunk k end (or the exit barrier of a failed speculation)
                if (misc[0]) break;
                rows_k = rows1;
                ya1 = ya2, rows1 = rows2, lo1 = lo2, cw1 = cw2;
                ya2 = ya3, rows2 = rows3, lo2 = lo3, cw2 = cw3;
            }
        }
        cp_async_wait<0>();
        (void) lo1;
        (void) cw1;
        (void) ya1;
    }

    // =================================================================================== exit / fallback
    if (tid < US_CT + 32) __syncthreads(); // compute + control: exit barrier (the producers already passed theirs)
    // commit what the producers have not: the verified rows of the last chunk the compute warps entered
    const int fb_row = misc[1], klast = misc[4];
    if (klast >= 0) {
        const int *dl = cdesc + (klast & 3) * 4;
        const int r_end = min(dl[1], max(min(fb_row, h) - dl[0], 0));
        us_commit(p, dl, ztile + (klast & 3) * US_TILE, mtile + (klast % 3) * US_TILE, ltile + (klast % 3) * US_TILE,
                  nk + (klast & 1) * US_MAXROWS * 32, rinfo + (klast & 1) * US_MAXROWS * 4, 0, r_end, tid, US_THREADS);
    }
    if (fb_row < h) {
        __syncthreads();
        update_rows_generic(p, fb_row, misc[2], misc[3], s_red);
    }
}

} // namespace b200c
