// mmap_update_tma.cuh -- K2b, the incremental m-map DP after one carve (liblqr lqr_carver_update_mmap,
// SURVEY.md A.8): warp-specialised, verified-speculative, staged and committed with TMA bulk copies.
//
// One CTA of 14 warps on one SM (the algorithm is a row-serial chain):
//
//   * 12 COMPUTE warps walk the rows.  Per row a thread does one cell (2 or 3 for very wide windows) entirely from
//     shared memory: parents from the previous row's ring (mrow/zrow), the cell's own id / energy / old m / old
//     parent from the chunk tile.  They do NOT wait for the exact band limits of the row: they recompute a
//     slightly wider ACTIVE range (the limits verified two rows earlier, grown by 2*delta_x and the energy bands
//     in between), apply liblqr's keep-old rule to every cell in it, leave the result in place in the tile and
//     emit one ballot word per warp marking the cells whose value changed.
//   * 1 CONTROL warp runs one row behind.  From the ballot words it recomputes liblqr's exact band limits (the
//     leading kept run advances x_min, a trailing kept run pulls x_max back) and VERIFIES the speculation:
//     every changed cell must lie inside the exact band.  It publishes the active / guard ranges two rows ahead.
//     If the band logic is sound -- a cell outside the band has unchanged parents, so recomputing it reproduces
//     the stored value within the keep tolerance -- the check never fires.  If it ever does, nothing wrong has
//     reached HBM: rows are COMMITTED only after verification, and the kernel finishes the remaining rows with
//     the exact generic row loop from the control warp's exact limits.  Bit-identical to liblqr in every case.
//   * 1 DMA warp moves the data with the TMA.  Pixel ids along a row are increasing and nearly contiguous (the
//     x-th visible pixel of row y is y*w0 + x + #seams removed on its left), so the en / m / least values of a
//     row window occupy one nearly contiguous PHYSICAL span.  Per row the warp issues four cp.async.bulk loads
//     (the raw-id window and the three spans, 16-byte aligned) two chunks ahead, completing on an mbarrier, and
//     -- one chunk behind, once verified -- two cp.async.bulk stores that write the m / least spans back.
//     No per-cell staging or commit instruction exists anywhere.
//
// Dependent chain per row on the compute warps: LDS parents -> min/select -> FADD -> keep test -> STS -> named
// barrier (416 threads).
#pragma once
#include <type_traits>

#include "carver_kernels.cuh"
#include "mmap_update_fast.cuh"

namespace b200c {

#define UT_NCW 12
#define UT_CT (UT_NCW * 32)
#define UT_THREADS (UT_CT + 64)
#define UT_TILE 11904 // words per chunk tile (4 tiles)
#define UT_RW 2048
#define UT_RWM (UT_RW - 1)
#define UT_MAXROWS 8
#define UT_MAXCW 1024
#define UT_NKS 40 // ballot words per row (36 used: 3 slots x 12 warps)

#define UT_NKW (2 * UT_MAXROWS * UT_NKS) // ballot words [chunk parity][row][UT_NKS]
#define UT_RIW (2 * UT_MAXROWS * 4)  // row info     [chunk parity][row]{guard base, slots}
#define UT_RTW (4 * UT_MAXROWS * 8)  // row tables   [tile][row]{xadd, eadd, madd, ladd, zlo, zhi, -, -}
#define UT_CTW (4 * (UT_MAXROWS + 1) * 8) // control-table slices [tile][row][side]{n[y-1], ext2, ext3, -}
static constexpr size_t ut_smem_bytes()
{
    return sizeof(int) * ((size_t) 4 * UT_TILE + 4 * UT_RW + UT_NKW + UT_RIW + UT_RTW + UT_CTW + 16 + 16 + 8 + 8 + 8 + 128);
}

__device__ __forceinline__ void ut_bar_rows() { asm volatile("bar.sync 1, %0;" ::"n"(UT_CT + 32) : "memory"); }
// control -> DMA hand-shake, once per chunk: "the previous chunk is verified" (arrive: control, sync: DMA warp)
__device__ __forceinline__ void ut_bar_commit_arrive() { asm volatile("bar.arrive 3, 64;" ::: "memory"); }
__device__ __forceinline__ void ut_bar_commit_wait() { asm volatile("bar.sync 3, 64;" ::: "memory"); }
__device__ __forceinline__ unsigned ut_saddr(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void ut_mbar_init(void *mbar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ut_saddr(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ut_mbar_expect(void *mbar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ut_saddr(mbar)), "r"(bytes) : "memory");
}
// bounded wait: a bulk copy that never completes is a bug, not a reason to hang the GPU
__device__ __forceinline__ bool ut_mbar_wait(void *mbar, unsigned parity)
{
    const unsigned a = ut_saddr(mbar);
    for (int tries = 0; tries < (1 << 22); ++tries) {
        unsigned ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}
// global -> shared bulk copy (TMA), completion counted in bytes on `mbar`
__device__ __forceinline__ void ut_bulk_load(void *dst_smem, const void *src, unsigned bytes, void *mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     ut_saddr(dst_smem)),
                 "l"(src), "r"(bytes), "r"(ut_saddr(mbar))
                 : "memory");
}
// shared -> global bulk copy (TMA), tracked by the thread's bulk async-group
__device__ __forceinline__ void ut_bulk_store(void *dst, const void *src_smem, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(ut_saddr(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void ut_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ut_bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void ut_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void ut_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Per-chunk constants of a compute thread: it owns the columns clo + tid + 384*j of the chunk window for the
// whole chunk, so every address that does not depend on the row is computed once per chunk.
struct UtSlot {
    int x;      // absolute column (clamped into the window for address purposes)
    int rb;     // ring index of x
    int rbl;    // ring index of x-1
    int rbr;    // ring index of x+1
    bool live;  // the column exists in this chunk's window
};

// One row of the compute warps.  Cells inside the row's guard range forward their value / id to the ring for the
// next row's parents; those inside the active range are recomputed first (parents from the previous row's ring).
// New value / parent are left in place in the tile; `changed` bits go to nkrow.  rt = {xadd, eadd, madd, ladd}:
// tile word offsets.  The body is branch-free (only the stores are predicated) so that every shared-memory load
// of the cell is issued up front instead of one round trip after the other.
template <int NS, bool D1>
__device__ __forceinline__ void ut_row(const DevP &p, bool row0, int4 rt, int *__restrict__ tile, int2 *ring_cur,
                                       const int2 *ring_prev, int act_lo, int act_hi, int gr_lo, int gr_hi,
                                       const UtSlot *sl, unsigned *nkrow, int lane, int warp, int clo)
{
    const int w = p.w;
    float *tilef = reinterpret_cast<float *>(tile);
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        if (NS > 1) { // a slot whose 32 columns miss the guard range has nothing to do (uniform per warp)
            const int wlo = clo + j * UT_CT + warp * 32;
            if (wlo > gr_hi || wlo + 31 < gr_lo) {
                if (lane == 0) nkrow[j * UT_NCW + warp] = 0u;
                continue;
            }
        }
        const int x = sl[j].x;
        const int z = tile[rt.x + x];
        float best;
        int parent;
        if (D1) {
            const int2 c0 = ring_prev[sl[j].rb];
            const int2 cl = ring_prev[sl[j].rbl];
            const int2 cr = ring_prev[sl[j].rbr];
            const float inf = __int_as_float(0x7f800000);
            // left-to-right scan with strict '<' == leftmost minimum; ties go right when leftright == 1.
            // (all m are finite: an out-of-image neighbour is replaced by +inf and can never win)
            const float m0 = __int_as_float(c0.x);
            const float ml = x > 0 ? __int_as_float(cl.x) : inf;
            const float mr = x < w - 1 ? __int_as_float(cr.x) : inf;
            best = fminf(fminf(ml, m0), mr);
            if (p.leftright)
                parent = mr == best ? cr.y : (m0 == best ? c0.y : cl.y);
            else
                parent = ml == best ? cl.y : (m0 == best ? c0.y : cr.y);
        } else {
            const int D = p.delta_x;
            const int dlo = max(-x, -D), dhi = min(w - 1 - x, D);
            int bdx = dlo;
            best = __int_as_float(ring_prev[(x + dlo) & UT_RWM].x);
            for (int dx = dlo + 1; dx <= dhi; ++dx) {
                const float cand = __int_as_float(ring_prev[(x + dx) & UT_RWM].x);
                if (cand < best || (cand == best && p.leftright == 1)) {
                    best = cand;
                    bdx = dx;
                }
            }
            parent = ring_prev[(x + bdx) & UT_RWM].y;
        }
        const float mo = tilef[rt.z + z];
        const float e = tilef[rt.y + z];
        const int lold = tile[rt.w + z];
        const bool in_gr = sl[j].live && x >= gr_lo && x <= gr_hi;
        const bool in_act = in_gr && x >= act_lo && x <= act_hi;
        const float new_m = __fadd_rn(e, best);
        // (double) |d| < 1e-5  <=>  |d| <= 0x3727C5AC: that float is the largest one below the double 1e-5
        const bool keep = (lold == parent) && (fabsf(__fsub_rn(mo, new_m)) <= __int_as_float(0x3727C5AC));
        const bool changed = in_act && !row0 && !keep;
        const float val = in_act ? (row0 ? e : (keep ? mo : new_m)) : mo; // row 0: m = en over the (exact) band
        if (changed || (in_act && row0)) tilef[rt.z + z] = val;
        if (changed) tile[rt.w + z] = parent;
        if (in_gr) ring_cur[sl[j].rb] = make_int2(__float_as_int(val), z);
        const unsigned word = __ballot_sync(0xffffffffu, changed);
        if (lane == 0) nkrow[j * UT_NCW + warp] = word;
    }
}

// Write the m / least spans of one verified row back to HBM: 16-byte aligned interior with a TMA bulk store,
// the (at most three) head / tail elements with plain stores.  Called by one lane per row.
__device__ __forceinline__ void ut_commit_row(const DevP &p, int y, const int *tile, const int *rtab)
{
    const int madd = rtab[2], ladd = rtab[3], zlo = rtab[4], zhi = rtab[5];
    const int a = (zlo + 3) & ~3, b = (zhi + 1) & ~3; // aligned interior [a, b)
    const float *tilef = reinterpret_cast<const float *>(tile);
    if (b > a) {
        ut_bulk_store(p.m + a, tilef + madd + a, (unsigned) (b - a) * 4u);
        if (y > 0) ut_bulk_store(p.least + a, tile + ladd + a, (unsigned) (b - a) * 4u);
        for (int z = zlo; z < a; ++z) {
            p.m[z] = tilef[madd + z];
            if (y > 0) p.least[z] = tile[ladd + z];
        }
        for (int z = b; z <= zhi; ++z) {
            p.m[z] = tilef[madd + z];
            if (y > 0) p.least[z] = tile[ladd + z];
        }
    } else {
        for (int z = zlo; z <= zhi; ++z) {
            p.m[z] = tilef[madd + z];
            if (y > 0) p.least[z] = tile[ladd + z];
        }
    }
}

template <bool D1>
__global__ void __launch_bounds__(UT_THREADS, 1) k_mmap_update_tma(DevP p)
{
    extern __shared__ __align__(128) unsigned char ut_smem[];
    int *tiles = reinterpret_cast<int *>(ut_smem);                    // [4][UT_TILE] chunk tiles
    int2 *ring = reinterpret_cast<int2 *>(tiles + 4 * UT_TILE);       // [2][RW] {m bits, id} of the previous / current row
    unsigned *nk = reinterpret_cast<unsigned *>(ring + 2 * UT_RW);    // [2][8][UT_NKS] "changed" ballot words
    int *rinfo = reinterpret_cast<int *>(nk + UT_NKW);                // [2][8][4] guard base, slots
    int *rtab = rinfo + UT_RIW;                                       // [4][8][8] row tables of the tiles
    int4 *ctile = reinterpret_cast<int4 *>(rtab + UT_RTW);           // [4][9][2] control-table slices of the tiles
    int *pub = rtab + UT_RTW + UT_CTW;                                         // [2][8] gr_lo, gr_hi, fail_row, -, act_lo, act_hi
    int *cdesc = pub + 16;                                            // [4][4] y0, rows, clo, cw
    int *clim = cdesc + 16;                                           // [2][4] x_min, x_max, y_v at chunk starts
    volatile int *misc = clim + 8;                                    // [8] 0 stop, 1 fb_row, 2 fb_xmin, 3 fb_xmax,
                                                                      //     4 last chunk entered, 5 rows verified, 6 last chunk issued
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(const_cast<int *>(misc) + 8); // [4]
    int *s_red = const_cast<int *>(misc) + 16;                        // [128] generic fallback scratch

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = p.delta_x, w = p.w, h = p.h;
    const bool is_compute = tid < UT_CT, is_control = warp == UT_NCW;

    if (tid == 0) {
        misc[0] = 0;
        misc[1] = h;
        misc[4] = -1;
        misc[5] = 0;
        misc[6] = -1;
        for (int i = 0; i < 4; ++i) ut_mbar_init(&mbar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads(); // mbarriers initialised

    if (is_compute) {
        // =============================================================================== COMPUTE
        __syncthreads(); // start 1: prologue chunks described, their loads issued
        __syncthreads(); // start 2: ranges of row 0 published
        int y = 0;
        bool failed = false;
        long long dbg_busy = 0, dbg_rel = clock64();
        int dbg_probe = 0;
        for (int k = 0;; ++k) {
            const int *dsc = cdesc + (k & 3) * 4;
            const int rows = dsc[1], clo = dsc[2], cw = dsc[3];
            if (rows == 0) break;
            if (misc[0]) { // a speculation failed during the previous chunk: stop here (same test in the control warp)
                failed = true;
                break;
            }
            if (tid == 0) misc[4] = k;
            int *tile = tiles + (k & 3) * UT_TILE;
            const int *rtc = rtab + (k & 3) * UT_MAXROWS * 8;
            unsigned *nkc = nk + (k & 1) * UT_MAXROWS * UT_NKS;
            const int ns = cw <= UT_CT ? 1 : (cw <= 2 * UT_CT ? 2 : 3);
            UtSlot sl[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int c = j * UT_CT + tid;
                sl[j].live = c < cw;
                sl[j].x = clo + min(c, max(cw - 1, 0));
                sl[j].rb = sl[j].x & UT_RWM;
                sl[j].rbl = (sl[j].x - 1) & UT_RWM;
                sl[j].rbr = (sl[j].x + 1) & UT_RWM;
            }
            if (!ut_mbar_wait(&mbar[k & 3], (unsigned) ((k >> 2) & 1))) // the chunk's bulk loads have landed
                atomicOr(p.err, 4);
            // Row time = the active warps' dependent chain, so the row loop is specialised on the slot count outside
            // the loop, carries no failure check (a failed speculation is noticed at the next chunk start; nothing
            // past the failed row is ever committed) and idle warps -- none of whose columns touch the row's guard
            // range -- read one record, clear their ballot words and go straight to the barrier.
            const int wfirst = clo + warp * 32; // first column of this warp's slot 0
            auto rows_loop = [&](auto ns_tag) {
                constexpr int NSC = decltype(ns_tag)::value;
                unsigned *nkrow = nkc;
                const int *rtr = rtc;
                for (int r = 0; r < rows; ++r, ++y, nkrow += UT_NKS, rtr += 8) {
                    const int par = y & 1;
                    const int4 pg = *reinterpret_cast<const int4 *>(pub + par * 8);     // gr_lo, gr_hi
                    const int2 pa = *reinterpret_cast<const int2 *>(pub + par * 8 + 4); // act_lo, act_hi
                    const int4 rt = *reinterpret_cast<const int4 *>(rtr);
                    bool active = false;
#pragma unroll
                    for (int j = 0; j < NSC; ++j) active |= !(wfirst + j * UT_CT > pg.y || wfirst + j * UT_CT + 31 < pg.x);
                    if (active)
                        ut_row<NSC, D1>(p, y == 0, rt, tile, ring + par * UT_RW, ring + (par ^ 1) * UT_RW, pa.x, pa.y, pg.x,
                                        pg.y, sl, nkrow, lane, warp, clo);
                    else if (lane < NSC)
                        nkrow[lane * UT_NCW + warp] = 0u;
                    if (p.dbg) dbg_busy += clock64() - dbg_rel;
                    ut_bar_rows();
                    if (p.dbg) {
                        dbg_probe += *reinterpret_cast<volatile int *>(pub); // the deferred barrier wait lands here
                        dbg_rel = clock64();
                    }
                }
            };
            if (cw <= 0) {
                for (int r = 0; r < rows; ++r, ++y) {
                    if (lane == 0) nkc[r * UT_NKS + warp] = 0u;
                    ut_bar_rows();
                }
            } else if (ns == 1)
                rows_loop(std::integral_constant<int, 1>{});
            else if (ns == 2)
                rows_loop(std::integral_constant<int, 2>{});
            else
                rows_loop(std::integral_constant<int, 3>{});
            ut_fence_async(); // tile writes (generic proxy) -> visible to the TMA stores (async proxy)
            __syncthreads(); // chunk end
        }
        if (!failed) {
            // drain: the last two rows still wait for verification
            for (int d = 0; d < 2; ++d, ++y) ut_bar_rows();
        }
        if (p.dbg && lane == 0) atomicAdd((unsigned long long *) &p.dbg[warp], (unsigned long long) dbg_busy + (dbg_probe == 0x7fffffff));
    } else if (is_control) {
        // =============================================================================== CONTROL
        // Lane-parallel: even lanes carry the LOW side (x_min, energy-band minima), odd lanes the HIGH side with
        // every quantity negated (-x_max, -maxima), so that one instruction stream computes both ends of every
        // range: "lower limit" = max(floor, min(...) - growth) on either side.
        const bool hi = lane & 1;
        const int sgn = hi ? -1 : 1;
        const int floor_s = hi ? -(w - 1) : 0;
        int xm = hi ? -min(p.nrg_xmax[0], w - 1) : max(p.nrg_xmin[0], 0); // x_min | -x_max
        int fail_row = INT_MAX;
        unsigned long long cells = 0;
        if (lane < 2) clim[lane] = sgn * xm;
        if (lane == 0) clim[2] = 0;
        __syncthreads(); // start 1: the DMA warp described chunks 0..2
        if (!ut_mbar_wait(&mbar[0], 0u)) atomicOr(p.err, 4);
        int4 clast = ctile[hi ? 1 : 0]; // this lane's side: {n[y-1], ext(n[y..y+1]), ext(n[y..y+2])} of row 0
        {
            const int4 cn = clast;
            // row 0: the active range is the exact band (m = en there); guard from rows (0, 0, 1)
            const int wb = hi ? -(cdesc[2] + cdesc[3] - 1) : cdesc[2];
            const int g = max(max(floor_s, min(xm, cn.y) - 4 * D), wb);
            const int a = max(xm, g);
            if (lane < 2) {
                pub[lane] = sgn * g;
                pub[4 + lane] = sgn * a;
            }
            if (lane == 0) pub[2] = INT_MAX;
        }
        __syncthreads(); // start 2
        int y = 0;
        bool stop = false;
        // Runs while the compute warps process row y: verify row y-1 (ballot words nkv, window start clo_v, nwords_v
        // words), publish the ranges of row y+1 (wb: this side's window bound of the chunk holding row y+1).
        auto iteration = [&](const unsigned *nkv, int clo_v, int nwords_v, int y_lim, int wb, int4 c) {
            const int a0 = y < h ? c.x : c.y; // energy-band limit of row y-1 (past the last row: the last record's n[h-1])
            const int yv = y - 1;
            if (yv >= 1 && yv < y_lim && fail_row == INT_MAX) {
                const unsigned wv = lane < nwords_v ? nkv[lane] : 0u;
                const int b = max(min(xm, a0) - D, floor_s); // bmin | -bmax of row yv
                const int bo = __shfl_xor_sync(0xffffffffu, b, 1);
                const unsigned any = __ballot_sync(0xffffffffu, wv != 0u);
                const int sel = hi ? 31 - __clz(any) : __ffs(any) - 1;
                const unsigned ws = __shfl_sync(0xffffffffu, wv, sel & 31);
                const int pos = hi ? 31 - __clz(ws) : __ffs(ws) - 1;
                const int v = sgn * (clo_v + 32 * sel + pos); // first changed column seen from this side (valid if any)
                const bool nonempty = b + bo <= 0;
                if (lane == 0 && nonempty) cells += (unsigned long long) (1 - b - bo);
                // low side: x_min = F, or bmax+1 when nothing changed; high side: x_max = (L == bmax ? bmax : L+1), or
                // bmin when nothing changed -- in negated coordinates; an empty band keeps its limits
                const int xm_any = hi ? (v == b ? b : v - 1) : v;
                const int xm_none = hi ? -bo : 1 - bo;
                const bool viol = any && (!nonempty || v < b);
                const int old = xm;
                xm = nonempty ? (any ? xm_any : xm_none) : b;
                if (__ballot_sync(0xffffffffu, viol)) {
                    fail_row = yv;
                    if (lane < 2) misc[2 + lane] = sgn * old;
                    if (lane == 0) {
                        misc[1] = yv;
                        __threadfence_block();
                        misc[0] = 1;
                    }
                }
            }
            // active range of row y+1 from the limits after row y-1 (two rows of growth), guard range (four rows),
            // both clamped to the staged window of that row's chunk; empty ([1, -1]) once a speculation has failed
            int g = max(max(floor_s, min(xm, c.z) - 4 * D), wb);
            int a = max(max(floor_s, min(xm, c.y) - 2 * D), g);
            if (fail_row != INT_MAX) g = a = 1;
            int *pb = pub + ((y + 1) & 1) * 8;
            if (lane < 2) {
                pb[lane] = sgn * g;
                pb[4 + lane] = sgn * a;
            }
        };
        // The loop mirrors the compute warps' exactly (same stop predicate at the top of every row), so both
        // roles execute the same sequence of row and chunk barriers.
        const unsigned *nk_last = nk;
        int clo_prev = 0, nw_prev = 0, klast_c = 0, rows_last = 0;
        long long cdbg_busy = 0, cdbg_rel = clock64(), cdbg_t0 = clock64();
        int cdbg_probe = 0;
        for (int k = 0;; ++k) {
            const int *dsc = cdesc + (k & 3) * 4;
            const int rows = dsc[1], clo = dsc[2], cw = dsc[3];
            if (rows == 0) break;
            const int *dn = cdesc + ((k + 1) & 3) * 4; // next chunk (described at least one chunk ago)
            const int wb_cur = hi ? -(clo + cw - 1) : clo, wb_nxt = hi ? -(dn[2] + dn[3] - 1) : dn[2];
            const int nw = (cw + 31) >> 5;
            const unsigned *nkc = nk + (k & 1) * UT_MAXROWS * UT_NKS;
            if (fail_row != INT_MAX) { // stop at chunk granularity (the compute warps test misc[0] at the same place)
                stop = true;
                if (k > 0) ut_bar_commit_arrive(); // the DMA warp waits for this chunk's hand-shake
                break;
            }
            if (!ut_mbar_wait(&mbar[k & 3], (unsigned) ((k >> 2) & 1))) atomicOr(p.err, 4); // table slice landed
            const int4 *ct = ctile + (k & 3) * (UT_MAXROWS + 1) * 2 + (hi ? 1 : 0);
            for (int r = 0; r < rows; ++r, ++y) {
                const int wb = r + 1 >= rows ? wb_nxt : wb_cur; // row y+1 may open the next chunk
                const int4 c = ct[2 * r];
                clast = c;
                if (r == 0) { // row y-1 is the last row of the previous chunk
                    iteration(nk_last, clo_prev, nw_prev, INT_MAX, wb, c);
                    if (k > 0) ut_bar_commit_arrive(); // chunk k-1 is verified to its last row: the DMA warp may commit it
                } else {
                    iteration(nkc + (r - 1) * UT_NKS, clo, nw, INT_MAX, wb, c);
                }
                if (r == rows - 1) {
                    // limits the DMA warp plans chunk k+3 from (it reads them after the chunk barrier)
                    int *cl = clim + ((k + 1) & 1) * 4;
                    if (lane < 2) cl[lane] = sgn * xm;
                    if (lane == 0) cl[2] = max(y - 1, 0);
                }
                if (p.dbg) cdbg_busy += clock64() - cdbg_rel;
                ut_bar_rows();
                if (p.dbg) {
                    cdbg_probe += *reinterpret_cast<volatile int *>(pub);
                    cdbg_rel = clock64();
                }
            }
            nk_last = nkc + (rows - 1) * UT_NKS;
            clo_prev = clo;
            nw_prev = nw;
            klast_c = k;
            rows_last = rows;
            __syncthreads(); // chunk end
        }
        if (!stop) {
            const int y_end = y;
            for (int d = 0; d < 2; ++d, ++y) {
                // d == 0 verifies row y_end-1: its n[] is the record of row y_end's .x, or the last record's .y at the image end
                iteration(nk_last, clo_prev, nw_prev, y_end, 0, y < h ? ctile[((klast_c & 3) * (UT_MAXROWS + 1) + rows_last) * 2 + (hi ? 1 : 0)] : clast);
                ut_bar_rows();
            }
            if (fail_row == INT_MAX && y_end < h) {
                // capacity stop: rows below y_end are exact; hand the exact limits to the generic loop
                if (lane < 2) misc[2 + lane] = sgn * xm;
                if (lane == 0) misc[1] = y_end;
            }
        }
        if (lane == 0 && p.cells) atomicAdd(p.cells, cells);
        if (p.dbg && lane == 0) {
            atomicAdd((unsigned long long *) &p.dbg[12], (unsigned long long) cdbg_busy + (cdbg_probe == 0x7fffffff));
            atomicAdd((unsigned long long *) &p.dbg[15], (unsigned long long) (clock64() - cdbg_t0));
        }
    } else {
        // =============================================================================== DMA warp
        // Lane r owns row r of the chunk being handled (chunks have at most 8 rows).
        // plan: rows / window of the chunk starting at row ya, from the limits (xv_min, xv_max) valid after row yv
        auto plan = [&](int ya, int xv_min, int xv_max, int yv, int &rows, int &lo, int &cw) {
            rows = 0, lo = 0, cw = 0;
            if (ya >= h) return;
            // energy-band limits of rows yv+1 .. yv+64 (two rows per lane), loaded once
            const int j0 = yv + 1 + lane, j1 = j0 + 32;
            const int n0 = j0 < h ? p.nrg_xmin[j0] : INT_MAX, x0 = j0 < h ? p.nrg_xmax[j0] : INT_MIN;
            const int n1 = j1 < h ? p.nrg_xmin[j1] : INT_MAX, x1 = j1 < h ? p.nrg_xmax[j1] : INT_MIN;
            for (int want = UT_MAXROWS; want >= 1; --want) {
                const int yb = min(ya + want, h) - 1;
                const int last = min(yb + 2, h - 1); // extremes over rows [yv+1, yb+2]
                int nlo = min(j0 <= last ? n0 : INT_MAX, j1 <= last ? n1 : INT_MAX);
                int nhi = max(j0 <= last ? x0 : INT_MIN, j1 <= last ? x1 : INT_MIN);
                nlo = __reduce_min_sync(0xffffffffu, nlo);
                nhi = __reduce_max_sync(0xffffffffu, nhi);
                // a band grows by at most delta_x per row beyond the energy bands; the guard range of the last
                // row reaches 4 rows of growth past the limits verified two rows before it
                const int dist = (yb + 3 - yv) * D;
                lo = max(0, min(xv_min, nlo) - dist);
                const int hi = min(w - 1, max(xv_max, nhi) + dist);
                cw = max(hi - lo + 1, 0);
                rows = yb - ya + 1;
                // capacity: ids (cw+6) + three spans (each about cw + removed pixels inside, checked at issue time)
                if (rows * (4 * cw + 64) <= UT_TILE && cw <= UT_MAXCW && yb + 1 - yv <= 63) return;
                rows = 0;
            }
        };
        // ids of the window ends of row ya + lane: the physical span [zlo, zhi] its en / m / least values occupy
        auto span_ends = [&](int ya, int rows, int lo, int cw, int &zlo, int &zhi) {
            zlo = 0, zhi = -1;
            if (lane < rows && cw > 0) {
                const int *rr = p.raw + (size_t) (ya + lane) * p.raw_stride;
                zlo = rr[lo];
                zhi = rr[lo + cw - 1];
            }
        };
        // lay the chunk out in its tile, write the row tables and issue the bulk loads.  Returns the number of
        // rows that fit (the spans are only known now); the chunk is cut there.
        int issued_upto = -1; // last chunk whose bulk loads were issued
        auto issue = [&](int kk, int ya, int rows, int lo, int cw, int zlo, int zhi) -> int {
            int *tile = tiles + (kk & 3) * UT_TILE;
            int *rtk = rtab + (kk & 3) * UT_MAXROWS * 8;
            const int y = ya + lane;
            const long long rawpos = (long long) y * p.raw_stride + lo;
            const int rawbase = (int) (rawpos & ~3LL), nraw = (int) ((rawpos + cw - rawbase + 3) & ~3LL);
            const int zbase = zlo & ~3, nsp = (zhi - zbase + 1 + 3) & ~3;
            const int need = lane < rows ? nraw + 3 * nsp : 0;
            // exclusive prefix of the per-row sizes -> word offset of each row in the tile
            int off = need;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, off, s);
                if (lane >= s) off += t;
            }
            const int fits = __popc(__ballot_sync(0xffffffffu, lane < rows && off <= UT_TILE));
            const int nrows = fits; // rows are laid out in order, so the ones that fit form a prefix
            off -= need;
            unsigned bytes = 0;
            if (lane < nrows) {
                rtk[lane * 8 + 0] = off + (int) ((long long) y * p.raw_stride - rawbase); // id of column x: tile[xadd + x]
                rtk[lane * 8 + 1] = off + nraw - zbase;                                    // e:     tile[eadd + z]
                rtk[lane * 8 + 2] = off + nraw + nsp - zbase;                              // m:     tile[madd + z]
                rtk[lane * 8 + 3] = off + nraw + 2 * nsp - zbase;                          // least: tile[ladd + z]
                rtk[lane * 8 + 4] = zlo;
                rtk[lane * 8 + 5] = zhi;
                bytes = (unsigned) (nraw + 3 * nsp) * 4u;
            }
            unsigned total = bytes;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) total += __shfl_xor_sync(0xffffffffu, total, s);
            // the control warp's table: records of rows ya .. ya+nrows (one extra: the row after the chunk)
            const int nrec = min(nrows + 1, h - ya);
            if (lane == 0 && nrows > 0) {
                ut_mbar_expect(&mbar[kk & 3], total + (unsigned) nrec * 32u);
                ut_bulk_load(ctile + (kk & 3) * (UT_MAXROWS + 1) * 2, p.ctab + 2 * ya, (unsigned) nrec * 32u, &mbar[kk & 3]);
            }
            __syncwarp();
            if (lane < nrows) {
                if (nraw > 0) ut_bulk_load(tile + off, p.raw + rawbase, (unsigned) nraw * 4u, &mbar[kk & 3]);
                if (nsp > 0) {
                    ut_bulk_load(tile + off + nraw, p.en + zbase, (unsigned) nsp * 4u, &mbar[kk & 3]);
                    ut_bulk_load(tile + off + nraw + nsp, p.m + zbase, (unsigned) nsp * 4u, &mbar[kk & 3]);
                    ut_bulk_load(tile + off + nraw + 2 * nsp, p.least + zbase, (unsigned) nsp * 4u, &mbar[kk & 3]);
                }
            }
            if (nrows > 0) issued_upto = kk;
            return nrows;
        };
        auto put_desc = [&](int slot, int ya, int rows, int lo, int cw) {
            if (lane == 0) {
                int *d = cdesc + slot * 4;
                d[0] = ya;
                d[1] = rows;
                d[2] = lo;
                d[3] = cw;
            }
        };
        const int b0min = max(p.nrg_xmin[0], 0), b0max = min(p.nrg_xmax[0], w - 1);
        // c1 = chunk k+1 (loads in flight), c2 = chunk k+2 (planned, span ends loaded, loads not yet issued)
        int ya1 = 0, rows1 = 0, ya2, rows2, lo2, cw2, zlo2, zhi2;
        int rows_k;
        {
            // ---- prologue: chunks 0 and 1 loading, chunk 2 planned
            int rows0, lo0, cw0, z0lo, z0hi;
            plan(0, b0min, b0max, 0, rows0, lo0, cw0);
            span_ends(0, rows0, lo0, cw0, z0lo, z0hi);
            rows0 = issue(0, 0, rows0, lo0, cw0, z0lo, z0hi);
            put_desc(0, 0, rows0, lo0, cw0);
            int lo1 = 0, cw1 = 0, z1lo, z1hi;
            ya1 = rows0;
            if (rows0 > 0) plan(ya1, b0min, b0max, 0, rows1, lo1, cw1);
            span_ends(ya1, rows1, lo1, cw1, z1lo, z1hi);
            rows1 = issue(1, ya1, rows1, lo1, cw1, z1lo, z1hi);
            put_desc(1, ya1, rows1, lo1, cw1);
            ya2 = ya1 + rows1, rows2 = 0, lo2 = 0, cw2 = 0;
            if (rows1 > 0) plan(ya2, b0min, b0max, 0, rows2, lo2, cw2);
            span_ends(ya2, rows2, lo2, cw2, zlo2, zhi2);
            rows_k = rows0; // chunk 2 is issued (and described) in iteration 0, before chunk 0 ends
        }
        __syncthreads(); // start 1
        __syncthreads(); // start 2
        // ---- steady state.  During chunk k: commit chunk k-1 (verified), issue the loads of chunk k+2, plan chunk
        // k+3 and fetch its span ends.  Everything waited for was issued a whole chunk earlier.
        for (int k = 0;; ++k) {
            if (rows_k == 0) {
                __syncthreads(); // matches the exit barrier of the other roles
                break;
            }
            // chunk k+3 first: plan it from the limits published at the end of chunk k-1 (or the initial band) and
            // fetch its span ends.  These loads are consumed in the next iteration, a whole chunk from now.
            const int *cl = clim + (k & 1) * 4;
            int ya3 = ya2 + rows2;
            int rows3 = 0, lo3 = 0, cw3 = 0, zlo3, zhi3;
            if (rows2 > 0) plan(ya3, cl[0], cl[1], cl[2], rows3, lo3, cw3);
            span_ends(ya3, rows3, lo3, cw3, zlo3, zhi3);
            if (k > 0) {
                // chunk k-1: verified to its last row during row 0 of chunk k -> write its m / least spans back
                const int *dp = cdesc + ((k - 1) & 3) * 4;
                const int yp0 = dp[0], prow = dp[1];
                ut_bar_commit_wait();
                // rows are verified in order: on a failure everything below the failed row is good
                const int r_end = min(prow, max(misc[0] ? (int) misc[1] - yp0 : prow, 0));
                if (lane < r_end)
                    ut_commit_row(p, yp0 + lane, tiles + ((k - 1) & 3) * UT_TILE, rtab + ((k - 1) & 3) * UT_MAXROWS * 8 + lane * 8);
                ut_bulk_commit();
            }
            // the tile of chunk k+2 last held chunk k-2, whose stores were issued a chunk ago: wait for all but the
            // newest store group to have read their shared-memory source
            ut_bulk_wait_read1();
            __syncwarp();
            // chunk k+2: its span ends were fetched during chunk k-1
            const int nrows2 = issue(k + 2, ya2, rows2, lo2, cw2, zlo2, zhi2);
            if (nrows2 != rows2) {
                // cut by the tile capacity (rare): the next chunk starts right after the cut, plan it again
                rows2 = nrows2;
                ya3 = ya2 + rows2;
                rows3 = 0, lo3 = 0, cw3 = 0;
                if (rows2 > 0) plan(ya3, cl[0], cl[1], cl[2], rows3, lo3, cw3);
                span_ends(ya3, rows3, lo3, cw3, zlo3, zhi3);
            }
            put_desc((k + 2) & 3, ya2, rows2, lo2, cw2);
            __syncthreads(); // chunk k end (or the exit barrier of a failed speculation)
            if (misc[0]) break;
            rows_k = rows1;
            ya1 = ya2, rows1 = rows2;
            ya2 = ya3, rows2 = rows3, lo2 = lo3, cw2 = cw3, zlo2 = zlo3, zhi2 = zhi3;
        }
        (void) ya1;
        // bulk loads the compute warps will never wait for must still land before the CTA's shared memory goes away
        if (lane == 0) misc[6] = issued_upto;
    }

    // =================================================================================== exit / fallback
    if (tid < UT_CT + 32) __syncthreads(); // compute + control: exit barrier (the DMA warp already passed its own)
    // commit what the DMA warp has not: the verified rows of the last chunk the compute warps entered
    const int fb_row = misc[1], klast = misc[4];
    if (warp == UT_NCW + 1) {
        if (klast >= 0) {
            const int *dl = cdesc + (klast & 3) * 4;
            const int r_end = min(dl[1], max(min(fb_row, h) - dl[0], 0));
            if (lane < r_end)
                ut_commit_row(p, dl[0] + lane, tiles + (klast & 3) * UT_TILE, rtab + (klast & 3) * UT_MAXROWS * 8 + lane * 8);
        }
        ut_bulk_commit();
        ut_bulk_wait_all(); // results are in HBM before the kernel ends / the generic loop reads them
        for (int kk = klast + 1; kk <= (int) misc[6]; ++kk) // chunks staged ahead that nobody consumed
            ut_mbar_wait(&mbar[kk & 3], (unsigned) ((kk >> 2) & 1));
    }
    if (fb_row < h) {
        __syncthreads();
        update_rows_generic(p, fb_row, misc[2], misc[3], s_red);
    }
}

} // namespace b200c
