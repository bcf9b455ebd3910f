// band_dp.cuh -- K2b, the incremental m-map DP after one carve (liblqr lqr_carver_update_mmap)
// on the compact maps: one CTA, trapezoid-tiled over warps, rows staged as 2-D TMA tiles.
//
// The update is a row-serial chain (row y needs row y-1), so the design minimises the latency of one row:
//
//   * The rows are cut into CHUNKS of 8 rows.  A chunk works on one column WINDOW starting at llo (16-cell aligned)
//     that contains every cell liblqr's self-trimming band could touch in those rows: the hull of the cells that
//     changed in the last finished row, grown by delta_x per row, merged with the energy bands of the rows in
//     between.  Every cell of the window's evaluation range [elo, ehi] is evaluated on every row of the chunk with
//     liblqr's keep-old rule.  Evaluating MORE cells than liblqr's band is exact: a cell whose parents, energy
//     and parent candidates did not change re-evaluates to "keep" (DESIGN.md section 5).
//   * COMPUTE warps own 128-column segments of the window, 4 consecutive cells per lane, all state in registers.
//     Inside a chunk a warp never talks to another warp: a row step is two shuffles (the neighbours' edge cells)
//     plus 4 cells of 3-input min / add / keep-test per lane.  Neighbouring segments overlap by 2*8*delta_x
//     columns; the cells a warp computes in that overlap go stale by delta_x columns per row (their parents belong
//     to the neighbour) and are simply not stored -- each column of the window is stored by exactly one warp, the
//     one in whose interior it lies.  Only at a chunk boundary do the warps meet: the last row is handed over
//     through shared memory and one named barrier.  The first three segments run on three different SM
//     sub-partitions (warps 1,2,3); sub-partition 0 belongs to the producer.
//   * One PRODUCER warp plans the windows of up to 3 chunks ahead (the band cannot outrun delta_x per row, so the
//     hull known now bounds it) and fetches them with the TMA as 2-D tiles of 128 columns x 8 rows (9 for m: the
//     row above the chunk) -- cp.async.bulk.tensor.2d from tensor maps over the compact m / en / pdx [/ rig]
//     arrays, completion on one mbarrier per chunk -- into a ring of box slots in shared memory.  Results go
//     straight back to HBM with vector stores (fire and forget; nothing waits for them).
//
// If a window ever outgrows 12 segments the kernel finishes the remaining rows with the exact generic row loop.
#pragma once
#include <cuda.h>

#include "carver_kernels.cuh"

#ifndef BD_UNROLL
#define BD_UNROLL 4 // unroll factor of the band DP's row loop
#endif

namespace b200c {

constexpr int bd_unroll = BD_UNROLL;
#define BD_NCW 12                 // compute warps = max segments of a window
#define BD_NWARPS 13              // warp 0 producer, warp s+1 = segment s: the first four segments sit on four different sub-partitions
#define BD_THREADS (BD_NWARPS * 32)
#define BD_SYNC_THREADS ((BD_NCW + 1) * 32)
#define BD_KMAX 16                // rows per chunk = height of a TMA box: 16 for delta_x <= 1 without rigidity, else 8
#define BD_BW 128                 // columns per TMA box
// cycle counters of the roles (B200C_DBG=1 at run time) are compiled in only with -DBD_PROFILE: they cost the
// compute warps ~100 cycles per chunk
#ifdef BD_PROFILE
#define BD_PROF true
#else
#define BD_PROF false
#endif
#define BD_LA 3                   // chunks planned ahead, at most
#define BD_NRING 8                // descriptor / mbarrier ring
#define BD_HMAX 4608              // rows whose energy bands fit the shared-memory table
#define BD_HANDW (BD_NCW * 128)
#define BD_RING_BYTES 170496      // 9 slots of 16 rows, 17 of 8 rows, 12 of 8 rows with the rigidity-mask box

// rows per chunk: a longer chunk amortises the chunk boundary, but its stale halo (rows * delta_x) eats the segment
__host__ __device__ constexpr int bd_rows(int delta_x, bool rig) { return (delta_x <= 1 && !rig) ? 16 : 8; }

template <int D, bool RIG>
struct BdSlot {
    static constexpr int K = bd_rows(D, RIG);
    static constexpr int box_m = (K + 1) * BD_BW * 4, box_e = K * BD_BW * 4, box_p = K * BD_BW;
    static constexpr int off_e = box_m;
    static constexpr int off_g = box_m + box_e;
    static constexpr int off_p = box_m + box_e + (RIG ? box_e : 0); // old parent offsets: read by the slow path only
    static constexpr int bytes = off_p + box_p;
    static constexpr int nslot = BD_RING_BYTES / bytes;
    static constexpr int HK = (K * D + 3) & ~3; // columns a segment edge goes stale over one chunk
    static constexpr int S = 128 - 2 * HK;      // stride of the segments = width of an interior
};

struct BdDesc {
    int seq;      // the chunk this entry describes, written LAST -- when its tiles have landed ("ready"): a compute warp
                  // reads the entry's first 16 bytes, and the rest once seq is the chunk it is about to start
    int y0, rows, llo;
    int nb, elo, ehi, nseg;
    int slot0;
    int ropen;    // the fetched columns reach the +inf sentinels right of the image: the right edge does not go stale
    int hlo, hhi; // columns [hlo, hhi) are handed over to the next chunk (the union of the interiors)
};
static_assert(sizeof(BdDesc) == 48, "BdDesc is read as three 16-byte vectors");
__device__ __forceinline__ int4 bd_lds_volatile4(const void *p)
{
    int4 v;
    asm volatile("ld.volatile.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"((unsigned) __cvta_generic_to_shared(p)) : "memory");
    return v;
}

static constexpr size_t bd_smem_bytes()
{
    return (size_t) BD_RING_BYTES + (size_t) BD_HMAX * 4 + 2 * BD_HANDW * 4 + BD_NRING * sizeof(BdDesc) + 256 + 64 + 64 +
           64 * 4 + BD_NCW * 4 * 128 * 4 + 128;
}

__device__ __forceinline__ unsigned bd_saddr(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void bd_mbar_init(void *mbar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bd_saddr(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bd_mbar_expect(void *mbar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bd_saddr(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bd_mbar_arrive(void *mbar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bd_saddr(mbar)) : "memory");
}
// bounded wait: a copy that never completes is a bug, not a reason to hang the GPU.  The bound is TIME (2 s on the
// global timer), not a number of polls: a tile delayed by preemption or a crowded device must not fail a resize.
__device__ __forceinline__ bool bd_mbar_wait(void *mbar, unsigned parity)
{
    const unsigned a = bd_saddr(mbar);
    unsigned long long t0 = 0;
    for (unsigned tries = 0;; ++tries) {
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
        if ((tries & 1023u) == 1023u) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 2000000000ull) return false;
        }
    }
}
// global -> shared 2-D tile copy (TMA), completion counted in bytes on `mbar`; out-of-range elements arrive as 0
__device__ __forceinline__ void bd_tma_load_2d(void *dst_smem, const CUtensorMap *tm, int c0, int c1, void *mbar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            bd_saddr(dst_smem)),
        "l"(reinterpret_cast<unsigned long long>(tm)), "r"(c0), "r"(c1), "r"(bd_saddr(mbar))
        : "memory");
}
// shared -> global 2-D tile store (TMA), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bd_tma_store_2d(const CUtensorMap *tm, int c0, int c1, const void *src_smem)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                     reinterpret_cast<unsigned long long>(tm)),
                 "r"(c0), "r"(c1), "r"(bd_saddr(src_smem))
                 : "memory");
}
// the compute warps arrive converged (aligned form); the producer's lanes leave spin loops one by one and need not have
// reconverged (compute-sanitizer's synccheck flags the aligned form there): it takes the unaligned form
__device__ __forceinline__ void bd_bar_chunk() { asm volatile("bar.sync 1, %0;" ::"n"(BD_SYNC_THREADS) : "memory"); }
__device__ __forceinline__ void bd_bar_chunk_producer()
{
    __syncwarp();
    asm volatile("barrier.sync 1, %0;" ::"n"(BD_SYNC_THREADS) : "memory");
}

// ------------------------------------------------------------------------------------------- compute warps
// Everything a single warp issues per row is on the critical path of the whole update, so the row body carries only
// what the NEXT row needs: the new cumulative values.
//
//   nm = en + min(parents);  d = m_old - nm
//   d == 0            -> nothing to decide (nm is m_old)
//   |d| > tol         -> liblqr stores nm
//   0 < |d| <= tol    -> "near": liblqr keeps m_old if the parent is unchanged, else stores nm.  Only this case needs
//                        the arg-min and the stored parent.
// Near rows are rare but not negligible (a few per cent of the rows of a 4K seam: near-ties between two paths), so
// they are settled one row at a time, one row late (see the row loop).  The new parent offsets are not needed by the
// chain at all (the next row reads values, not parents): k_fix_parents recomputes them from the final values, in
// parallel.
template <int D, bool LR>
__device__ __forceinline__ int bd_argmin(const float (&cand)[2 * D + 1], float best)
{
    // the scan "cand < best || (cand == best && leftright)" keeps the FIRST minimum, or the LAST when leftright
    int bdx = LR ? -D : D;
#pragma unroll
    for (int j = 1; j <= 2 * D; ++j) {
        const int jj = LR ? j : 2 * D - j;
        if (cand[jj] == best) bdx = jj - D;
    }
    return bdx;
}

// candidates of cell i of a lane: row y-1 at columns x0+i-D .. x0+i+D (the rigidity term is added by the callers)
template <int D>
__device__ __forceinline__ void bd_parents(const float (&prev)[4], float leftfloor, float (&v)[4 + 2 * D])
{
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const float l = __shfl_up_sync(full, prev[4 - D + j], 1);
        v[j] = fmaxf(l, leftfloor); // columns < 0 do not exist (+inf); columns >= w hold +inf in the maps (sentinels)
        v[4 + D + j] = __shfl_down_sync(full, prev[j], 1);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) v[D + i] = prev[i];
}

#define BD_NEAR_MAX (2u * 0x3727C5ACu - 2u)

// four cells to the m-map in HBM (an explicit global-space store: the pointer comes out of an argument block that may
// itself live in HBM, which the compiler would otherwise treat as a generic address)
__device__ __forceinline__ void bd_st_global(float *dst, const float (&v)[4])
{
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}

// one row, the fast way: out = en + min(parents); returns the lane's near key (<= BD_NEAR_MAX: a cell is "near").
// near <=> 0 < |d| <= tol <=> 2 <= 2*bits(d) (sign shifted out) <= 2*bits(tol) <=> 2*bits(d) - 2 <= 2*bits(tol) - 2 as
// unsigned; the minimum over the four cells decides for the lane.
template <int D, bool RIG>
__device__ __forceinline__ unsigned bd_eval(const float (&prev)[4], const float4 ce, const float4 co, const float4 cg,
                                            const float (&rmap)[2 * D + 1], float leftfloor, float (&out)[4])
{
    float v[4 + 2 * D];
    bd_parents<D>(prev, leftfloor, v);
    const float en[4] = {ce.x, ce.y, ce.z, ce.w};
    const float mo[4] = {co.x, co.y, co.z, co.w};
    const float rf[4] = {cg.x, cg.y, cg.z, cg.w};
    unsigned u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float best = RIG ? __fadd_rn(v[i], __fmul_rn(rf[i], rmap[0])) : v[i];
#pragma unroll
        for (int j = 1; j <= 2 * D; ++j)
            best = fminf(best, RIG ? __fadd_rn(v[i + j], __fmul_rn(rf[i], rmap[j])) : v[i + j]);
        out[i] = __fadd_rn(en[i], best);
        u[i] = (unsigned) __float_as_int(__fsub_rn(mo[i], out[i])) * 2u - 2u;
    }
    return min(min(u[0], u[1]), min(u[2], u[3]));
}

// The slow path of the row loop, kept OUT OF LINE so that the fast path stays a few dozen instructions: row q (the
// pending one) is settled with liblqr's full rule from its parents `mq` -- a near cell keeps its old value when its
// parent is unchanged -- stored, and row q+1 is evaluated again from the settled values.
struct BdSlow {
    float mp[4], nv[4];            // out: the pending row settled, the row after it redone
    const float *par;              // in: the pending row's parents (this lane's four cells: in the tile or the warp's value ring)
    const float *e, *o, *g;         // operands of the pending row in the tile (the next row's are BD_BW floats further)
    const unsigned char *pold;      // old parent offsets of the pending row (in the tile)
    float *dst;                     // where the pending row is stored (tile or value ring)
    float leftfloor;
    const float *rigmap;
    unsigned key;                   // out: near key of the row after
    int has_next;
};

template <int D, bool RIG, bool LR>
__device__ __forceinline__ void bd_slow_row(BdSlow &c)
{
    const float inf = __int_as_float(0x7f800000);
    const float tol = __int_as_float(0x3727C5AC); // (double) |d| < 1e-5  <=>  |d| <= this float
    float rmap[2 * D + 1];
#pragma unroll
    for (int j = 0; j <= 2 * D; ++j) rmap[j] = RIG ? c.rigmap[j - D] : 0.f;
    const unsigned pwo = *reinterpret_cast<const unsigned *>(c.pold);
    const float4 ce = *reinterpret_cast<const float4 *>(c.e), co = *reinterpret_cast<const float4 *>(c.o);
    // the next row's operands, fetched now: the redo below does not wait for them (rows past the chunk: the next box)
    const float4 ne = *reinterpret_cast<const float4 *>(c.e + BD_BW), no = *reinterpret_cast<const float4 *>(c.o + BD_BW);
    const float4 ng = RIG ? *reinterpret_cast<const float4 *>(c.g + BD_BW) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 cg = RIG ? *reinterpret_cast<const float4 *>(c.g) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 pv = *reinterpret_cast<const float4 *>(c.par);
    const float prev[4] = {pv.x, pv.y, pv.z, pv.w};
    float v[4 + 2 * D];
    bd_parents<D>(prev, c.leftfloor, v);
    const float en[4] = {ce.x, ce.y, ce.z, ce.w};
    const float mo[4] = {co.x, co.y, co.z, co.w};
    const float rf[4] = {cg.x, cg.y, cg.z, cg.w};
    float out[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float cand[2 * D + 1];
        float best = inf;
#pragma unroll
        for (int j = 0; j <= 2 * D; ++j) {
            cand[j] = RIG ? __fadd_rn(v[i + j], __fmul_rn(rf[i], rmap[j])) : v[i + j];
            best = fminf(best, cand[j]);
        }
        out[i] = __fadd_rn(en[i], best);
        const float df = __fsub_rn(mo[i], out[i]);
        if (df != 0.f && fabsf(df) <= tol && (int) (signed char) (pwo >> (8 * i)) == bd_argmin<D, LR>(cand, best))
            out[i] = mo[i];
        c.mp[i] = out[i];
    }
    *reinterpret_cast<float4 *>(c.dst) = make_float4(out[0], out[1], out[2], out[3]);
    if (c.has_next) {
        float nv[4];
        c.key = bd_eval<D, RIG>(out, ne, no, ng, rmap, c.leftfloor, nv);
#pragma unroll
        for (int i = 0; i < 4; ++i) c.nv[i] = nv[i];
    }
}

// Chunk protocol (one named barrier per chunk, nothing else on the critical path between two chunks): the producer
// plans chunk k+1 and WAITS for its tiles before it arrives at the barrier that ends chunk k, so a compute warp that
// leaves that barrier finds descriptor and tiles of chunk k+1 in place -- it never polls an mbarrier itself, it reads
// the producer's "ready" word (and spins on it only when the window is so wide that chunk k+1 did not fit into the
// ring beside chunk k and is fetched after the barrier).
template <int D, bool RIG, bool LR>
__device__ __forceinline__ void bd_compute(const DevP &p, unsigned char *ring, float *hand, float *vring,
                                           const BdDesc *desc, int *hull, const volatile int *ready, int seg, int lane)
{
    using SL = BdSlot<D, RIG>;
    constexpr int HK = SL::HK, S = SL::S;
    const float inf = __int_as_float(0x7f800000);
    const unsigned full = 0xffffffffu;
    float rmap[2 * D + 1];
#pragma unroll
    for (int j = 0; j <= 2 * D; ++j) rmap[j] = RIG ? p.rigmap[j - D] : 0.f;
    float mp[4] = {0.f, 0.f, 0.f, 0.f}; // row y-1 at this lane's cells
    int hl_lo = 0, hl_hi = 0;           // hand-over range of the previous chunk
    long long t_bar = 0, t_rows = 0, t_pro = 0, t_epi = 0, t_all = (BD_PROF && p.dbg) ? clock64() : 0;
    int n_rows = 0, n_slow = 0;
    auto ld4 = [](const float *q) { return *reinterpret_cast<const float4 *>(q); };
    const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f);

    bd_bar_chunk(); // chunk 0 is planned and has landed
    long long t_start = 0, t_gap = 0, t_prev = 0, t_pro1 = 0, t_epi1 = 0;
    if (BD_PROF && p.dbg) t_prev = clock64(), t_start = t_prev - t_all;
    for (int k = 0;; ++k) {
        long long tp0 = 0;
        if (BD_PROF && p.dbg) tp0 = clock64(), t_gap += tp0 - t_prev;
        BdDesc d;
        {
            const BdDesc *dq = desc + (k % BD_NRING);
            int4 v0 = bd_lds_volatile4(dq);
            while (v0.x != k) v0 = bd_lds_volatile4(dq); // normally true at once: see the producer
            const int4 v1 = bd_lds_volatile4(reinterpret_cast<const int4 *>(dq) + 1);
            const int4 v2 = bd_lds_volatile4(reinterpret_cast<const int4 *>(dq) + 2);
            d.seq = v0.x, d.y0 = v0.y, d.rows = v0.z, d.llo = v0.w;
            d.nb = v1.x, d.elo = v1.y, d.ehi = v1.z, d.nseg = v1.w;
            d.slot0 = v2.x, d.ropen = v2.y, d.hlo = v2.z, d.hhi = v2.w;
        }
        if (d.rows == 0) break;
        int *hull_k = hull + (k & 1) * (2 * BD_NCW);
        if (seg < d.nseg) {
            const int rows = d.rows;
            const int lw = d.nb * BD_BW;
            const int x0 = d.llo + seg * S + 4 * lane;
            const int c = min(x0 - d.llo, lw - 4);
            // a segment edge next to another segment -- or to columns that were not fetched -- goes stale; an edge
            // at the image border (or inside the +inf sentinel columns) does not
            const int ilo = (seg == 0 && d.llo == 0) ? 0 : HK;
            const int ihi = (seg == d.nseg - 1 && d.ropen) ? 128 : 128 - HK;
            const bool interior = 4 * lane >= ilo && 4 * lane < ihi && x0 < d.hhi; // hhi: see the producer
#ifdef BD_DEBUG_STORE_ALL
            const bool st = interior && x0 < p.pitch;
#else
            const bool st = interior && x0 >= d.elo && x0 <= d.ehi;
#endif
            const float leftfloor = x0 == 0 ? inf : -inf; // +inf in the lane whose first column is column 0

            int slot = d.slot0 + (c >> 7);
            if (slot >= SL::nslot) slot -= SL::nslot;
            const unsigned char *sb = ring + (size_t) slot * SL::bytes;
            const int cc = c & (BD_BW - 1);
            float *op = const_cast<float *>(reinterpret_cast<const float *>(sb)) + cc; // m rows -1 .. K-1: old values in, new out
            const float *ep = reinterpret_cast<const float *>(sb + SL::off_e) + cc;   // en rows 0 .. K-1
            const float *gq = reinterpret_cast<const float *>(sb + SL::off_g) + cc;   // rigidity mask rows (RIG)

            if (BD_PROF && p.dbg) t_pro1 += clock64() - tp0;
            if (d.y0 > 0) {
                const float4 v = (x0 >= hl_lo && x0 < hl_hi) ? ld4(hand + ((k - 1) & 1) * BD_HANDW + (x0 - hl_lo)) : ld4(op);
                mp[0] = v.x, mp[1] = v.y, mp[2] = v.z, mp[3] = v.w;
            }
            op += BD_BW; // row 0 of the chunk
            const float4 po = ld4(op + (size_t) (rows - 1) * BD_BW); // old values of the last row
            const unsigned char *pp = sb + SL::off_p + cc;                            // pdx rows 0 .. K-1 (slow path only)
            int r = 0;
            if (d.y0 == 0) { // row 0 of the image: m = en (A.8; true of every cell of the row, evaluated or not)
                const float4 e4 = ld4(ep);
                mp[0] = e4.x, mp[1] = e4.y, mp[2] = e4.z, mp[3] = e4.w;
                if (st) *reinterpret_cast<float4 *>(op) = e4;
                r = 1;
            }
            long long tr0 = 0;
            if (BD_PROF && p.dbg) {
                tr0 = clock64();
                t_pro += tr0 - tp0;
            }
            // The row loop is SPECULATIVE by one row: row r is computed from row r-1's fast values while the question
            // "did row r-1 have a near cell?" is still open, so the vote + branch that answers it is off the chain
            // (loop-carried dependency: shuffle -> 3-input min -> add).  If it did (a few per cent of the rows), the
            // out-of-line slow path settles row r-1 with the full rule -- its operands are still in the tile, its
            // parents (row r-2) in the warp's value ring, its old parent offsets in the tile as well -- and
            // redoes row r.
            // Every row is stored ONCE, one row late -- when the vote has confirmed it -- by one 16-byte store per lane:
            // the lanes that own their columns (st) write into the tile, over the old values of the same cells (the
            // producer writes the boxes back to HBM with bulk tensor stores after the chunk barrier: no global store on
            // the chain); the other lanes (halo columns, whose values still feed their neighbours' shuffles) into a
            // per-warp ring of four rows (slot = row index relative to the loop entry, mod 4).  Either way the values
            // of the rows up to r-2 are in shared memory when the loop is left at row r, so the slow path needs no
            // register state beyond the row counter and the unrolled fast rows stay free of moves.
            float *vr = vring + 4 * lane; // [4][128] per warp
            bool pend = false;            // row r-1 has a near cell in this lane
            int r0 = r;                   // entry row of the fast loop
            unsigned key = 0xffffffffu;
            auto slow = [&](int q, bool has_next) { // settles row q (stored), redoes row q+1 (-> mp, stored once confirmed)
                BdSlow c;
                c.par = st ? op + (q - 1) * BD_BW : vr + ((q - 1 - r0) & 3) * 128;
                c.e = ep + q * BD_BW, c.o = op + q * BD_BW, c.g = gq + q * BD_BW;
                c.pold = pp + q * BD_BW;
                // over the row's old values, which the call reads first / slot 2 of the ring as re-based after the call
                c.dst = st ? op + q * BD_BW : vr + 2 * 128;
                c.leftfloor = leftfloor;
                c.rigmap = p.rigmap;
                c.has_next = has_next;
                bd_slow_row<D, RIG, LR>(c);
#pragma unroll
                for (int i = 0; i < 4; ++i) mp[i] = has_next ? c.nv[i] : c.mp[i];
                key = c.key;
                ++n_slow;
            };
            // one fast row: returns true when the row before it turns out to be pending (nothing was stored).  The
            // operands of the row (ce / co / cg) were fetched while the previous row was computed; the next row's are
            // fetched first thing here (past the last row of the chunk that reads the neighbouring box of the ring and
            // is never used), so the chain never waits for shared memory.
            float4 ce = ld4(ep + r * BD_BW), co = ld4(op + r * BD_BW), cg = RIG ? ld4(gq + r * BD_BW) : one4;
            auto fast_row = [&](const float *e_next, const float *o_next, const float *g_next, int slot_prev) -> bool {
                const float4 ne = ld4(e_next), no = ld4(o_next), ng = RIG ? ld4(g_next) : one4;
                float nv[4];
                key = bd_eval<D, RIG>(mp, ce, co, cg, rmap, leftfloor, nv);
#ifdef BD_DEBUG_ALWAYS_SLOW
                pend = e_next > ep + (r0 + 1) * BD_BW;
#endif
#ifdef BD_EXP_NO_SLOW
                pend = false; // timing experiment only: never leave the fast loop (results are wrong)
#endif
                if (__any_sync(full, pend)) return true;
                // row r-1 is confirmed: o_next is row r+1 of the tile.
                // The neighbouring segment's halo lanes read these very cells as THEIR "old value" of the row, possibly
                // after this store (compute-sanitizer's racecheck reports the pair): harmless, because the rule is
                // idempotent -- a cell that kept its old value reads back the old value; a cell that took the new
                // value nm reads back nm, finds d = 0 and takes its own, identical, nm.  Each 4-byte word is one or the
                // other, never torn.
                float *dst = st ? const_cast<float *>(o_next) - 2 * BD_BW : vr + slot_prev * 128;
                *reinterpret_cast<float4 *>(dst) = make_float4(mp[0], mp[1], mp[2], mp[3]);
#pragma unroll
                for (int i = 0; i < 4; ++i) mp[i] = nv[i];
                ce = ne, co = no, cg = ng;
                pend = key <= BD_NEAR_MAX;
                return false;
            };
            for (;;) {
                bool hit = false;
                for (; r + 4 <= rows; r += 4) { // (r - r0) is a multiple of 4 here: the slots are compile-time constants
                    const float *e0 = ep + (r + 1) * BD_BW, *o0 = op + (r + 1) * BD_BW, *g0 = gq + (r + 1) * BD_BW;
                    if (fast_row(e0, o0, g0, 3)) { hit = true; break; }
                    if (fast_row(e0 + BD_BW, o0 + BD_BW, g0 + BD_BW, 0)) { r += 1; hit = true; break; }
                    if (fast_row(e0 + 2 * BD_BW, o0 + 2 * BD_BW, g0 + 2 * BD_BW, 1)) { r += 2; hit = true; break; }
                    if (fast_row(e0 + 3 * BD_BW, o0 + 3 * BD_BW, g0 + 3 * BD_BW, 2)) { r += 3; hit = true; break; }
                }
                if (!hit)
                    for (; r < rows; ++r)
                        if (fast_row(ep + (r + 1) * BD_BW, op + (r + 1) * BD_BW, gq + (r + 1) * BD_BW, (r - 1 - r0) & 3)) { hit = true; break; }
                if (!hit) break;
                // operands of the row after the redone one: fetched before the slow path, which does not touch them
                const float4 ce_n = ld4(ep + (r + 1) * BD_BW), co_n = ld4(op + (r + 1) * BD_BW);
                const float4 cg_n = RIG ? ld4(gq + (r + 1) * BD_BW) : one4;
                slow(r - 1, true); // row r-1 settled (and stored), row r redone from it; it is stored once confirmed
                pend = key <= BD_NEAR_MAX;
                r0 = ++r;
                ce = ce_n, co = co_n, cg = cg_n;
            }
            if (__any_sync(full, pend)) slow(rows - 1, false); // the last row of the chunk is still open
            else if (st) *reinterpret_cast<float4 *>(op + (rows - 1) * BD_BW) = make_float4(mp[0], mp[1], mp[2], mp[3]);
            long long te0 = 0;
            if (BD_PROF && p.dbg) {
                te0 = clock64();
                t_rows += te0 - tr0;
                n_rows += rows;
            }
            // hand the last row over and publish the hull of the cells whose VALUE changed in it (a changed parent
            // alone does not matter to the next row); po = the old values of the last row computed
            if (interior) *reinterpret_cast<float4 *>(hand + (k & 1) * BD_HANDW + (x0 - d.hlo)) = make_float4(mp[0], mp[1], mp[2], mp[3]);
            // the hull at lane granularity (4 cells): a superset of the changed cells, which is all the planner needs
            if (BD_PROF && p.dbg) t_epi1 += clock64() - te0;
            bool chg = st;
            if (st && (rows > 1 || d.y0 > 0)) chg = mp[0] != po.x || mp[1] != po.y || mp[2] != po.z || mp[3] != po.w;
            const unsigned bal = __ballot_sync(full, chg);
            if (lane == 0) {
                const int xs = d.llo + seg * S;
                hull_k[2 * seg] = bal ? xs + 4 * (__ffs(bal) - 1) : INT_MAX;
                hull_k[2 * seg + 1] = bal ? xs + 4 * (31 - __clz(bal)) + 3 : INT_MIN;
            }
            if (BD_PROF && p.dbg) t_epi += clock64() - te0;
        } else if (lane == 0) {
            hull_k[2 * seg] = INT_MAX;
            hull_k[2 * seg + 1] = INT_MIN;
        }
        long long t1 = 0;
        if (BD_PROF && p.dbg) t1 = clock64();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the tile rows written above, for the bulk stores
        bd_bar_chunk();
        if (BD_PROF && p.dbg) t_prev = clock64(), t_bar += t_prev - t1;
        hl_lo = d.hlo;
        hl_hi = d.hhi;
    }
    if (BD_PROF && p.dbg && lane == 0) {
        atomicAdd((unsigned long long *) &p.dbg[seg == 0 ? 1 : 3], (unsigned long long) t_bar);
        if (seg == 0) {
            atomicAdd((unsigned long long *) &p.dbg[0], (unsigned long long) t_pro);
            atomicAdd((unsigned long long *) &p.dbg[2], (unsigned long long) t_epi);
            atomicAdd((unsigned long long *) &p.dbg[5], (unsigned long long) t_rows);
            atomicAdd((unsigned long long *) &p.dbg[6], (unsigned long long) (clock64() - t_all));
            atomicAdd((unsigned long long *) &p.dbg[7], (unsigned long long) n_rows);
            atomicAdd((unsigned long long *) &p.dbg[10], (unsigned long long) n_slow);
            atomicAdd((unsigned long long *) &p.dbg[13], (unsigned long long) t_start);
            atomicAdd((unsigned long long *) &p.dbg[26], (unsigned long long) t_pro1);
            atomicAdd((unsigned long long *) &p.dbg[27], (unsigned long long) t_epi1);
            atomicAdd((unsigned long long *) &p.dbg[15], (unsigned long long) t_gap);
        } else {
            atomicAdd((unsigned long long *) &p.dbg[8], (unsigned long long) n_rows);
            atomicAdd((unsigned long long *) &p.dbg[11], (unsigned long long) n_slow);
        }
    }
}

// ------------------------------------------------------------------------------------------- producer warp
struct BdMaps {
    CUtensorMap m, en, pdx, rig, mst; // mst: the m-map again with boxes of K rows, for the stores // 2-D tiled maps over the compact arrays: boxes of 128 x (K+1) (m) / 128 x K rows
};

template <int D, bool RIG>
__device__ __forceinline__ void bd_producer(const DevP &p, const BdMaps &tm, unsigned char *ring, const unsigned *nrg,
                                            BdDesc *desc, const int *hull, unsigned long long *mbar, volatile int *misc,
                                            int lane, int seam)
{
    using SL = BdSlot<D, RIG>;
    constexpr int HK = SL::HK, S = SL::S, K = SL::K;
    const unsigned full = 0xffffffffu;
    int ya = 0;                                  // first row of the next chunk to plan
    int hlo = INT_MAX, hhi = INT_MIN, yl = -1;   // hull of the changed cells of row yl
    bool ended = false;
    int kp = 0;                                  // next chunk to plan
    int slot_next = 0, slots_free = SL::nslot;
    unsigned long long cells = 0;
    long long t_plan = 0;

    // The FAR phase of this seam's carve (k_carve phase 2) may still be running beside this kernel: in row y the columns
    // from max(s[y] - delta_x - 1, 0) + B200C_CARVE_SPAN on are its business until it has counted all h rows.  A window
    // that reaches that far (a band over a thousand columns wide) waits for it; so does the end of the tiled path when
    // rows are left to the tail kernel / the wide-window loop, which read whole rows.
    bool far_done = p.far == nullptr;
    auto far_wait = [&]() {
        const int need = (seam + 1) * p.h;
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (*reinterpret_cast<volatile int *>(p.far) < need) {
            // FAR never waits for anybody, so this ends as soon as it has run; should the two kernels ever be put in
            // one hardware queue, give up after 2 s and flag the session instead of hanging the device
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 2000000000ull) {
                if (lane == 0) atomicOr(p.err, 8);
                break;
            }
        }
        __threadfence();
        far_done = true;
    };
    // The write-back of a finished chunk (bulk tensor stores out of its boxes) is only waited for when the ring runs out
    // of boxes: planning the next chunk and fetching it into free boxes goes on while the copies read.
    int store_nb = 0; // boxes of the chunk whose stores are in flight
    auto reclaim = [&]() {
        if (store_nb) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
            slots_free += store_nb;
            store_nb = 0;
        }
    };
    // plans chunk kp and issues its tiles; false when it has to wait for ring space
    auto plan_issue = [&]() -> bool {
        const long long tp0 = (BD_PROF && p.dbg) ? clock64() : 0;
        BdDesc *dd = desc + (kp % BD_NRING);
        void *mb = &mbar[kp % BD_NRING];
        int end_y = -1;
        int elo = 0, ehi = -1, llo = 0, nb = 0, nseg = 0, rows = 0, ropen = 0, lw = 0;
        if (ya >= p.h) {
            end_y = p.h;
        } else {
            rows = min(K, p.h - ya);
            const int yb = ya + rows - 1;
            int nlo = INT_MAX, nhi = INT_MIN; // extremes of the energy bands of rows (yl, yb]
            for (int j = yl + 1 + lane; j <= yb; j += 32) {
                const unsigned pk = nrg[j];
                const int a = (int) (pk & 0xffffffu);
                nlo = min(nlo, a);
                nhi = max(nhi, a + (int) (pk >> 24) - 1);
            }
            nlo = __reduce_min_sync(full, nlo);
            nhi = __reduce_max_sync(full, nhi);
            const int g = (yb - yl) * D; // the band grows by at most delta_x per row past the hull / energy bands
            elo = max(min(hlo, nlo) - g, 0) & ~3;       // whole lanes: evaluating more cells is exact, and
            ehi = min(max(hhi, nhi) + g, p.w - 1) | 3;  // cells right of w-1 are +inf sentinels nobody reads
            // the window reaches HK columns past the evaluation range on either side: its outer edges go stale like
            // the edges between segments, unless they are the image border / the sentinel columns
            const int wlim = min((p.w + 4 + 3) & ~3, p.pitch);
            llo = max(min(elo, p.w - 1) - HK, 0) & ~15;
            const int need = max(min(max(ehi, elo) + HK + 1, wlim) - llo, 16);
            nseg = need <= 128 ? 1 : (need - 2 * HK + S - 1) / S;
            nb = (need + BD_BW - 1) / BD_BW;
            lw = nb * BD_BW;
            ropen = llo + min(lw, (nseg - 1) * S + 128) >= wlim; // fetched AND covered by the last segment
            // too wide for the tiled path: the tail kernel / the exact generic loop takes over at row ya
            if (nseg > min(BD_NCW, p.bd_maxseg) || nb > SL::nslot) end_y = ya;
        }
        if (end_y < 0 && nb > slots_free) {
            reclaim(); // the boxes of the chunk that is being written back, once the copies have read them
            if (nb > slots_free) return false;
        }
        if (end_y < 0 && !far_done) {
            int cmin = INT_MAX; // leftmost energy-band start of the rows the tiles cover (incl. the row above the chunk)
            for (int j = max(ya - 1, 0) + lane; j <= ya + rows - 1; j += 32) cmin = min(cmin, (int) (nrg[j] & 0xffffffu));
            cmin = __reduce_min_sync(full, cmin);
            if (llo + nb * BD_BW > max(cmin - D - 1, 0) + B200C_CARVE_SPAN) far_wait();
        }
        if (end_y >= 0 && end_y < p.h && !far_done) far_wait();
        if (end_y >= 0) {
            if (lane == 0) {
                dd->rows = 0;
                dd->y0 = end_y;
                bd_mbar_arrive(mb);
            }
            ended = true;
        } else {
            if (lane == 0) {
                dd->y0 = ya;
                dd->rows = rows;
                dd->llo = llo;
                dd->nb = nb;
                dd->elo = elo;
                dd->ehi = ehi;
                dd->nseg = nseg;
                dd->slot0 = slot_next;
                dd->ropen = ropen;
                dd->hlo = llo + (llo == 0 ? 0 : HK);
                // the fetched columns end at llo + lw: unless that is past the image, the last HK of them go stale too
                dd->hhi = llo + (ropen ? min(lw, (nseg - 1) * S + 128) : min(lw - HK, nseg * S + HK));
                if (p.fix) p.fix[kp] = make_int4(ya, rows, elo, ehi);
                bd_mbar_expect(mb, (unsigned) nb * (unsigned) SL::bytes);
            }
            __syncwarp();
            if (lane < nb) {
                int slot = slot_next + lane;
                if (slot >= SL::nslot) slot -= SL::nslot;
                unsigned char *sb = ring + (size_t) slot * SL::bytes;
                const int cx = llo + lane * BD_BW;
                bd_tma_load_2d(sb, &tm.m, cx, ya - 1, mb);
                bd_tma_load_2d(sb + SL::off_e, &tm.en, cx, ya, mb);
                if (RIG) bd_tma_load_2d(sb + SL::off_g, &tm.rig, cx, ya, mb);
                bd_tma_load_2d(sb + SL::off_p, &tm.pdx, cx, ya, mb);
            }
            slot_next += nb;
            if (slot_next >= SL::nslot) slot_next -= SL::nslot;
            slots_free -= nb;
            if (ehi >= elo) cells += (unsigned long long) (ehi - elo + 1) * rows;
            ya += rows;
        }
        ++kp;
        if (BD_PROF && p.dbg) t_plan += clock64() - tp0;
        return true;
    };

    // misc[4] = the last chunk whose descriptor is written and whose tiles have landed ("ready"); the compute warps
    // read it once after every barrier and only wait on it when a window is so wide that the next chunk did not fit
    // into the ring beside the current one
    auto publish_ready = [&](int kk) {
        if (!bd_mbar_wait(&mbar[kk % BD_NRING], (unsigned) ((kk / BD_NRING) & 1))) atomicOr(p.err, 4);
        __threadfence_block();
        if (lane == 0) {
            *reinterpret_cast<volatile int *>(&desc[kk % BD_NRING].seq) = kk;
            misc[4] = kk;
        }
        __syncwarp();
    };
    int ready = -1;
    // chunk 0: planned, landed, then the start barrier
    while (!ended && kp <= BD_LA)
        if (!plan_issue()) break;
    publish_ready(0), ready = 0;
    bd_bar_chunk_producer();
    long long t_tiles = 0, t_pbar = 0, t_pstore = 0;
    for (int k = 0;; ++k) {
        // a chunk that did not fit before the last barrier is planned now that its predecessor's slots are free: the
        // compute warps are waiting for it
        while (!ended && kp <= k + BD_LA)
            if (!plan_issue()) break;
        if (ready < k) publish_ready(k), ready = k;
        const BdDesc *dk = desc + (k % BD_NRING);
        const int rows_k = dk->rows, y0_k = dk->y0, nb_k = dk->nb;
        if (rows_k == 0) {
            if (lane == 0) {
                misc[0] = y0_k; // first row left to the generic loop (h: none)
                misc[1] = hlo;
                misc[2] = hhi;
            }
            break;
        }
        // if chunk k+1 is planned, its tiles must have landed before anybody leaves the barrier that ends chunk k
        {
            const long long tw0 = (BD_PROF && p.dbg) ? clock64() : 0;
            if (kp >= k + 2) publish_ready(k + 1), ready = k + 1;
            if (BD_PROF && p.dbg) t_tiles += clock64() - tw0;
        }
        const long long tb0 = (BD_PROF && p.dbg) ? clock64() : 0;
        bd_bar_chunk_producer(); // end of chunk k: the hull of its last row is known, its new values are in the tiles
        if (BD_PROF && p.dbg) { // the barrier instruction does not hold back a clock read: read the clock after a load
            const int probe = *reinterpret_cast<const volatile int *>(hull);
            long long tb1 = clock64();
            if (probe == 0x7ffffff1) tb1 = 0;
            t_pbar += tb1 - tb0;
        }
        const long long ts0 = (BD_PROF && p.dbg) ? clock64() : 0;
        // write chunk k's boxes back (rows 0 .. K-1 of every m box; rows past the image are clipped by the tensor map),
        // and let the copies READ the slots before these are handed to a later chunk's loads
        reclaim(); // (the previous chunk's, if nobody needed its boxes in the meantime)
        if (lane < nb_k) {
            int slot = dk->slot0 + lane;
            if (slot >= SL::nslot) slot -= SL::nslot;
            bd_tma_store_2d(&tm.mst, dk->llo + lane * BD_BW, y0_k, ring + (size_t) slot * SL::bytes + BD_BW * 4);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
        store_nb = nb_k;
        if (BD_PROF && p.dbg) t_pstore += clock64() - ts0;
        const int *hull_k = hull + (k & 1) * (2 * BD_NCW);
        int lo = lane < BD_NCW ? hull_k[2 * lane] : INT_MAX;
        int hi = lane < BD_NCW ? hull_k[2 * lane + 1] : INT_MIN;
        hlo = __reduce_min_sync(full, lo);
        hhi = __reduce_max_sync(full, hi);
        yl = y0_k + rows_k - 1;
    }
    if (BD_PROF && p.dbg && lane == 0) {
        atomicAdd((unsigned long long *) &p.dbg[12], (unsigned long long) t_tiles);
        atomicAdd((unsigned long long *) &p.dbg[24], (unsigned long long) t_pbar);
        atomicAdd((unsigned long long *) &p.dbg[25], (unsigned long long) t_pstore);
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // every store has landed before the kernel goes on / ends
    asm volatile("fence.proxy.async;" ::: "memory");
    __threadfence();
    if (lane == 0 && p.cells) atomicAdd(p.cells, cells);
    if (lane == 0 && p.fixn) *p.fixn = kp - 1; // chunks whose parents k_fix_parents has to recompute (the last is the end marker)
    if (BD_PROF && p.dbg && lane == 0) {
        atomicAdd((unsigned long long *) &p.dbg[4], (unsigned long long) kp);
        atomicAdd((unsigned long long *) &p.dbg[9], (unsigned long long) t_plan);
    }
}

// Parents of the cells the band DP evaluated, recomputed from the final values: one CTA per chunk of the table the
// producer left in p.fix ({y0, rows, elo, ehi}); 4 cells per thread, packed store.  liblqr writes least[] for every
// cell of its band (A.8); for a kept cell the arg-min equals the stored parent, so writing the arg-min everywhere
// in the (larger) evaluated range is the same map.
template <int D, bool RIG, bool LR>
__global__ void __launch_bounds__(256) k_fix_parents(const DevP pin0, const DevP *tab)
{
    pdl_entry();
    const DevP pin = pick_image(pin0, tab);
    const DevP p = seam_view(pin, 1);
    if ((int) blockIdx.x >= *p.fixn) return;
    const int4 f = p.fix[blockIdx.x];
    const int y0 = f.x, rows = f.y, elo = f.z, ehi = f.w;
    const int nq = (ehi - elo + 1) >> 2; // groups of 4 cells per row
    const float inf = __int_as_float(0x7f800000);
    float rmap[2 * D + 1];
#pragma unroll
    for (int j = 0; j <= 2 * D; ++j) rmap[j] = RIG ? p.rigmap[j - D] : 0.f;
    // gridDim.y CTAs share a chunk: the kernel is a few dependent trips to L2 per thread, not bandwidth
    for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < rows * nq; t += blockDim.x * gridDim.y) {
        const int r = t / nq, x0 = elo + 4 * (t - r * nq), y = y0 + r;
        if (y == 0) continue;
        const float *up = p.m + (size_t) (y - 1) * p.pitch;
        float v[4 + 2 * D];
#pragma unroll
        for (int j = 0; j < 4 + 2 * D; ++j) {
            const int x = x0 - D + j;
            v[j] = (x >= 0 && x < p.pitch) ? up[x] : inf;
        }
        unsigned pk = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float rf = RIG ? p.rig[(size_t) y * p.pitch + x0 + i] : 1.f;
            float cand[2 * D + 1];
            float best = inf;
#pragma unroll
            for (int j = 0; j <= 2 * D; ++j) {
                cand[j] = RIG ? __fadd_rn(v[i + j], __fmul_rn(rf, rmap[j])) : v[i + j];
                best = fminf(best, cand[j]);
            }
            pk |= ((unsigned) (bd_argmin<D, LR>(cand, best) & 0xff)) << (8 * i);
        }
        *reinterpret_cast<unsigned *>(p.pdx + (size_t) y * p.pitch + x0) = pk;
    }
}

// The rows the tiled path cannot take (window wider than BD_NCW segments): the same superset evaluation as
// update_rows_generic, row by row with the whole CTA, but with the previous row kept in shared memory (the tile ring is
// free by now) and every operand of a row fetched in one batch of independent loads, so a row costs about one
// trip to L2 instead of one per candidate.  prev / cur: two row buffers of `cap` floats (cap >= pitch).
template <int D, bool RIG, bool LR>
__device__ void bd_rows_wide(const DevP &p, int y_from, int lo, int hi, float *prev, float *cur, int *s_red)
{
    constexpr int NB = 8; // cells per thread and batch
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x, nwarp = nthr >> 5;
    const float inf = __int_as_float(0x7f800000);
    float rmap[2 * D + 1];
#pragma unroll
    for (int j = 0; j <= 2 * D; ++j) rmap[j] = RIG ? p.rigmap[j - D] : 0.f;
    int plo = 1, phi = 0; // columns of `prev` holding row y-1 (empty: parents come from HBM)
    unsigned long long cells = 0;
    for (int y = y_from; y < p.h; ++y) {
        const size_t o = (size_t) y * p.pitch;
        const int elo = max(min(lo, p.nrg_xmin[y]) - (y ? D : 0), 0);
        const int ehi = min(max(hi, p.nrg_xmax[y]) + (y ? D : 0), p.w - 1);
        if (y > 0) { // parents of the range that row y-1 did not evaluate: unchanged, from HBM
            const int need_lo = max(elo - D, 0), need_hi = min(ehi + D, p.w - 1);
            const float *up = p.m + o - p.pitch;
            if (plo > phi) {
                for (int x = need_lo + tid; x <= need_hi; x += nthr) prev[x] = up[x];
            } else {
                for (int x = need_lo + tid; x < plo; x += nthr) prev[x] = up[x];
                for (int x = phi + 1 + tid; x <= need_hi; x += nthr) prev[x] = up[x];
            }
        }
        __syncthreads();
        int first = INT_MAX, last = INT_MIN;
        for (int base = elo; base <= ehi; base += nthr * NB) {
            float en[NB], mo[NB], rf[NB];
            int pd[NB];
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const int x = base + j * nthr + tid;
                en[j] = mo[j] = 0.f, rf[j] = 1.f, pd[j] = 0;
                if (x <= ehi) {
                    en[j] = p.en[o + x];
                    mo[j] = p.m[o + x];
                    pd[j] = p.pdx[o + x];
                    if (RIG) rf[j] = p.rig[o + x];
                }
            }
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const int x = base + j * nthr + tid;
                if (x > ehi) continue;
                float val;
                if (y == 0) { // row 0: m = en over the energy band; the band carries over to row 1 (A.8)
                    val = en[j];
                    p.m[x] = val;
                    first = min(first, x), last = max(last, x);
                } else {
                    float cand[2 * D + 1];
                    float best = inf;
#pragma unroll
                    for (int k = 0; k <= 2 * D; ++k) {
                        const int xx = x + k - D;
                        const float pv = (xx >= 0 && xx <= p.w - 1) ? prev[xx] : inf;
                        cand[k] = RIG ? __fadd_rn(pv, __fmul_rn(rf[j], rmap[k])) : pv;
                        best = fminf(best, cand[k]);
                    }
                    const int bdx = bd_argmin<D, LR>(cand, best);
                    const float nm = __fadd_rn(en[j], best);
                    val = mo[j];
                    if (!keep_old(pd[j], bdx, mo[j], nm)) {
                        val = nm;
                        p.m[o + x] = nm;
                        p.pdx[o + x] = (int8_t) bdx;
                        first = min(first, x), last = max(last, x);
                    }
                }
                cur[x] = val;
            }
        }
        if (tid == 0 && ehi >= elo) cells += (unsigned long long) (ehi - elo + 1);
        first = __reduce_min_sync(0xffffffffu, first);
        last = __reduce_max_sync(0xffffffffu, last);
        const int buf = (y & 1) * 32;
        if (lane == 0) {
            s_red[buf + warp] = first;
            s_red[buf + 16 + warp] = last;
        }
        __syncthreads(); // publishes this row's values (cur) and the per-warp extremes
        lo = INT_MAX, hi = INT_MIN;
        for (int i = 0; i < nwarp; ++i) {
            lo = min(lo, s_red[buf + i]);
            hi = max(hi, s_red[buf + 16 + i]);
        }
        plo = elo, phi = ehi;
        float *t = prev;
        prev = cur;
        cur = t;
    }
    if (tid == 0 && p.cells) atomicAdd(p.cells, cells);
}

template <int D, bool RIG, bool LR>
__global__ void __launch_bounds__(BD_THREADS, 1) k_band_dp(const DevP pin0, const __grid_constant__ BdMaps tm0, const DevP *tab,
                                                           const BdMaps *mtab)
{
    pdl_entry();
    const DevP pin = pick_image(pin0, tab);
    const BdMaps &tm = mtab ? mtab[blockIdx.z] : tm0; // tensor maps of this image: kernel parameter, or table in HBM
    int seam;
    const DevP p = seam_view(pin, 1, &seam);
    extern __shared__ __align__(128) unsigned char bd_smem[];
    unsigned char *ring = bd_smem;
    unsigned *nrg = reinterpret_cast<unsigned *>(ring + BD_RING_BYTES);
    float *hand = reinterpret_cast<float *>(nrg + BD_HMAX);
    BdDesc *desc = reinterpret_cast<BdDesc *>(hand + 2 * BD_HANDW);
    int *hull = reinterpret_cast<int *>(desc + BD_NRING);                         // [2][BD_NCW][2], 256 bytes reserved
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(hull + 64); // [BD_NRING]
    volatile int *misc = reinterpret_cast<volatile int *>(mbar + BD_NRING);
    int *s_red = const_cast<int *>(misc) + 16;
    float *vals = reinterpret_cast<float *>(s_red + 64); // [BD_NCW][4][128]: each compute warp's last rows

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int y = tid; y < p.h; y += BD_THREADS) nrg[y] = p.nrg_pack[y];
    if (tid < BD_NRING) desc[tid].seq = -1;
    if (tid == 0) {
        for (int i = 0; i < BD_NRING; ++i) bd_mbar_init(&mbar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        misc[0] = p.h;
        misc[4] = -1;
    }
    __syncthreads();

    if (warp == 0) {
        bd_producer<D, RIG>(p, tm, ring, nrg, desc, hull, mbar, misc, lane, seam);
    } else {
        bd_compute<D, RIG, LR>(p, ring, hand, vals + (size_t) (warp - 1) * 4 * 128, desc, hull, misc + 4, warp - 1, lane);
    }
    __syncthreads();
    const int y_from = misc[0];
    if (p.tail) { // the multi-SM tail kernel (band_tail.cuh) takes the rows this CTA cannot tile
        if (tid == 0) p.tail[0] = y_from, p.tail[1] = misc[1], p.tail[2] = misc[2];
        return;
    }
    if (y_from < p.h) {
        if (BD_PROF && p.dbg && tid == 0) atomicAdd((unsigned long long *) &p.dbg[14], (unsigned long long) (p.h - y_from));
        constexpr int cap = BD_RING_BYTES / 8; // two row buffers in the tile ring, which nobody uses any more
        if (p.pitch <= cap)
            bd_rows_wide<D, RIG, LR>(p, y_from, misc[1], misc[2], reinterpret_cast<float *>(ring), reinterpret_cast<float *>(ring) + cap, s_red);
        else
            update_rows_generic(p, y_from, misc[1], misc[2], s_red);
    }
}

} // namespace b200c
