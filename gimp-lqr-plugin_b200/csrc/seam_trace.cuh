// seam_trace.cuh -- K3, the minimum-cost seam (liblqr lqr_carver_build_vpath) on the compact maps, as two
// kernels that replace the single-CTA chase of seam_path.cuh on the per-seam critical path.
//
// With parent offsets stored per cell a chase step is x += pdx[y][x]: h dependent loads.  The chain is cut by
// pointer jumping on two levels: BLOCKS of R rows (ST_R = 32 for delta_x <= 3, 28 for delta_x 4) made of four
// SUB-BLOCKS of q = R / 4 rows.
//
//   k_seam_jumps   grid (column chunks, row blocks[, images]): every CTA stages the parent offsets of its R rows x
//                  (ST_COLS + 2 R delta_x) columns in shared memory; every thread walks the four sub-blocks from its
//                  column (four independent chains of q steps), a few threads also the sub-blocks of the halo columns
//                  the composition can reach, and the block jump is the composition of the four sub-block jumps:
//                      J8[4 b + j][x] = (column a path entering sub-block j of block b at column x leaves it with) - x
//                      J32[b][x]      = the same over the whole block
//                  one signed byte each (ST_BAD / 0 when the path meets a parent that was carved away: the walk below
//                  notices).  The CTAs of the bottom block also leave the arg-min of their 256 columns of the last row
//                  of m.  All SMs, ~1 byte read per cell.
//   k_seam_chase   one CTA per image: arg-min over the partial arg-mins; then the chase through J32, one dependent
//                  shared-memory load per BLOCK: the jump rows of a GROUP of blocks are fetched around the column the
//                  chase holds (a path drifts at most R delta_x columns per block, so block k of the group needs
//                  2 k R delta_x + 1 columns: a triangle of R delta_x G^2 bytes, one bulk copy (TMA) per row); then
//                  one thread per block resolves the entry columns of its sub-blocks through J8 (3 dependent loads from
//                  L2) and one thread per sub-block walks its q rows through the parent map itself (q dependent loads
//                  from L2) and writes the seam column of every row.  A single SM pulls only ~32-50 bytes per clock
//                  out of L2 (tools/ubench_stage.cu), so everything but the triangle is read in place.
//
// Blocks count from the bottom: block b walks the parents of rows ybot(b) = h-1 - b R down to ytop(b) = max(ybot - R + 1, 1);
// sub-block j of it the rows ybot - j q down to max(ybot - j q - q + 1, ytop).
#pragma once
#include "carver_kernels.cuh"

namespace b200c {

#define ST_COLS 256
#define ST_THREADS 256
#define ST_CHASE_THREADS 1024
#define ST_HMAX 8192
#define ST_WMAX 16384 // partial arg-mins: one per ST_COLS columns, at most ST_NPART
#define ST_NPART 64
#define ST_MAXBLK ((ST_HMAX + 27) / 28 + 1)
#define ST_CHASE_DYN (176 * 1024) // staging area of the chase kernel (jump triangle)
#define ST_BAD (-128)             // sub-block jump of a path that meets a dead parent
#define ST_PAD 1024               // bytes in front of k_seam_jumps' tile (see there)

__host__ __device__ inline int st_rows(int delta_x) { return delta_x <= 3 ? 32 : 28; } // R * delta_x <= 127: a jump fits a byte
__host__ __device__ inline int st_nblk(int h, int delta_x) { return h > 1 ? (h - 1 + st_rows(delta_x) - 1) / st_rows(delta_x) : 0; }
// the jump buffer of a carver: J32 rows, then J8 rows, then the partial arg-mins (ST_NPART values, ST_NPART columns)
__host__ __device__ inline size_t st_jump_bytes(int h, int delta_x, int pitch) { return (size_t) 5 * st_nblk(h, delta_x) * pitch + 64 + ST_NPART * 8; }
__device__ __forceinline__ signed char *st_j8(const DevP &p, int nblk) { return p.jump + (size_t) nblk * p.pitch; }
__device__ __forceinline__ float *st_part_v(const DevP &p, int nblk) { return reinterpret_cast<float *>(p.jump + (((size_t) 5 * nblk * p.pitch + 63) & ~(size_t) 63)); }
__device__ __forceinline__ int *st_part_x(const DevP &p, int nblk) { return reinterpret_cast<int *>(st_part_v(p, nblk) + ST_NPART); }
static inline size_t st_jump_smem(int delta_x)
{
    const int D = delta_x ? delta_x : 1, R = st_rows(delta_x);
    return ST_PAD + (size_t) R * (ST_COLS + 2 * R * D + 32) + 4 * (size_t) (ST_COLS + 6 * (R / 4) * D + 16);
}
static inline size_t st_chase_smem() { return (size_t) ST_CHASE_DYN + (4 * ST_MAXBLK + 4) * 4 + (ST_MAXBLK + 1) * 4 * 3 + 512; }

__device__ __forceinline__ void st_cp16(void *dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned) __cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void st_cp_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// the view of the iteration that is about to start: the seam counter is advanced by k_seam_chase, which runs after
// k_seam_jumps
__device__ __forceinline__ DevP seam_view_next(DevP p)
{
    if (p.dyn) p.w -= *reinterpret_cast<volatile int *>(p.dyn) + 1;
    return p;
}

__device__ __forceinline__ unsigned st_saddr(const void *q) { return (unsigned) __cvta_generic_to_shared(q); }
// one row of the jump triangle: global -> shared bulk copy (TMA), completion counted in bytes on `mbar`
__device__ __forceinline__ void st_bulk(void *dst_smem, const void *src, unsigned bytes, void *mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(st_saddr(dst_smem)),
                 "l"(src), "r"(bytes), "r"(st_saddr(mbar))
                 : "memory");
}
__device__ __forceinline__ bool st_mbar_wait(void *mbar, unsigned parity)
{
    unsigned long long t0 = 0;
    for (unsigned tries = 0;; ++tries) {
        unsigned ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok)
                     : "r"(st_saddr(mbar)), "r"(parity)
                     : "memory");
        if (ok) return true;
        if ((tries & 1023u) == 1023u) { // a copy that never lands is a bug, not a reason to hang the GPU: give up after 2 s
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 2000000000ull) return false;
        }
    }
}

__global__ void __launch_bounds__(ST_THREADS) k_seam_jumps(const DevP pin0, const DevP *tab)
{
    pdl_entry();
    const DevP pin = pick_image(pin0, tab);
    extern __shared__ __align__(16) unsigned char st_smem[];
    __shared__ float s_v[ST_THREADS / 32];
    __shared__ int s_x[ST_THREADS / 32];
    const int D = max(pin.delta_x, 1), R = st_rows(pin.delta_x), q = R >> 2, reach = R * D;
    const int b = blockIdx.y, c0 = blockIdx.x * ST_COLS, tid = threadIdx.x;
    const int nblk = st_nblk(pin.h, pin.delta_x);
    const int ybot = pin.h - 1 - b * R, ytop = max(ybot - R + 1, 1), rows = ybot - ytop + 1;
    const int tlo = max(c0 - reach, 0) & ~15;                           // first staged column (16-byte aligned)
    const int thi = min((c0 + ST_COLS + reach + 15) & ~15, pin.pitch);  // one past the last
    const int tw = thi - tlo, pieces = max(tw, 0) >> 4;
    // the tile sits ST_PAD bytes into the buffer: a walk that meets a dead parent (offset -128) goes on from a column that
    // does not matter, possibly left of the tile, before it is flagged
    unsigned char *tile = st_smem + ST_PAD;
    // the copies are issued before the current width is known (it takes a trip to the seam counter in HBM): the tile
    // only depends on the pitch
    for (int r = tid >> 5; r < rows; r += ST_THREADS / 32) { // a warp per row: no index arithmetic per piece
        const signed char *src = pin.pdx + (size_t) (ybot - r) * pin.pitch + tlo;
        for (int c = tid & 31; c < pieces; c += 32) st_cp16(tile + (size_t) r * tw + (c << 4), src + (c << 4));
    }
    const DevP p = seam_view_next(pin);
    if (c0 >= p.w) { // (uniform over the CTA) nothing to do here any more: the image has shrunk past this chunk
        st_cp_wait();
        return;
    }
    const int x = c0 + tid;
    // the bottom block's CTAs: arg-min of their columns of the last row (the rule of last_row_argmin)
    float av = 536870912.f; // (float) (1 << 29)
    int ax = -1;
    if (b == 0) {
        if (x < p.w) {
            const float v = p.m[(size_t) (p.h - 1) * p.pitch + x];
            if (v < av || (v == av && p.leftright == 1)) av = v, ax = x;
        }
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_down_sync(0xffffffffu, av, off);
            const int ox = __shfl_down_sync(0xffffffffu, ax, off);
            if (seam_better(ov, ox, av, ax, p.leftright)) av = ov, ax = ox;
        }
        if ((tid & 31) == 0) s_v[tid >> 5] = av, s_x[tid >> 5] = ax;
    }
    st_cp_wait();
    __syncthreads();
    if (b == 0 && tid == 0) {
        for (int i = 1; i < ST_THREADS / 32; ++i)
            if (seam_better(s_v[i], s_x[i], av, ax, p.leftright)) av = s_v[i], ax = s_x[i];
        st_part_v(p, nblk)[blockIdx.x] = av;
        st_part_x(p, nblk)[blockIdx.x] = ax;
    }

    const signed char *t = reinterpret_cast<const signed char *>(tile);
    const int Wj = ST_COLS + 6 * q * D + 16, jbase = c0 - 3 * q * D; // sub-block jumps of the columns a composition can reach
    signed char *j8s = reinterpret_cast<signed char *>(tile) + (size_t) R * tw;
    // The four sub-blocks from this thread's own column: four independent chains.  A step is one shared-memory load and
    // one add: the tile index moves on by a row plus the parent offset, the minimum of the offsets seen tells afterwards
    // whether a dead parent (B200C_PDX_NONE = -128, smaller than any live offset) was among them.
    int own[4] = {0, 0, 0, 0};
    if (x < p.w) {
        int a[4], mn[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = j * q * tw + (x - tlo), mn[j] = 0;
        if (rows == R) {
            for (int sidx = 0; sidx < q; ++sidx) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int d = t[a[j]];
                    a[j] += tw + d;
                    mn[j] = min(mn[j], d);
                }
            }
        } else {
            for (int sidx = 0; sidx < q; ++sidx) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j * q + sidx < rows) {
                        const int d = t[a[j]];
                        a[j] += tw + d;
                        mn[j] = min(mn[j], d);
                    }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = min(max(rows - j * q, 0), q); // rows of sub-block j
            own[j] = mn[j] == B200C_PDX_NONE ? ST_BAD : a[j] - (j * q + n) * tw - (x - tlo);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) j8s[j * Wj + (x - jbase)] = (signed char) own[j];
    // ... and from the halo columns: sub-block j can be entered j q D columns either side of the CTA's own columns
    const int qd = q * D;
    for (int i = tid; i < 12 * qd; i += ST_THREADS) {
        const int j = i < 2 * qd ? 1 : (i < 6 * qd ? 2 : 3);
        const int o = i - (j == 1 ? 0 : (j == 2 ? 2 * qd : 6 * qd));
        const int col = o < j * qd ? c0 - j * qd + o : c0 + ST_COLS + (o - j * qd);
        if (col < 0 || col >= p.w) continue;
        const int n = min(max(rows - j * q, 0), q);
        int a = j * q * tw + (col - tlo), mn = 0;
        for (int r = 0; r < n; ++r) {
            const int d = t[a];
            a += tw + d;
            mn = min(mn, d);
        }
        j8s[j * Wj + (col - jbase)] = (signed char) (mn == B200C_PDX_NONE ? ST_BAD : a - (j * q + n) * tw - (col - tlo));
    }
    __syncthreads();
    if (x >= p.w) return;
    // the block jump: the four sub-block jumps composed (a path that meets a dead parent gets 0: the chase kernel walks
    // every row again and notices)
    int xc = x;
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int v = bad ? 0 : (int) j8s[j * Wj + (xc - jbase)];
        bad |= v == ST_BAD;
        xc += v == ST_BAD ? 0 : v;
    }
    p.jump[(size_t) b * p.pitch + x] = bad ? (signed char) 0 : (signed char) (xc - x);
    signed char *J8 = st_j8(p, nblk) + (size_t) (4 * b) * p.pitch + x;
#pragma unroll
    for (int j = 0; j < 4; ++j) J8[(size_t) j * p.pitch] = (signed char) own[j];
}

#ifdef BD_PROFILE
#define ST_MARK(i)                                                                                       \
    do {                                                                                                 \
        if (pin.dbg && threadIdx.x == 0) {                                                               \
            const long long t_now = clock64();                                                           \
            atomicAdd((unsigned long long *) &pin.dbg[16 + (i)], (unsigned long long) (t_now - t_mark)); \
            t_mark = t_now;                                                                              \
        }                                                                                                \
    } while (0)
#else
#define ST_MARK(i) do {} while (0)
#endif

__global__ void __launch_bounds__(ST_CHASE_THREADS, 1) k_seam_chase(const DevP pin0, const DevP *tab)
{
    pdl_entry();
    const DevP pin = pick_image(pin0, tab);
    long long t_mark = clock64();
    (void) t_mark;
    if (pin.dyn && threadIdx.x == 0) advance_seam(pin); // this iteration's seam
    __syncthreads();
    const DevP p = seam_view(pin, 0);
    extern __shared__ __align__(16) unsigned char st_smem[];
    unsigned char *stage = st_smem;                                      // [ST_CHASE_DYN]
    int *sub = reinterpret_cast<int *>(st_smem + ST_CHASE_DYN);          // [4 nblk] column entering sub-block j of block b
    int *ent = sub + 4 * ST_MAXBLK + 4;                                  // [nblk + 1] column entering block b
    int *roff = ent + ST_MAXBLK + 1;                                     // per staged row: offset in `stage`
    int *rcol = roff + ST_MAXBLK + 1;                                    // ... and the stage index of its column 0
    __shared__ float s_v[32];
    __shared__ int s_x[32];
    __shared__ int s_bad, s_tot;
    __shared__ __align__(8) unsigned long long mbar;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int D = max(p.delta_x, 1), R = st_rows(p.delta_x), q = R >> 2, reach = R * D;
    const int nblk = st_nblk(p.h, p.delta_x);
    if (nblk == 0) { // a one-row image: no parents to follow
        const int x_end = last_row_argmin(p, s_v, s_x);
        if (tid == 0) p.vpath_x[0] = x_end;
        return;
    }
    if (tid == 0) {
        s_bad = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(st_saddr(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    ST_MARK(0);
    if (wid == 0) { // the arg-min of the last row out of k_seam_jumps' partial arg-mins
        const int npart = (p.w + ST_COLS - 1) / ST_COLS;
        const float *pv = st_part_v(p, nblk);
        const int *px = st_part_x(p, nblk);
        float best = lane < npart ? pv[lane] : 0.f;
        int bx = lane < npart ? px[lane] : -1;
        if (lane + 32 < npart && seam_better(pv[lane + 32], px[lane + 32], best, bx, p.leftright)) best = pv[lane + 32], bx = px[lane + 32];
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_down_sync(0xffffffffu, best, off);
            const int ox = __shfl_down_sync(0xffffffffu, bx, off);
            if (seam_better(ov, ox, best, bx, p.leftright)) best = ov, bx = ox;
        }
        if (lane == 0) ent[0] = bx < 0 ? 0 : bx;
    }
    __syncthreads();
    ST_MARK(1);

    // ---- the chase through the block jumps, a group of G blocks at a time
    int G = min(nblk, (int) sqrtf((float) ST_CHASE_DYN / (float) reach)); // G^2 reach + 48 G <= ST_CHASE_DYN
    while (G > 1 && G * G * reach + G * 48 > ST_CHASE_DYN) --G;
    unsigned phase = 0;
    for (int b0 = 0; b0 < nblk; b0 += G, phase ^= 1) {
        const int g = min(G, nblk - b0), xg = ent[b0];
        // layout of the triangle: row k holds columns [a_k, a_k + n_k) with a_k 16-byte aligned; the row offsets are an
        // exclusive prefix sum of the widths (two-level warp scan over at most 1024 rows)
        int wdt = 0, a = 0;
        if (tid < g) {
            a = max(xg - tid * reach, 0) & ~15;
            wdt = min((xg + tid * reach + 16) & ~15, p.pitch) - a;
        }
        int incl = wdt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_x[wid] = incl;
        __syncthreads();
        if (tid < g) {
            int base = 0;
            for (int i = 0; i < wid; ++i) base += s_x[i];
            const int off = base + incl - wdt;
            roff[tid] = off;
            rcol[tid] = off - a;
            if (tid == g - 1) s_tot = off + wdt;
            if (b0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the stage was read by the chase of the last group
            st_bulk(stage + off, p.jump + (size_t) (b0 + tid) * p.pitch + a, (unsigned) wdt, &mbar);
        }
        __syncthreads();
        ST_MARK(6);
        if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st_saddr(&mbar)), "r"((unsigned) s_tot) : "memory");
        if (!st_mbar_wait(&mbar, phase) && tid == 0) atomicOr(p.err, 8);
        ST_MARK(2);
        if (tid == 0) { // the chain: one shared-memory load + one add per block
            const signed char *S = reinterpret_cast<const signed char *>(stage);
            int x = xg;
#pragma unroll 4
            for (int k = 0; k < g; ++k) {
                x += S[rcol[k] + x];
                ent[b0 + k + 1] = x;
            }
        }
        __syncthreads();
        ST_MARK(3);
    }

    // ---- the entry columns of the sub-blocks: one thread per block, three dependent loads through J8
    const int wmax = max(p.w - 1, 0);
    for (int b = tid; b < nblk; b += ST_CHASE_THREADS) {
        const signed char *J8 = st_j8(p, nblk) + (size_t) (4 * b) * p.pitch;
        int e = min(max(ent[b], 0), wmax);
        sub[4 * b] = e;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int v = __ldcg(J8 + (size_t) j * p.pitch + e);
            if (v == ST_BAD) s_bad = 1;
            e = min(max(e + (v == ST_BAD ? 0 : v), 0), wmax);
            sub[4 * b + j + 1] = e;
        }
    }
    __syncthreads();
    ST_MARK(4);
    // ---- the rows: one thread per sub-block follows the parents themselves and writes the seam
    for (int t = tid; t < 4 * nblk; t += ST_CHASE_THREADS) { // (more than 1024 sub-blocks: delta_x 4 and over 7168 rows)
        const int b = t >> 2, j = t & 3;
        const int ybot = p.h - 1 - b * R, ytop = max(ybot - R + 1, 1);
        const int yhi = ybot - j * q, ylo = max(yhi - q + 1, ytop);
        int x = sub[t];
        bool bad = false;
        for (int y = yhi; y >= ylo; --y) {
            p.vpath_x[y] = x;
            const int d = __ldcg(p.pdx + (size_t) y * p.pitch + x);
            bad |= d == B200C_PDX_NONE;
            x = min(max(x + (d == B200C_PDX_NONE ? 0 : d), 0), wmax); // a live parent is at most delta_x columns away
        }
        if (bad || x != (j == 3 ? ent[b + 1] : sub[t + 1])) s_bad = 1;
    }
    if (tid == 0) p.vpath_x[0] = ent[nblk];
    __syncthreads();
    ST_MARK(5);
    if (tid == 0 && s_bad) atomicOr(p.err, 2);
}

} // namespace b200c
