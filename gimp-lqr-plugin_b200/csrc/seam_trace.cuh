// seam_trace.cuh -- K3, the minimum-cost seam (liblqr lqr_carver_build_vpath) on the compact maps, as two
// kernels that replace the single-CTA chase of seam_path.cuh on the per-seam critical path.
//
// With parent offsets stored per cell a chase step is x += pdx[y][x]: h dependent loads.  The chain is cut by
// pointer jumping over BLOCKS of R rows (ST_R = 32 for delta_x <= 3, 28 for delta_x 4):
//
//   k_seam_jumps   grid (column chunks, row blocks[, images]): every CTA stages the parent offsets of its R rows x
//                  (ST_COLS + 2 R delta_x) columns in shared memory and one thread per column walks them:
//                  J[b][x] = (column a path entering block b at column x leaves it with) - x, one signed byte,
//                  (0 when the path meets a parent that was carved away: the re-walk below notices).  All SMs, ~1 byte
//                  read per cell.
//   k_seam_chase   one CTA per image: arg-min of the last row of m; then the chase, one dependent shared-memory load
//                  per BLOCK: the jump rows of a GROUP of blocks are fetched around the column the chase holds (a path
//                  drifts at most R delta_x columns per block, so block k of the group needs 2 k R delta_x + 1 columns:
//                  a triangle of R delta_x G^2 bytes); then every block is re-walked row by row from its entry column by
//                  one thread (blocks in parallel, parent offsets staged the same way) to get the seam column of every
//                  row, and vpath_x is written back coalesced.
//
// Blocks count from the bottom: block b walks the parents of rows ybot(b) = h-1 - b R down to ytop(b) = max(ybot - R + 1, 1).
#pragma once
#include "carver_kernels.cuh"

namespace b200c {

#define ST_COLS 256
#define ST_THREADS 256
#define ST_CHASE_THREADS 1024
#define ST_HMAX 8192
#define ST_MAXBLK ((ST_HMAX + 27) / 28 + 1)
#define ST_CHASE_DYN (176 * 1024) // staging area of the chase kernel (jump triangle, then parent tiles)

__host__ __device__ inline int st_rows(int delta_x) { return delta_x <= 3 ? 32 : 28; } // R * delta_x <= 127: a jump fits a byte
__host__ __device__ inline int st_nblk(int h, int delta_x) { return h > 1 ? (h - 1 + st_rows(delta_x) - 1) / st_rows(delta_x) : 0; }
static inline size_t st_jump_smem(int delta_x) { return (size_t) st_rows(delta_x) * (ST_COLS + 2 * st_rows(delta_x) * (delta_x ? delta_x : 1) + 32); }
static inline size_t st_chase_smem() { return (size_t) ST_CHASE_DYN + ST_HMAX * 4 + (ST_MAXBLK + 1) * 4 * 3 + 512; }

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

__global__ void __launch_bounds__(ST_THREADS) k_seam_jumps(const DevP pin0, const DevP *tab)
{
    const DevP pin = pick_image(pin0, tab);
    const DevP p = seam_view_next(pin);
    extern __shared__ __align__(16) unsigned char st_smem[];
    const int D = max(p.delta_x, 1), R = st_rows(p.delta_x), reach = R * D;
    const int b = blockIdx.y, c0 = blockIdx.x * ST_COLS;
    if (c0 >= p.w) return;
    const int ybot = p.h - 1 - b * R, ytop = max(ybot - R + 1, 1), rows = ybot - ytop + 1;
    const int tlo = max(c0 - reach, 0) & ~15;                         // first staged column (16-byte aligned)
    const int thi = min((c0 + ST_COLS + reach + 15) & ~15, p.pitch);  // one past the last
    const int tw = thi - tlo, pieces = tw >> 4;
    for (int i = threadIdx.x; i < rows * pieces; i += ST_THREADS) {
        const int r = i / pieces, c = (i - r * pieces) << 4;
        st_cp16(st_smem + (size_t) r * tw + c, p.pdx + (size_t) (ybot - r) * p.pitch + tlo + c);
    }
    st_cp_wait();
    __syncthreads();
    const int x = c0 + threadIdx.x;
    if (x >= p.w) return;
    int xx = x - tlo;
    bool bad = false;
    const signed char *t = reinterpret_cast<const signed char *>(st_smem);
    for (int r = 0; r < rows; ++r) {
        const int d = t[r * tw + xx];
        bad |= d == B200C_PDX_NONE;
        xx = min(max(xx + (d == B200C_PDX_NONE ? 0 : d), 0), tw - 1);
    }
    // a path that meets a dead parent gets jump 0: the chase kernel re-walks every block row by row and notices
    p.jump[(size_t) b * p.pitch + x] = bad ? (signed char) 0 : (signed char) (xx + tlo - x);
}

__global__ void __launch_bounds__(ST_CHASE_THREADS, 1) k_seam_chase(const DevP pin0, const DevP *tab)
{
    const DevP pin = pick_image(pin0, tab);
    if (pin.dyn && threadIdx.x == 0) advance_seam(pin); // this iteration's seam
    __syncthreads();
    const DevP p = seam_view(pin, 0);
    extern __shared__ __align__(16) unsigned char st_smem[];
    unsigned char *stage = st_smem;                                      // [ST_CHASE_DYN]
    int *sx = reinterpret_cast<int *>(st_smem + ST_CHASE_DYN);           // [ST_HMAX] seam column per row
    int *ent = sx + ST_HMAX;                                             // [nblk + 1] column entering block b
    int *roff = ent + ST_MAXBLK + 1;                                     // per staged row / tile: offset in `stage`
    int *rcol = roff + ST_MAXBLK + 1;                                    // ... and its first column
    __shared__ float s_v[32];
    __shared__ int s_x[32];
    __shared__ int s_bad;
    const int tid = threadIdx.x;
    const int D = max(p.delta_x, 1), R = st_rows(p.delta_x), reach = R * D;
    const int nblk = st_nblk(p.h, p.delta_x);
    if (tid == 0) s_bad = 0;

    const int x_end = last_row_argmin(p, s_v, s_x); // valid in thread 0
    if (tid == 0) ent[0] = x_end;
    __syncthreads();

    // ---- the chase through the jump tables, a group of G blocks at a time
    int G = 1;
    while ((G + 1) * (G + 1) * reach + (G + 1) * 48 <= ST_CHASE_DYN && G < nblk) ++G;
    for (int b0 = 0; b0 < nblk; b0 += G) {
        const int g = min(G, nblk - b0), xg = ent[b0];
        // layout of the triangle: row k holds columns [a_k, a_k + n_k) with a_k 16-byte aligned; the row offsets are an
        // exclusive prefix sum of the widths (two-level warp scan over at most 1024 rows)
        {
            const int lane = tid & 31, wid = tid >> 5;
            int wdt = 0;
            if (tid < g) {
                const int a = max(xg - tid * reach, 0) & ~15, e = min((xg + tid * reach + 16) & ~15, p.pitch);
                wdt = e - a;
            }
            int incl = wdt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) s_x[wid] = incl;
            __syncthreads();
            int base = 0;
            for (int i = 0; i < wid; ++i) base += s_x[i];
            if (tid < g) {
                const int off = base + incl - wdt;
                roff[tid] = off;
                rcol[tid] = off - (max(xg - tid * reach, 0) & ~15); // stage index of column 0 of this row
            }
        }
        __syncthreads();
        for (int k = tid >> 5; k < g; k += ST_CHASE_THREADS / 32) { // a warp per row of the triangle; copies are asynchronous
            const int a = max(xg - k * reach, 0) & ~15, e = min((xg + k * reach + 16) & ~15, p.pitch), pieces = (e - a) >> 4;
            const signed char *src = p.jump + (size_t) (b0 + k) * p.pitch + a;
            for (int i = tid & 31; i < pieces; i += 32) st_cp16(stage + roff[k] + (i << 4), src + (i << 4));
        }
        st_cp_wait();
        __syncthreads();
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
    }

    // ---- the rows inside the blocks: one thread per block re-walks it from its entry column through staged parents
    const int twf = min((2 * reach + 1 + 15 + 15) & ~15, p.pitch); // staged columns per row: covers any 16-aligned start
    const int per_pass = max(1, ST_CHASE_DYN / (R * twf));
    for (int b0 = 0; b0 < nblk; b0 += per_pass) {
        const int nb = min(per_pass, nblk - b0);
        const int pieces = twf >> 4;
        for (int i = tid; i < nb * R * pieces; i += ST_CHASE_THREADS) {
            const int bb = i / (R * pieces), rem = i - bb * (R * pieces), r = rem / pieces, c = (rem - r * pieces) << 4;
            const int b = b0 + bb, ybot = p.h - 1 - b * R, y = ybot - r;
            if (y < 1) continue;
            const int lo = min(max(ent[b] - reach, 0) & ~15, p.pitch - twf);
            st_cp16(stage + ((size_t) bb * R + r) * twf + c, p.pdx + (size_t) y * p.pitch + lo + c);
        }
        st_cp_wait();
        __syncthreads();
        if (tid < nb) {
            const int b = b0 + tid, ybot = p.h - 1 - b * R, ytop = max(ybot - R + 1, 1);
            const int lo = min(max(ent[b] - reach, 0) & ~15, p.pitch - twf);
            const signed char *t = reinterpret_cast<const signed char *>(stage) + (size_t) tid * R * twf;
            int xx = ent[b] - lo;
            bool bad = false;
            for (int y = ybot; y >= ytop; --y) {
                sx[y] = xx + lo;
                const int d = t[(ybot - y) * twf + xx];
                bad |= d == B200C_PDX_NONE;
                xx += d == B200C_PDX_NONE ? 0 : d; // a live parent is at most delta_x columns away: xx stays in the tile
            }
            if (bad || xx + lo != ent[b + 1]) s_bad = 1;
        }
        __syncthreads();
    }
    if (tid == 0) {
        sx[0] = ent[nblk];
        if (s_bad) atomicOr(p.err, 2);
    }
    __syncthreads();
    for (int y = tid; y < p.h; y += ST_CHASE_THREADS) p.vpath_x[y] = sx[y];
}

} // namespace b200c
