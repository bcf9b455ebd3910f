// seam_path.cuh -- K3, the minimum-cost seam (liblqr lqr_carver_build_vpath) on the compact maps.
//
// Arg-min over the last row of m (one coalesced pass of the CTA), then the parent chase upwards.  With parent
// offsets stored per cell in current coordinates a chase step is one dependent load, x += pdx[y][x] -- but h of
// them in a row is still h x (shared-memory latency + add) ~ 48 cycles each.  The chain is cut by POINTER JUMPING:
//
//   * rows are staged in tiles of R rows (R a multiple of 4); the seam moves at most delta_x columns per row, so the
//     tile two chunks ahead is known to stay within 3*R*delta_x columns of the column the chase holds when it
//     ENTERS the current chunk: warps 1..15 fetch that window with cp.async;
//   * the same warps turn the tile one chunk ahead into a table of 4-row jumps, J4[b][x] = where a path entering
//     block b (4 rows) at column x leaves it -- every entry an independent 4-load chain, thousands in parallel;
//   * thread 0 chases the current chunk through its J4 table: one dependent load per FOUR rows;
//   * the lanes of warp 0 then fill in the three rows inside every block from the block entries, in parallel.
//
// The seam columns are collected in shared memory and written back coalesced.
#pragma once
#include "carver_kernels.cuh"

namespace b200c {

#define SP_THREADS 512
#define SP_TILE_BYTES 32768 // one staged tile of pdx rows
#define SP_JUMP_BYTES 16384 // one table of 4-row jumps (int16 per column and block)
#define SP_HMAX 8192
#define SP_BAD 32767        // jump-table entry of a path that meets a dead parent or leaves the window

static constexpr size_t sp_smem_bytes() { return 3 * SP_TILE_BYTES + 2 * SP_JUMP_BYTES + SP_HMAX * 4 + 256; }

// rows per chunk: a multiple of 4 with R * (6 * R * delta_x + 32) <= SP_TILE_BYTES
__device__ __forceinline__ int sp_rows_per_chunk(int delta_x)
{
    const int r = (int) sqrtf(5300.f / (float) max(delta_x, 1)) & ~3;
    return max(4, min(60, r)); // at most 15 blocks of 4 rows: one per helper warp
}

__device__ __forceinline__ void sp_cp_async16(void *dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned) __cvta_generic_to_shared(dst_smem)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void sp_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void sp_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void sp_cp_async_wait_but1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void sp_bar_helpers()
{
    __syncwarp();
    asm volatile("barrier.sync 1, %0;" ::"n"(SP_THREADS - 32) : "memory");
}

// stage rows [y_top, y_bot] (y_bot >= y_top), columns [wlo, wlo + tw) of pdx into `tile` (row r = y_bot - y, pitch tw);
// the work is spread over threads first .. SP_THREADS-1
__device__ __forceinline__ void sp_stage(const DevP &p, signed char *tile, int y_top, int y_bot, int wlo, int tw, int first)
{
    const int pieces = tw >> 4, n = (y_bot - y_top + 1) * pieces;
    for (int i = (int) threadIdx.x - first; i >= 0 && i < n; i += SP_THREADS - first) {
        const int r = i / pieces, c = (i - r * pieces) << 4;
        sp_cp_async16(tile + (size_t) r * tw + c, p.pdx + (size_t) (y_bot - r) * p.pitch + wlo + c);
    }
}

__device__ __forceinline__ int sp_lds8(unsigned a)
{
    int v;
    asm volatile("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int sp_lds16(unsigned a)
{
    int v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// J4[b][c] = net column offset of a path that enters tile row 4b at window column c and climbs 4 rows.  Only the
// columns a seam can reach are built: xc +- (drift + 4b + 4) * delta_x, where xc is a column the seam is known to hold
// `drift / delta_x` rows below the tile.  Warp wi of nw takes blocks wi, wi + nw, ... (a tile has at most 15 blocks:
// one per helper warp); its lanes stride the columns two at a time (two independent 4-load chains in flight).  A path
// that meets a dead parent gets jump 0: the fill-in pass of the chase re-walks every block row by row and notices.
__device__ __forceinline__ void sp_build_jumps(const signed char *tile, short *jump, int nblocks, int tw, int xc, int drift,
                                               int delta_x, int wi, int nw, int lane)
{
    const unsigned tbase = (unsigned) __cvta_generic_to_shared(tile);
    for (int b = wi; b < nblocks; b += nw) {
        const int half = drift + (4 * b + 4) * delta_x;
        const int c_lo = max(xc - half, 0), c_hi = min(xc + half, tw - 1);
        const unsigned t0 = tbase + (unsigned) (4 * b * tw);
        short *jb = jump + b * tw;
        for (int c = c_lo + lane; c <= c_hi; c += 64) {
            const int c2 = min(c + 32, c_hi); // (a duplicate of the last column when the range ends in between)
            int x0 = c, x1 = c2;
            bool bad0 = false, bad1 = false;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int d0 = sp_lds8(t0 + (unsigned) (k * tw + x0)), d1 = sp_lds8(t0 + (unsigned) (k * tw + x1));
                bad0 |= d0 == B200C_PDX_NONE, bad1 |= d1 == B200C_PDX_NONE;
                x0 = min(max(x0 + d0, 0), tw - 1), x1 = min(max(x1 + d1, 0), tw - 1);
            }
            jb[c] = bad0 ? (short) 0 : (short) (x0 - c);
            jb[c2] = bad1 ? (short) 0 : (short) (x1 - c2);
        }
    }
}

__global__ void __launch_bounds__(SP_THREADS, 1) k_seam_path(const DevP pin0, const DevP *tab)
{
    pdl_entry();
    const DevP pin = pick_image(pin0, tab);
    if (pin.dyn && threadIdx.x == 0) advance_seam(pin); // this iteration's seam
    __syncthreads();
    const DevP p = seam_view(pin, 0);
    extern __shared__ __align__(128) unsigned char sp_smem[];
    signed char *tiles = reinterpret_cast<signed char *>(sp_smem);                  // [3][SP_TILE_BYTES]
    short *jumps = reinterpret_cast<short *>(sp_smem + 3 * SP_TILE_BYTES);           // [2][SP_JUMP_BYTES / 2]
    int *sx = reinterpret_cast<int *>(sp_smem + 3 * SP_TILE_BYTES + 2 * SP_JUMP_BYTES);
    int *ctl = sx + SP_HMAX; // [0] column at the last chunk boundary, [1..16] block entry columns of the current chunk
    __shared__ float s_v[32];
    __shared__ int s_x[32];

    const int R = sp_rows_per_chunk(p.delta_x);
    const int reach = 3 * R * p.delta_x; // three chunks of drift
    const int tw = min((2 * reach + 1 + 15 + 15) & ~15, p.pitch); // bytes per staged row; covers any 16-aligned start
    const bool collect = p.h <= SP_HMAX;
    // window of a tile whose rows lie within 3R rows of the row where the chase held column xc
    auto window_lo = [&](int xc) { return min(max(xc - reach, 0) & ~15, p.pitch - tw); };
    auto put = [&](int y, int x) {
        if (collect)
            sx[y] = x;
        else
            p.vpath_x[y] = x;
    };
    const int tid = threadIdx.x, lane = tid & 31;

    const int x_end = last_row_argmin(p, s_v, s_x);
    if (tid == 0) ctl[0] = x_end;
    __syncthreads();
    // chunk j covers rows ybot(j) = h-1 - j*R down to max(ybot - R + 1, 1); its tile lives in tiles[j % 3], its jump
    // table in jumps[j % 2]
    const int nchunks = p.h > 1 ? (p.h - 1 + R - 1) / R : 0;
    auto ybot = [&](int j) { return p.h - 1 - j * R; };
    auto ytop = [&](int j) { return max(p.h - 1 - j * R - R + 1, 1); };
    int c_cur = ctl[0], c_nxt = c_cur; // the columns the windows of tile j and tile j+1 are centred on
    int x = c_cur;                     // column the chase enters chunk j with
    // prologue: tiles 0 and 1 (both centred on the arg-min), jump table 0
    if (nchunks > 0) sp_stage(p, tiles, ytop(0), ybot(0), window_lo(c_cur), tw, 0);
    if (nchunks > 1) sp_stage(p, tiles + SP_TILE_BYTES, ytop(1), ybot(1), window_lo(c_nxt), tw, 0);
    sp_cp_async_commit();
    sp_cp_async_wait_all();
    __syncthreads();
    if (nchunks > 0) // the chase enters tile 0 at the arg-min itself
        sp_build_jumps(tiles, jumps, (ybot(0) - ytop(0) + 1) >> 2, tw, c_cur - window_lo(c_cur), 0, p.delta_x, tid >> 5, SP_THREADS / 32, lane);
    __syncthreads();

    for (int j = 0; j < nchunks; ++j) {
        const int yb = ybot(j), rows = yb - ytop(j) + 1, nblocks = rows >> 2;
        const signed char *T = tiles + (size_t) (j % 3) * SP_TILE_BYTES;
        const int wlo = window_lo(c_cur);
        const int c_nn = x; // the tile after next is centred on this chunk's entry column
        if (tid >= 32) {
            // ---- warps 1..15: fetch tile j+2, then turn tile j+1 (fetched during the previous chunk) into jumps
            if (j + 2 < nchunks)
                sp_stage(p, tiles + (size_t) ((j + 2) % 3) * SP_TILE_BYTES, ytop(j + 2), ybot(j + 2), window_lo(c_nn), tw, 32);
            sp_cp_async_commit();
            sp_cp_async_wait_but1();
            sp_bar_helpers(); // tile j+1 has landed for every helper
            if (j + 1 < nchunks) // the chase holds column x now, R rows below tile j+1
                sp_build_jumps(tiles + (size_t) ((j + 1) % 3) * SP_TILE_BYTES, jumps + (size_t) ((j + 1) & 1) * (SP_JUMP_BYTES / 2),
                               (ybot(j + 1) - ytop(j + 1) + 1) >> 2, tw, x - window_lo(c_nxt), R * p.delta_x, p.delta_x,
                               (tid >> 5) - 1, SP_THREADS / 32 - 1, lane);
        } else {
            // ---- warp 0: thread 0 chases the chunk through its jump table, one dependent load per 4 rows
            const short *J = jumps + (size_t) (j & 1) * (SP_JUMP_BYTES / 2);
            const unsigned jbase = (unsigned) __cvta_generic_to_shared(J);
            const unsigned tbase = (unsigned) __cvta_generic_to_shared(T);
            if (lane == 0) {
                // the chased quantity is the ADDRESS of the jump entry: a += 2 * (d + tw), one load + one add per block
                unsigned a = jbase + 2u * (unsigned) (x - wlo);
                for (int b = 0; b < nblocks; ++b) {
                    ctl[1 + b] = (int) a;
                    a += 2u * (unsigned) (sp_lds16(a) + tw);
                }
                ctl[1 + nblocks] = (int) a;
                int xx = (int) ((a - jbase) >> 1) - nblocks * tw; // window column after the last full block
                bool bad = false;
                // rows past the last full block (only the topmost chunk): plain steps
                for (int r = 4 * nblocks; r < rows; ++r) {
                    put(yb - r, xx + wlo);
                    const int d = sp_lds8(tbase + (unsigned) (r * tw + xx));
                    bad |= d == B200C_PDX_NONE;
                    xx = min(max(xx + (d == B200C_PDX_NONE ? 0 : d), 0), tw - 1);
                }
                if (bad) atomicOr(p.err, 2);
                ctl[0] = min(max(xx + wlo, 0), p.w - 1);
            }
            __syncwarp();
            // the rows inside the blocks, from the block entries: lane b walks block b row by row -- and checks that it
            // comes out where the jump said (a dead parent on the way shows here)
            for (int b = lane; b < nblocks; b += 32) {
                int xx = (int) (((unsigned) ctl[1 + b] - jbase) >> 1) - b * tw;
                const int x_out = (int) (((unsigned) ctl[2 + b] - jbase) >> 1) - (b + 1) * tw;
                bool bad = false;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    put(yb - (4 * b + k), xx + wlo);
                    const int d = sp_lds8(tbase + (unsigned) ((4 * b + k) * tw + xx));
                    bad |= d == B200C_PDX_NONE;
                    xx = min(max(xx + d, 0), tw - 1);
                }
                if (bad || xx != x_out) atomicOr(p.err, 2);
            }
        }
        __syncthreads();
        c_cur = c_nxt;
        c_nxt = c_nn;
        x = ctl[0];
        __syncthreads(); // ctl is rewritten in the next round
    }
    if (tid == 0) put(0, x);
    __syncthreads();
    if (collect)
        for (int y = tid; y < p.h; y += SP_THREADS) p.vpath_x[y] = sx[y];
}

} // namespace b200c
