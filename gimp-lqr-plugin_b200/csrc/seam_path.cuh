// seam_path.cuh -- K3, the minimum-cost seam (liblqr lqr_carver_build_vpath) on the compact maps.
//
// Arg-min over the last row of m (one coalesced pass of the CTA), then the parent chase upwards.  With parent
// offsets stored per cell in current coordinates a chase step is ONE dependent load: x += pdx[y][x].  The loads
// are made shared-memory loads: rows are staged in chunks of R rows; the seam moves at most delta_x columns per
// row, so the chunk after next is known to stay within 2*R*delta_x columns of the column the chase holds when it
// ENTERS the current chunk -- the other warps fetch that window with cp.async while thread 0 chases the current
// chunk out of shared memory.  The seam columns are collected in shared memory and written back coalesced.
#pragma once
#include "carver_kernels.cuh"

namespace b200c {

#define SP_THREADS 512
#define SP_TILE_BYTES 73728
#define SP_HMAX 8192

static constexpr size_t sp_smem_bytes() { return 2 * SP_TILE_BYTES + SP_HMAX * 4 + 64; }

__device__ __forceinline__ int sp_rows_per_chunk(int delta_x)
{
    // a chunk must take longer to chase (~35 cycles per row) than the next one takes to arrive (~2 us), and two
    // tiles of R rows x (4*R*delta_x + 32) bytes must fit SP_TILE_BYTES each
    int r = (int) sqrtf(17000.f / (float) max(delta_x, 1));
    return max(4, min(128, r));
}

__device__ __forceinline__ void sp_cp_async16(void *dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned) __cvta_generic_to_shared(dst_smem)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void sp_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

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

__global__ void __launch_bounds__(SP_THREADS, 1) k_seam_path(DevP p)
{
    extern __shared__ __align__(128) unsigned char sp_smem[];
    signed char *tiles = reinterpret_cast<signed char *>(sp_smem);
    int *sx = reinterpret_cast<int *>(sp_smem + 2 * SP_TILE_BYTES);
    int *ctl = sx + SP_HMAX; // [0] column the chase holds at the last chunk boundary
    __shared__ float s_v[32];
    __shared__ int s_x[32];

    const int R = sp_rows_per_chunk(p.delta_x);
    const int reach = 2 * R * p.delta_x; // two chunks of drift
    const int tw = min((2 * reach + 1 + 15 + 15) & ~15, p.pitch); // bytes per staged row; covers any 16-aligned start
    const bool collect = p.h <= SP_HMAX;
    // window of a chunk whose rows lie within 2R rows of the row where the chase held column xc
    auto window_lo = [&](int xc) { return min(max(xc - reach, 0) & ~15, p.pitch - tw); };

    const int x_end = last_row_argmin(p, s_v, s_x);
    if (threadIdx.x == 0) ctl[0] = x_end;
    __syncthreads();
    int centre = ctl[0]; // centre of the window of the chunk about to be chased
    int x = centre;      // column the chase enters the chunk with (thread 0)
    int y_bot = p.h - 1;
    if (y_bot >= 1) sp_stage(p, tiles, max(y_bot - R + 1, 1), y_bot, window_lo(centre), tw, 0);
    sp_cp_async_wait_all();
    __syncthreads();

    for (int buf = 0; y_bot >= 1; buf ^= 1, y_bot -= R) {
        const int y_top = max(y_bot - R + 1, 1);
        if (threadIdx.x == 0) {
            const signed char *t = tiles + (size_t) buf * SP_TILE_BYTES;
            const int wlo = window_lo(centre);
            // one dependent shared-memory load per row: xx += pdx[y][xx].  A dead parent (PDX_NONE, -128) cannot occur
            // on a seam (the band DP re-evaluates every cell whose parent was carved); it is flagged after the chunk
            // and the column is clamped there, so a corrupted map can never walk the chase out of its window.
            // (explicit ld.shared with a 32-bit address: the chain is LDS -> IADD -> LDS ...)
            unsigned ta = (unsigned) __cvta_generic_to_shared(t) - (unsigned) wlo;
            int xx = x, bad = 0;
            auto lds8 = [](unsigned a) {
                int v;
                asm volatile("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(a));
                return v;
            };
            // the chased quantity is the ADDRESS of the seam's cell: a += d + tw (one 3-input add per row on the chain)
            unsigned ca = ta + (unsigned) xx;
            if (collect) {
                for (int y = y_bot; y >= y_top; --y, ta += tw) {
                    sx[y] = (int) (ca - ta);
                    const int d = lds8(ca);
                    bad |= d == B200C_PDX_NONE;
                    ca += (unsigned) d + (unsigned) tw;
                }
            } else {
                for (int y = y_bot; y >= y_top; --y, ta += tw) {
                    p.vpath_x[y] = (int) (ca - ta);
                    const int d = lds8(ca);
                    bad |= d == B200C_PDX_NONE;
                    ca += (unsigned) d + (unsigned) tw;
                }
            }
            xx = (int) (ca - ta);
            if (bad || xx < 0 || xx > p.w - 1) {
                atomicOr(p.err, 2);
                xx = min(max(xx, 0), p.w - 1);
            }
            ctl[0] = xx;
        } else if (y_bot - R >= 1) {
            // the next chunk lies within 2R rows of this chunk's entry row: its window is centred on `x` at entry,
            // which every thread knows (ctl[0] of the previous boundary) -- fetched while thread 0 chases
            sp_stage(p, tiles + (size_t) (buf ^ 1) * SP_TILE_BYTES, max(y_bot - 2 * R + 1, 1), y_bot - R, window_lo(x), tw, 1);
            sp_cp_async_wait_all();
        }
        __syncthreads();
        centre = x;      // the window just fetched was centred on this chunk's entry column
        x = ctl[0];      // entry column of the next chunk
        __syncthreads(); // ctl[0] is rewritten in the next round
    }
    if (threadIdx.x == 0) {
        if (collect) sx[0] = x; else p.vpath_x[0] = x;
    }
    __syncthreads();
    if (collect)
        for (int y = threadIdx.x; y < p.h; y += SP_THREADS) p.vpath_x[y] = sx[y];
}

} // namespace b200c
