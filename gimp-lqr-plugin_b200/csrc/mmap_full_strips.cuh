// mmap_full_strips.cuh -- K2, the full m-map DP (liblqr lqr_carver_build_mmap) on the compact maps, as
// h/32 strip launches: the FALLBACK of the one-launch cluster kernel in mmap_cluster.cuh (B200C_CLUSTER=0, or images wider
// than 8192 columns).
//
// m[y][x] = en[y][x] + min over |dx| <= delta_x of m[y-1][x+dx] (+ rigidity term) is a chain of h dependent rows.
// The image is cut into column STRIPS of 128 columns, one warp per strip, 4 consecutive cells per lane, the row
// held in registers: a row step is two shuffles plus 4 cells of 3-input min / add per lane -- no shared-memory
// round trip and no barrier on the chain.  Strips overlap: a strip recomputes HK = rows * delta_x columns of each
// neighbour, which go stale by delta_x columns per row and are not stored (trapezoid tiling), so the strips of one
// launch are independent.  One launch advances all strips by MF_ROWS rows; the launch boundary is the only
// grid-wide synchronisation.  The energy rows of a strip are fetched up front with cp.async (they do not depend
// on the chain) and read back from shared memory row by row.
#pragma once
#include "carver_kernels.cuh"

namespace b200c {

#define MF_WARPS 4
#define MF_THREADS (MF_WARPS * 32)

__host__ __device__ inline int mf_rows(int delta_x) { return delta_x <= 1 ? 32 : (delta_x == 2 ? 16 : 8); }
__host__ __device__ inline int mf_hk(int delta_x) { return (mf_rows(delta_x) * delta_x + 3) & ~3; }
static inline size_t mf_smem_bytes(int delta_x, bool rig) { return (size_t) MF_WARPS * mf_rows(delta_x) * 512 * (rig ? 2 : 1); }

__device__ __forceinline__ void mf_cp_async16(void *dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned) __cvta_generic_to_shared(dst_smem)), "l"(src)
                 : "memory");
}

// rows [y0, y0 + rows) of the full DP for one strip per warp; row y0-1 of m is final in HBM (y0 == 0: no parents)
template <int D, bool RIG, bool LR>
__global__ void __launch_bounds__(MF_THREADS) k_mmap_full_strips(const DevP p0, int y0, int rows, const DevP *tab)
{
    const DevP p = pick_image(p0, tab);
    extern __shared__ __align__(16) unsigned char mf_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int R = mf_rows(D), HK = mf_hk(D), S = 128 - 2 * HK;
    const int strip = blockIdx.x * MF_WARPS + warp;
    const int x0 = strip * S - HK + 4 * lane; // first of this lane's 4 columns
    const int wlim = min((p.w + 4 + 3) & ~3, p.pitch); // the image plus the +inf sentinel columns a parent scan can reach
    if (strip * S >= wlim) return;             // no interior column to produce (warp-uniform)
    const bool inmem = x0 >= 0 && x0 < p.pitch; // pitch is a multiple of 4: a lane is inside or outside as a whole
    const bool interior = 4 * lane >= HK && 4 * lane < 128 - HK && inmem && x0 < wlim;
    const float inf = __int_as_float(0x7f800000);
    const unsigned full = 0xffffffffu;

    float *es = reinterpret_cast<float *>(mf_smem) + (size_t) warp * R * 128 * (RIG ? 2 : 1);
    float *gs = es + (size_t) R * 128;
    if (inmem) {
        for (int r = 0; r < rows; ++r) {
            const size_t o = (size_t) (y0 + r) * p.pitch + x0;
            mf_cp_async16(es + r * 128 + 4 * lane, p.en + o);
            if (RIG) mf_cp_async16(gs + r * 128 + 4 * lane, p.rig + o);
        }
    }
    float rmap[2 * D + 1];
#pragma unroll
    for (int j = 0; j <= 2 * D; ++j) rmap[j] = RIG ? p.rigmap[j - D] : 0.f;
    float mp[4] = {inf, inf, inf, inf};
    if (y0 > 0 && inmem) {
        const float4 v = *reinterpret_cast<const float4 *>(p.m + (size_t) (y0 - 1) * p.pitch + x0);
        mp[0] = v.x, mp[1] = v.y, mp[2] = v.z, mp[3] = v.w;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();

    unsigned go = (unsigned) y0 * (unsigned) p.pitch + (unsigned) x0;
    int r = 0;
    if (y0 == 0) { // row 0: m = en
        if (inmem) {
            const float4 e = *reinterpret_cast<const float4 *>(es + 4 * lane);
            mp[0] = e.x, mp[1] = e.y, mp[2] = e.z, mp[3] = e.w;
            if (interior) *reinterpret_cast<float4 *>(p.m + go) = e;
        }
        go += p.pitch;
        r = 1;
    }
#pragma unroll 4
    for (; r < rows; ++r, go += p.pitch) {
        float4 e4 = make_float4(inf, inf, inf, inf), g4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (inmem) {
            e4 = *reinterpret_cast<const float4 *>(es + r * 128 + 4 * lane);
            if (RIG) g4 = *reinterpret_cast<const float4 *>(gs + r * 128 + 4 * lane);
        }
        const float en[4] = {e4.x, e4.y, e4.z, e4.w};
        const float rf[4] = {g4.x, g4.y, g4.z, g4.w};
        float v[4 + 2 * D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const float l = __shfl_up_sync(full, mp[4 - D + j], 1);
            v[j] = x0 <= 0 ? inf : l; // columns < 0 do not exist; columns >= w hold +inf (sentinels)
            v[4 + D + j] = __shfl_down_sync(full, mp[j], 1);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) v[D + i] = mp[i];
        unsigned pk = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float cand[2 * D + 1];
#pragma unroll
            for (int j = 0; j <= 2 * D; ++j) cand[j] = RIG ? __fadd_rn(v[i + j], __fmul_rn(rf[i], rmap[j])) : v[i + j];
            float best = cand[0];
#pragma unroll
            for (int j = 1; j <= 2 * D; ++j) best = fminf(best, cand[j]);
            unsigned b = ((unsigned) ((LR ? -D : D) & 0xff)) << (8 * i);
#pragma unroll
            for (int j = 1; j <= 2 * D; ++j) {
                const int jj = LR ? j : 2 * D - j;
                if (cand[jj] == best) b = ((unsigned) ((jj - D) & 0xff)) << (8 * i);
            }
            pk |= b;
            mp[i] = __fadd_rn(en[i], best);
        }
        if (interior) {
            *reinterpret_cast<float4 *>(p.m + go) = make_float4(mp[0], mp[1], mp[2], mp[3]);
            *reinterpret_cast<unsigned *>(p.pdx + go) = pk;
        }
    }
}

} // namespace b200c
