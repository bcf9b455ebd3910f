// band_tail.cuh -- K2c, the rows of the incremental m-map DP (A.8) that the one-CTA band kernel cannot tile.
//
// k_band_dp hands over at the first row whose window is wider than its 12 segments (deep rows of large images with
// delta_x >= 2: the band widens by up to delta_x columns per row).  From there on the band is a large part of the
// row, so this kernel evaluates EVERY column of the remaining rows -- any superset of liblqr's band is exact
// (DESIGN.md section 5): a cell whose parents' values did not change re-evaluates to "keep" -- with all SMs: column
// strips of 128 columns per warp, 4 cells per lane, the row in registers, the per-cell rule of the band kernel's
// settle path (candidates, arg-min, keep-old test).  Strips overlap by rows * delta_x columns (trapezoid), so the
// strips of one block of bt_rows() rows are independent; between row blocks the CTAs meet at a grid barrier
// (cooperative launch).  The operands of a row block (en, old m, old parents, rigidity mask) do not depend on the
// chain and are fetched up front with cp.async.
#pragma once
#include <cooperative_groups.h>

#include "band_dp.cuh"

namespace b200c {

#define BT_WARPS 4
#define BT_THREADS (BT_WARPS * 32)

// rows per grid barrier: more rows amortise the barrier, but the strips overlap by rows * delta_x columns on each side
__host__ __device__ constexpr int bt_rows(int delta_x) { return delta_x <= 2 ? 16 : 8; }
__host__ __device__ constexpr int bt_hk(int delta_x) { return (bt_rows(delta_x) * delta_x + 3) & ~3; }
__host__ __device__ constexpr int bt_strip(int delta_x) { return 128 - 2 * bt_hk(delta_x); }
static inline size_t bt_smem_bytes(int delta_x, bool rig) { return (size_t) BT_WARPS * bt_rows(delta_x) * 128 * (4 + 4 + 1 + (rig ? 4 : 0)); }
// CTAs that cover the widest row of a session (width w plus the sentinel columns a parent scan can reach)
static inline int bt_grid(int w, int delta_x) { return (((w + 4 + 3) & ~3) + BT_WARPS * bt_strip(delta_x) - 1) / (BT_WARPS * bt_strip(delta_x)); }

__device__ __forceinline__ void bt_cp_async(void *dst_smem, const void *src, int bytes16)
{
    if (bytes16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned) __cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned) __cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}

// p.tail = {first row left, ...} written by k_band_dp of the same seam; p.tail[0] >= h: nothing to do
template <int D, bool RIG, bool LR>
__global__ void __launch_bounds__(BT_THREADS) k_band_tail(const DevP pin0, const DevP *tab)
{
    pdl_entry();
    const DevP pin = pick_image(pin0, tab);
    const DevP p = seam_view(pin, 1);
    const int y_from = *reinterpret_cast<volatile int *>(p.tail);
    if (y_from >= p.h) return; // the same for every thread of the grid
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    extern __shared__ __align__(16) unsigned char bt_smem[];
    constexpr int R = bt_rows(D), HK = bt_hk(D), S = bt_strip(D);
    static_assert(S >= 32, "strips must keep an interior");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int strip = blockIdx.x * BT_WARPS + warp;
    const int x0 = strip * S - HK + 4 * lane; // first of this lane's 4 columns
    const int wlim = min((p.w + 4 + 3) & ~3, p.pitch);
    const bool active = strip * S < wlim;      // warp-uniform; idle warps only keep the barrier count
    const bool inmem = x0 >= 0 && x0 < p.pitch; // pitch is a multiple of 4
    const bool interior = active && 4 * lane >= HK && 4 * lane < 128 - HK && inmem && x0 < wlim;
    const float inf = __int_as_float(0x7f800000);
    const unsigned full = 0xffffffffu;

    constexpr int per_warp = R * 128 * (4 + 4 + 1 + (RIG ? 4 : 0));
    unsigned char *base = bt_smem + (size_t) warp * per_warp;
    float *es = reinterpret_cast<float *>(base);                 // en   [R][128]
    float *os = es + R * 128;                                    // old m
    float *gs = os + R * 128;                                    // rigidity mask (RIG)
    unsigned char *ps = base + (size_t) R * 128 * (RIG ? 12 : 8); // old parents [R][128] bytes
    float rmap[2 * D + 1];
#pragma unroll
    for (int j = 0; j <= 2 * D; ++j) rmap[j] = RIG ? p.rigmap[j - D] : 0.f;
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.cells) atomicAdd(p.cells, (unsigned long long) (p.h - y_from) * (unsigned long long) p.w);

    for (int yb = y_from; yb < p.h; yb += R) {
        const int rows = min(R, p.h - yb);
        if (active) {
            if (inmem) {
                for (int r = 0; r < rows; ++r) {
                    const size_t o = (size_t) (yb + r) * p.pitch + x0;
                    bt_cp_async(es + r * 128 + 4 * lane, p.en + o, 1);
                    bt_cp_async(os + r * 128 + 4 * lane, p.m + o, 1);
                    if (RIG) bt_cp_async(gs + r * 128 + 4 * lane, p.rig + o, 1);
                    bt_cp_async(ps + r * 128 + 4 * lane, p.pdx + o, 0);
                }
            }
            float mp[4] = {inf, inf, inf, inf}; // row yb-1: written by other SMs in the previous block -> past L1
            if (yb > 0 && inmem) {
                const float4 v = __ldcg(reinterpret_cast<const float4 *>(p.m + (size_t) (yb - 1) * p.pitch + x0));
                mp[0] = v.x, mp[1] = v.y, mp[2] = v.z, mp[3] = v.w;
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            unsigned go = (unsigned) yb * (unsigned) p.pitch + (unsigned) x0;
            for (int r = 0; r < rows; ++r, go += p.pitch) {
                float4 e4 = make_float4(inf, inf, inf, inf), o4 = e4, g4 = make_float4(1.f, 1.f, 1.f, 1.f);
                unsigned pw = 0;
                if (inmem) {
                    e4 = *reinterpret_cast<const float4 *>(es + r * 128 + 4 * lane);
                    o4 = *reinterpret_cast<const float4 *>(os + r * 128 + 4 * lane);
                    if (RIG) g4 = *reinterpret_cast<const float4 *>(gs + r * 128 + 4 * lane);
                    pw = *reinterpret_cast<const unsigned *>(ps + r * 128 + 4 * lane);
                }
                const float en[4] = {e4.x, e4.y, e4.z, e4.w};
                const float mo[4] = {o4.x, o4.y, o4.z, o4.w};
                const float rf[4] = {g4.x, g4.y, g4.z, g4.w};
                float nv[4];
                unsigned pk = 0;
                if (yb + r == 0) { // row 0: m = en (true of every cell of the row); parents are not defined there
#pragma unroll
                    for (int i = 0; i < 4; ++i) nv[i] = en[i];
                    pk = pw;
                } else {
                    float v[4 + 2 * D];
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        const float l = __shfl_up_sync(full, mp[4 - D + j], 1);
                        v[j] = x0 <= 0 ? inf : l; // columns < 0 do not exist; columns >= w hold +inf (sentinels)
                        v[4 + D + j] = __shfl_down_sync(full, mp[j], 1);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[D + i] = mp[i];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float cand[2 * D + 1];
                        float best = inf;
#pragma unroll
                        for (int j = 0; j <= 2 * D; ++j) {
                            cand[j] = RIG ? __fadd_rn(v[i + j], __fmul_rn(rf[i], rmap[j])) : v[i + j];
                            best = fminf(best, cand[j]);
                        }
                        const int bdx = bd_argmin<D, LR>(cand, best);
                        const float nm = __fadd_rn(en[i], best);
                        const int pold = (int) (signed char) (pw >> (8 * i));
                        nv[i] = keep_old(pold, bdx, mo[i], nm) ? mo[i] : nm;
                        pk |= ((unsigned) (bdx & 0xff)) << (8 * i);
                    }
                }
                if (interior) {
                    *reinterpret_cast<float4 *>(p.m + go) = make_float4(nv[0], nv[1], nv[2], nv[3]);
                    *reinterpret_cast<unsigned *>(p.pdx + go) = pk;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) mp[i] = nv[i];
            }
        }
        grid.sync(); // row block done everywhere (and visible) before anyone reads its last row
    }
}

} // namespace b200c
