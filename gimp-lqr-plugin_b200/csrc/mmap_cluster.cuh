// mmap_cluster.cuh -- K2, the full m-map DP (liblqr lqr_carver_build_mmap) as ONE launch per pass: a
// thread-block CLUSTER per image, halo exchange through distributed shared memory.
//
// m[y][x] = en[y][x] + min over |dx| <= delta_x of m[y-1][x+dx] (+ rigidity term) is a chain of h dependent rows.  The
// width is cut into SEGMENTS of 128 columns, one warp each, 4 consecutive cells per lane, the row in registers: a row
// step is two shuffles plus 4 cells of 3-input min / add -- the VALUES only; the parent offsets are not needed by the
// chain and are recomputed from the finished map by k_parents_full, fully parallel.  Segments overlap by MC_HK = 32
// columns on each side (trapezoid tiling: an edge goes stale by delta_x columns per row), so a warp runs MC_K = 32 /
// delta_x rows without talking to anybody; its interior is the 64 columns in the middle.  Every MC_K rows the interiors'
// outer 32 columns are written into the NEIGHBOUR segment's halo buffer -- plain shared memory inside a CTA,
// st.shared::cluster across CTAs -- followed by one cluster barrier.  The segment geometry never changes, so interior
// lanes keep their row in registers for the whole pass and only the 16 halo lanes reload.  Energy rows are streamed
// through a per-warp ring in shared memory with cp.async, MC_AHEAD rows ahead of the chain.
//
// Grid: (cluster size, 1, images); cluster size 1, 2, 4, 8 or 16 (non-portable) CTAs of MC_WARPS..8 warps.
#pragma once
#include <cooperative_groups.h>

#include "band_dp.cuh"

namespace b200c {

#define MC_HK 32     // halo columns per side
#define MC_S 64      // interior columns per segment
#define MC_RING 16   // energy rows per warp in shared memory
#define MC_AHEAD 12  // rows fetched ahead of the chain

__host__ __device__ constexpr int mc_rows(int delta_x) { return delta_x <= 1 ? 32 : (delta_x == 2 ? 16 : 8); } // rows * delta_x <= MC_HK
__host__ __device__ constexpr int mc_batch(int delta_x) { return mc_rows(delta_x) < 16 ? mc_rows(delta_x) : 16; } // rows per fetch
// bytes of shared memory per warp: energy ring (+ rigidity-mask ring), two parities of two halo buffers
// (without a mask: two buffers of a batch of at most 16 rows; with one: the two row rings)
__host__ __device__ constexpr int mc_warp_bytes(bool rig) { return (rig ? MC_RING * 512 * 2 : 2 * 16 * 512) + 2 * 2 * MC_HK * 4; }

__device__ __forceinline__ void mc_cp16(void *dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned) __cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
// address of `p` (shared memory of this CTA) in the shared memory of CTA `rank` of the cluster
__device__ __forceinline__ unsigned mc_mapa(const void *p, unsigned rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"((unsigned) __cvta_generic_to_shared(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void mc_st_cluster(unsigned addr, float4 v)
{
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mc_cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int D, bool RIG>
__global__ void k_mmap_full_cluster(const DevP p0, int nwarps, const DevP *tab)
{
    const DevP p = pick_image(p0, tab);
    extern __shared__ __align__(16) unsigned char mc_smem[];
    constexpr int K = mc_rows(D);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned crank = blockIdx.x, csize = gridDim.x; // one cluster per image along x
    const int g = (int) crank * nwarps + warp;            // segment index
    const int x0 = g * MC_S - MC_HK + 4 * lane;           // first of this lane's 4 columns
    const int wlim = min((p.w + 4 + 3) & ~3, p.pitch);    // the image plus the +inf sentinel columns a parent scan reaches
    const bool inmem = x0 >= 0 && x0 < p.pitch;
    const bool interior = 4 * lane >= MC_HK && 4 * lane < MC_HK + MC_S && inmem && x0 < wlim;
    const float inf = __int_as_float(0x7f800000);
    const unsigned full = 0xffffffffu;

    unsigned char *wb = mc_smem + (size_t) warp * mc_warp_bytes(RIG);
    float *es = reinterpret_cast<float *>(wb);                                  // [MC_RING][128]
    float *gs = es + MC_RING * 128;                                             // [MC_RING][128] (RIG)
    float *halo = reinterpret_cast<float *>(wb + (RIG ? MC_RING * 512 * 2 : 2 * 16 * 512)); // [2 parities][2 sides][MC_HK]
    // where this warp's interior edges go: the right halo of segment g-1, the left halo of segment g+1
    const int gl = g - 1, gr = g + 1;
    const bool has_l = gl >= 0, has_r = gr < (int) csize * nwarps;
    auto halo_of = [&](int seg, int parity, int side) -> unsigned {
        const unsigned char *wbs = mc_smem + (size_t) (seg % nwarps) * mc_warp_bytes(RIG);
        const float *h = reinterpret_cast<const float *>(wbs + (RIG ? MC_RING * 512 * 2 : 2 * 16 * 512)) + (parity * 2 + side) * MC_HK;
        return mc_mapa(h, (unsigned) (seg / nwarps));
    };

    float rmap[2 * D + 1];
#pragma unroll
    for (int j = 0; j <= 2 * D; ++j) rmap[j] = RIG ? p.rigmap[j - D] : 0.f;

    // one row of the chain from its operands; stores the interior
    auto row_step = [&](int y, const float4 e4, const float4 g4, float (&mp)[4]) {
        float nv[4];
        if (y == 0) { // row 0: m = en
            nv[0] = e4.x, nv[1] = e4.y, nv[2] = e4.z, nv[3] = e4.w;
        } else {
            float v[4 + 2 * D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                v[j] = __shfl_up_sync(full, mp[4 - D + j], 1); // columns < 0 belong to lanes that hold +inf for good
                v[4 + D + j] = __shfl_down_sync(full, mp[j], 1);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) v[D + i] = mp[i];
            const float en[4] = {e4.x, e4.y, e4.z, e4.w};
            const float rf[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float best = RIG ? __fadd_rn(v[i], __fmul_rn(rf[i], rmap[0])) : v[i];
#pragma unroll
                for (int j = 1; j <= 2 * D; ++j)
                    best = fminf(best, RIG ? __fadd_rn(v[i + j], __fmul_rn(rf[i], rmap[j])) : v[i + j]);
                nv[i] = __fadd_rn(en[i], best);
            }
        }
        if (interior) bd_st_global(p.m + (size_t) y * p.pitch + x0, nv);
#pragma unroll
        for (int i = 0; i < 4; ++i) mp[i] = nv[i];
    };
    // end of a trapezoid: hand the interior edges over, take the halos in
    auto exchange = [&](float (&mp)[4], int parity) {
        const float4 mine = make_float4(mp[0], mp[1], mp[2], mp[3]);
        const int c = 4 * lane - MC_HK; // column inside the interior, 0 .. 63
        if (c >= 0 && c < MC_HK && has_l) mc_st_cluster(halo_of(gl, parity, 1) + (unsigned) c * 4u, mine);
        if (c >= MC_S - MC_HK && c < MC_S && has_r) mc_st_cluster(halo_of(gr, parity, 0) + (unsigned) (c - (MC_S - MC_HK)) * 4u, mine);
        __syncwarp();
        mc_cluster_sync();
        if (4 * lane < MC_HK) { // left halo lanes
            if (has_l) {
                const float4 h = *reinterpret_cast<const float4 *>(halo + (parity * 2 + 0) * MC_HK + 4 * lane);
                mp[0] = h.x, mp[1] = h.y, mp[2] = h.z, mp[3] = h.w;
            }
        } else if (4 * lane >= MC_HK + MC_S) { // right halo lanes
            if (has_r) {
                const float4 h = *reinterpret_cast<const float4 *>(halo + (parity * 2 + 1) * MC_HK + (4 * lane - MC_HK - MC_S));
                mp[0] = h.x, mp[1] = h.y, mp[2] = h.z, mp[3] = h.w;
            }
        }
        if (!inmem || x0 >= wlim) mp[0] = mp[1] = mp[2] = mp[3] = inf; // outside the image for good
    };

    float mp[4] = {inf, inf, inf, inf};
    int parity = 0;
    // distributed shared memory may only be written once its CTA is known to run: every CTA of the cluster passes here
    // before the first exchange
    mc_cluster_sync();
    if constexpr (!RIG) {
        // the energy rows are fetched a BATCH (16 rows, 8 for delta_x >= 3) ahead: two buffers per warp
        constexpr int FB = mc_batch(D);
        float *buf = es; // [2][FB][128]
        auto fetch_batch = [&](int c) {
            const int y0 = c * FB;
            if (inmem)
                for (int r = 0; r < FB && y0 + r < p.h; ++r)
                    mc_cp16(buf + ((c & 1) * FB + r) * 128 + 4 * lane, p.en + (size_t) (y0 + r) * p.pitch + x0);
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        const int nbatch = (p.h + FB - 1) / FB;
        fetch_batch(0);
        for (int c = 0; c < nbatch; ++c) {
            if (c + 1 < nbatch) fetch_batch(c + 1);
            else asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();
            const int y0 = c * FB, rows = min(FB, p.h - y0);
            const float *eb = buf + (c & 1) * FB * 128 + 4 * lane;
#pragma unroll 8
            for (int r = 0; r < rows; ++r) {
                const float4 e4 = inmem ? *reinterpret_cast<const float4 *>(eb + r * 128) : make_float4(inf, inf, inf, inf);
                row_step(y0 + r, e4, make_float4(1.f, 1.f, 1.f, 1.f), mp);
            }
            if ((y0 + rows) % K == 0 && y0 + rows < p.h) { // a trapezoid ends with this batch
                exchange(mp, parity);
                parity ^= 1;
            }
        }
    } else {
        auto fetch = [&](int y) { // energy and rigidity-mask row y into the ring, asynchronously
            if (y < p.h && inmem) {
                const size_t o = (size_t) y * p.pitch + x0;
                mc_cp16(es + (y & (MC_RING - 1)) * 128 + 4 * lane, p.en + o);
                mc_cp16(gs + (y & (MC_RING - 1)) * 128 + 4 * lane, p.rig + o);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int y = 0; y < MC_AHEAD; ++y) fetch(y);
        for (int y = 0; y < p.h; ++y) {
            fetch(y + MC_AHEAD);
            asm volatile("cp.async.wait_group %0;" ::"n"(MC_AHEAD) : "memory");
            __syncwarp();
            float4 e4 = make_float4(inf, inf, inf, inf), g4 = make_float4(1.f, 1.f, 1.f, 1.f);
            if (inmem) {
                e4 = *reinterpret_cast<const float4 *>(es + (y & (MC_RING - 1)) * 128 + 4 * lane);
                g4 = *reinterpret_cast<const float4 *>(gs + (y & (MC_RING - 1)) * 128 + 4 * lane);
            }
            row_step(y, e4, g4, mp);
            if ((y + 1) % K == 0 && y + 1 < p.h) {
                exchange(mp, parity);
                parity ^= 1;
            }
        }
    }
    mc_cluster_sync(); // nobody leaves while a neighbour may still write into its halo buffers
}

// Parent offsets of the whole image from the finished m-map (A.5: scan left to right, strict '<' keeps the leftmost
// minimum, leftright turns ties to the right): one thread per 4 cells, packed store.
template <int D, bool RIG, bool LR>
__global__ void __launch_bounds__(256) k_parents_full(const DevP p0, const DevP *tab)
{
    const DevP p = pick_image(p0, tab);
    const int x0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    const int y = blockIdx.y + 1;
    if (y >= p.h || x0 >= p.w) return;
    const float inf = __int_as_float(0x7f800000);
    float rmap[2 * D + 1];
#pragma unroll
    for (int j = 0; j <= 2 * D; ++j) rmap[j] = RIG ? p.rigmap[j - D] : 0.f;
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

} // namespace b200c
