// vpath_tma.cuh -- K3, the seam backtrack (liblqr lqr_carver_build_vpath), staged with TMA bulk
// copies.  After the block-wide arg-min over the last row, two warps finish the job:
//
//   * the CHASER (one thread) walks up the rows entirely in shared memory.  Per row it needs the parent id of
//     the cell it stands on (least[id], one dependent shared-memory load) and the column of that id in the row
//     above (2*delta_x+1 independent loads from the staged raw window, issued in parallel with the first);
//   * the DMA warp stages the next chunk of rows while the chaser works on the current one: for every row one
//     cp.async.bulk load of the raw-id window around the seam and one of the physical span of `least` those ids
//     cover (ids along a row are increasing and nearly contiguous).  The window of chunk c+1 is centred on the
//     seam position at the top of chunk c and is twice as wide as the seam can travel, so it is known one chunk
//     ahead.  No per-cell gather or resolve pass exists.
//
// Exactly liblqr's semantics, including the corner where a parent is not found within delta_x (the column then
// stays and the next parent is looked up through raw[y][x], not through the parent id).
#pragma once
#include "carver_kernels.cuh"
#include "mmap_update_tma.cuh"

namespace b200c {

#define VT_THREADS 256
#define VT_TILE 20480 // words per tile (2 tiles)
#define VT_MAXR 64    // rows (transitions) per chunk, at most 2 per DMA lane

static constexpr size_t vt_smem_bytes() { return sizeof(int) * ((size_t) 2 * VT_TILE + 2 * (VT_MAXR + 1) * 4 + 16 + 8 + 64); }

__global__ void __launch_bounds__(VT_THREADS, 1) k_vpath_tma(DevP p)
{
    extern __shared__ __align__(128) unsigned char vt_smem[];
    int *tiles = reinterpret_cast<int *>(vt_smem);              // [2][VT_TILE]
    int *rtab = tiles + 2 * VT_TILE;                            // [2][VT_MAXR+1][4] rawadd, ladd, -, -
    int *cdesc = rtab + 2 * (VT_MAXR + 1) * 4;                  // [2][8] top, ntrans, -, -, cx (chaser -> DMA), last
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(cdesc + 16); // [2] (+pad)
    float *s_v = reinterpret_cast<float *>(cdesc + 16 + 8);     // [32]
    int *s_x = reinterpret_cast<int *>(s_v + 32);               // [32]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = VT_THREADS / 32;
    const int h = p.h, w = p.w, D = p.delta_x;

    if (tid == 0) {
        ut_mbar_init(&mbar[0], 1);
        ut_mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- arg-min over the last row (A.6 tie rule)
    const int *row = p.raw + (size_t) (h - 1) * p.raw_stride;
    float best = 536870912.f;
    int bx = -1;
    for (int x = tid; x < w; x += VT_THREADS) {
        const float v = p.m[row[x]];
        if (v < best || (v == best && p.leftright == 1)) {
            best = v;
            bx = x;
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_down_sync(0xffffffffu, best, off);
        const int ox = __shfl_down_sync(0xffffffffu, bx, off);
        if (seam_better(ov, ox, best, bx, p.leftright)) {
            best = ov;
            bx = ox;
        }
    }
    if (lane == 0) {
        s_v[warp] = best;
        s_x[warp] = bx;
    }
    __syncthreads();
    if (tid == 0) {
        for (int i = 1; i < nwarp; ++i)
            if (seam_better(s_v[i], s_x[i], best, bx, p.leftright)) {
                best = s_v[i];
                bx = s_x[i];
            }
        const int last_x = bx < 0 ? 0 : bx;
        cdesc[4] = last_x;      // seam column at the top row of chunk 0
        cdesc[5] = row[last_x]; // its pixel id
    }
    __syncthreads();
    if (warp >= 2) return;

    // chunk geometry: R transitions per chunk, window half-width 2*R*delta_x (the seam moves <= delta_x per row and
    // the window of a chunk is centred on the seam position one chunk earlier)
    int R = VT_MAXR - 1;
    while (R > 1 && (R + 1) * (2 * (4 * R * D + 1) + 24) > VT_TILE) --R;
    const int HW = 2 * R * D;

    if (warp == 1) {
        // =============================================================================== DMA warp
        // stage the chunk with top row `top` centred on column cx into tile t; returns the transitions staged
        auto stage = [&](int c, int top, int cx) {
            const int t = c & 1;
            int *tile = tiles + t * VT_TILE;
            int *rt = rtab + t * (VT_MAXR + 1) * 4;
            const int lo = max(cx - HW, 0), hi = min(cx + HW, w - 1);
            const int cw = hi - lo + 1;
            const int nrows = min(R, top) + 1; // raw rows top, top-1, ..., top-nrows+1 (transitions: nrows-1)
            int need[2], nraw[2], nsp[2], rawbase[2], zbase[2];
            int tot = 0;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int r = lane + 32 * q;
                need[q] = 0, nraw[q] = 0, nsp[q] = 0, rawbase[q] = 0, zbase[q] = 0;
                if (r < nrows) {
                    const int y = top - r;
                    const long long rawpos = (long long) y * p.raw_stride + lo;
                    rawbase[q] = (int) (rawpos & ~3LL);
                    nraw[q] = (int) ((rawpos + cw - rawbase[q] + 3) & ~3LL);
                    if (r < nrows - 1 && y > 0) { // this row's parents are looked up: stage its span of `least`
                        const int *rr = p.raw + (size_t) y * p.raw_stride;
                        const int zlo = rr[lo], zhi = rr[hi];
                        zbase[q] = zlo & ~3;
                        nsp[q] = (zhi - zbase[q] + 1 + 3) & ~3;
                    }
                    need[q] = nraw[q] + nsp[q];
                }
                tot += need[q];
            }
            // exclusive prefix over rows in row order: lane-major for q = 0, then q = 1
            int inc = need[0];
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, s);
                if (lane >= s) inc += v;
            }
            const int sum0 = __shfl_sync(0xffffffffu, inc, 31);
            int inc1 = need[1];
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc1, s);
                if (lane >= s) inc1 += v;
            }
            int off[2] = {inc - need[0], sum0 + inc1 - need[1]};
            // rows that fit the tile form a prefix
            const unsigned fit0 = __ballot_sync(0xffffffffu, lane < nrows && off[0] + need[0] <= VT_TILE);
            const unsigned fit1 = __ballot_sync(0xffffffffu, lane + 32 < nrows && off[1] + need[1] <= VT_TILE);
            const int nfit = __popc(fit0) + __popc(fit1);
            unsigned bytes = 0;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int r = lane + 32 * q;
                if (r < nfit) {
                    const int y = top - r;
                    rt[r * 4 + 0] = off[q] + (int) ((long long) y * p.raw_stride - rawbase[q]); // raw[y][x] = tile[rawadd + x]
                    rt[r * 4 + 1] = off[q] + nraw[q] - zbase[q];                                  // least[z]  = tile[ladd + z]
                    bytes += (unsigned) (nraw[q] + nsp[q]) * 4u;
                }
            }
            unsigned total = bytes;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) total += __shfl_xor_sync(0xffffffffu, total, s);
            if (lane == 0) {
                cdesc[t * 8 + 0] = top;
                cdesc[t * 8 + 1] = max(nfit - 1, 0); // transitions available in this chunk
                cdesc[t * 8 + 2] = lo;
                cdesc[t * 8 + 3] = hi;
                ut_mbar_expect(&mbar[t], total);
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int r = lane + 32 * q;
                if (r < nfit) {
                    if (nraw[q] > 0) ut_bulk_load(tile + off[q], p.raw + rawbase[q], (unsigned) nraw[q] * 4u, &mbar[t]);
                    if (nsp[q] > 0)
                        ut_bulk_load(tile + off[q] + nraw[q], p.least + zbase[q], (unsigned) nsp[q] * 4u, &mbar[t]);
                }
            }
            (void) tot;
        };
        int top = h - 1;
        stage(0, top, cdesc[4]);
        for (int c = 0;; ++c) {
            asm volatile("bar.sync 2, 64;" ::: "memory"); // the chaser has published the seam column at `top` of chunk c
            const int ntr = cdesc[(c & 1) * 8 + 1];
            if (ntr == 0 || top - ntr <= 0) break; // chunk c reaches row 0 (or nothing could be staged)
            const int cx = cdesc[(c & 1) * 8 + 4];
            top -= ntr;
            stage(c + 1, top, cx); // tile (c+1)&1 held chunk c-1, which the chaser has left
        }
    } else {
        // =============================================================================== CHASER
        int y = h - 1;
        bool ok = true, found = true; // found: the last parent was located within delta_x (always, in practice)
        int x = 0, last = 0;
        for (int c = 0; ok; ++c) {
            const int t = c & 1;
            if (c == 0) {
                x = cdesc[4];
                last = cdesc[5];
            } else if (lane == 0) {
                cdesc[t * 8 + 4] = x; // seam column at the top row of chunk c: the DMA warp centres chunk c+1 on it
            }
            asm volatile("bar.sync 2, 64;" ::: "memory");
            if (!ut_mbar_wait(&mbar[t], (unsigned) ((c >> 1) & 1))) atomicOr(p.err, 16);
            const int ntr = cdesc[t * 8 + 1], wlo = cdesc[t * 8 + 2], whi = cdesc[t * 8 + 3];
            const int *tile = tiles + t * VT_TILE;
            const int *rt = rtab + t * (VT_MAXR + 1) * 4;
            if (lane == 0) {
                for (int r = 0; r < ntr; ++r, --y) {
                    p.vpath[y] = last;
                    p.vpath_x[y] = x;
                    const int rawadd = rt[r * 4 + 0], ladd = rt[r * 4 + 1], rawadd_up = rt[r * 4 + 4];
                    const int zc = found ? last : tile[rawadd + x]; // liblqr: least[raw[y][last_x]]
                    const int l = tile[ladd + zc];
                    const int x_lo = max(max(x - D, 0), wlo), x_hi = min(min(x + D, w - 1), whi);
                    int nx = -1;
                    for (int xx = x_lo; xx <= x_hi; ++xx)
                        if (nx < 0 && tile[rawadd_up + xx] == l) nx = xx;
                    found = nx >= 0;
                    if (found) x = nx;
                    last = l;
                }
                if (x < wlo || x > whi) atomicOr(p.err, 2); // cannot happen: the window is twice the seam's reach
            }
            y = __shfl_sync(0xffffffffu, y, 0);
            x = __shfl_sync(0xffffffffu, x, 0);
            last = __shfl_sync(0xffffffffu, last, 0);
            if (ntr == 0 || y <= 0) ok = false;
        }
        if (lane == 0 && y >= 0) {
            // rows the chunks did not reach (only row 0 in the regular case; more if staging was impossible, which the
            // error word reports through a non-zero remainder)
            if (y > 0) atomicOr(p.err, 32);
            p.vpath[y] = last;
            p.vpath_x[y] = x;
        }
    }
}

} // namespace b200c
