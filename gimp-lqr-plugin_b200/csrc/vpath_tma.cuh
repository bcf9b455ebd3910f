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
#define VT_NT 3       // tiles: the chunk being chased and the next two being staged
#define VT_TILE 13568 // words per tile
#define VT_MAXR 64    // rows per chunk, at most 2 per DMA lane

static constexpr size_t vt_smem_bytes()
{
    return sizeof(int) * ((size_t) VT_NT * VT_TILE + VT_NT * (VT_MAXR + 1) * 4 + VT_NT * 8 + 8 + 64 + 4 * (VT_MAXR + 1));
}

template <bool D1>
__global__ void __launch_bounds__(VT_THREADS, 1) k_vpath_tma(DevP p)
{
    extern __shared__ __align__(128) unsigned char vt_smem[];
    int *tiles = reinterpret_cast<int *>(vt_smem);              // [NT][VT_TILE]
    int *rtab = tiles + VT_NT * VT_TILE;                        // [NT][VT_MAXR+1][4] rawadd, ladd, -, -
    int *cdesc = rtab + VT_NT * (VT_MAXR + 1) * 4;              // [NT][8] top, ntrans, lo, hi, cx (chaser -> DMA), last
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(cdesc + VT_NT * 8); // [NT] (+pad)
    float *s_v = reinterpret_cast<float *>(cdesc + VT_NT * 8 + 8); // [32]
    int *s_x = reinterpret_cast<int *>(s_v + 32);               // [32]
    int *outbuf = s_x + 32;                                     // [2][2][VT_MAXR+1] vpath / vpath_x of a chunk

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = VT_THREADS / 32;
    const int h = p.h, w = p.w, D = p.delta_x;

    if (tid == 0) {
        for (int i = 0; i < VT_NT; ++i) ut_mbar_init(&mbar[i], 1);
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

    // chunk geometry: R rows per chunk; a chunk's window is centred on the seam position TWO chunks earlier (it is
    // staged while the chaser is still two chunks away), so its half-width is 3*R*delta_x
    int R = VT_MAXR - 1;
    while (R > 1 && (R + 1) * (2 * (6 * R * D + 1) + 24) > VT_TILE) --R;
    const int HW = 3 * R * D;

    if (warp == 1) {
        // =============================================================================== DMA warp
        // stage the chunk with top row `top`, centred on column cx, into tile c % NT; returns its transitions
        auto stage = [&](int c, int top, int cx) -> int {
            const int t = c % VT_NT;
            int *tile = tiles + t * VT_TILE;
            int *rt = rtab + t * (VT_MAXR + 1) * 4;
            const int lo = max(cx - HW, 0), hi = min(cx + HW, w - 1);
            const int cw = hi - lo + 1;
            const int nrows = top >= 0 ? min(R, top) + 1 : 0; // raw rows top, top-1, ... (transitions: nrows-1)
            int need[2], nraw[2], nsp[2], rawbase[2], zbase[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int r = lane + 32 * q;
                need[q] = 0, nraw[q] = 0, nsp[q] = 0, rawbase[q] = 0, zbase[q] = 0;
                if (r < nrows) {
                    const int y = top - r;
                    const long long rawpos = (long long) y * p.raw_stride + lo;
                    rawbase[q] = (int) (rawpos & ~3LL);
                    nraw[q] = (int) ((rawpos + cw - rawbase[q] + 3) & ~3LL);
                    if (y > 0) { // the row's parents may be looked up: stage its span of `least`
                        const int *rr = p.raw + (size_t) y * p.raw_stride;
                        const int zlo = rr[lo], zhi = rr[hi];
                        zbase[q] = zlo & ~3;
                        nsp[q] = (zhi - zbase[q] + 1 + 3) & ~3;
                    }
                    need[q] = nraw[q] + nsp[q];
                }
            }
            // exclusive prefix over rows in row order: rows 0..31 (q = 0), then rows 32..63 (q = 1)
            int inc = need[0];
#pragma unroll
            for (int s2 = 1; s2 < 32; s2 <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, s2);
                if (lane >= s2) inc += v;
            }
            const int sum0 = __shfl_sync(0xffffffffu, inc, 31);
            int inc1 = need[1];
#pragma unroll
            for (int s2 = 1; s2 < 32; s2 <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc1, s2);
                if (lane >= s2) inc1 += v;
            }
            const int off[2] = {inc - need[0], sum0 + inc1 - need[1]};
            // rows that fit the tile form a prefix
            const unsigned fit0 = __ballot_sync(0xffffffffu, lane < nrows && off[0] + need[0] <= VT_TILE);
            const unsigned fit1 = __ballot_sync(0xffffffffu, lane + 32 < nrows && off[1] + need[1] <= VT_TILE);
            int nfit = __popc(fit0) + __popc(fit1);
            if (nfit < 2) nfit = 0; // a chunk needs at least one transition (two rows); otherwise nothing is staged
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
            for (int s2 = 16; s2 > 0; s2 >>= 1) total += __shfl_xor_sync(0xffffffffu, total, s2);
            const int ntr = max(nfit - 1, 0);
            if (lane == 0) {
                cdesc[t * 8 + 0] = top;
                cdesc[t * 8 + 1] = ntr; // transitions available in this chunk
                cdesc[t * 8 + 2] = lo;
                cdesc[t * 8 + 3] = hi;
                if (nfit > 0) ut_mbar_expect(&mbar[t], total);
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
            return ntr;
        };
        // tops of the chunks are fixed by the number of transitions each one could stage
        int top_next = h - 1; // top row of the next chunk to stage
        const int cx0 = cdesc[4];
        int n0 = stage(0, top_next, cx0);
        top_next -= n0;
        int n1 = 0;
        if (n0 > 0 && top_next > 0) {
            n1 = stage(1, top_next, cx0); // at most 2R rows away from the arg-min: inside the half-width
            top_next -= n1;
        } else if (lane == 0) {
            cdesc[1 * 8 + 1] = 0;
        }
        bool more = n0 > 0 && n1 > 0 && top_next > 0;
        for (int c = 0;; ++c) {
            asm volatile("bar.sync 2, 64;" ::: "memory"); // the chaser published the seam column at the top of chunk c
            // the chaser stops after the first chunk that has no transitions or reaches row 0
            const int ntr_c = cdesc[(c % VT_NT) * 8 + 1], top_c = cdesc[(c % VT_NT) * 8 + 0];
            if (ntr_c == 0 || top_c - ntr_c <= 0) break;
            const int cx = cdesc[(c % VT_NT) * 8 + 4];
            if (more) {
                const int n = stage(c + 2, top_next, cx); // tile (c+2) % NT held chunk c-1, which the chaser has left
                top_next -= n;
                more = n > 0 && top_next > 0;
            } else if (lane == 0) {
                cdesc[((c + 2) % VT_NT) * 8 + 1] = 0; // nothing beyond: the chaser will see an empty chunk
                cdesc[((c + 2) % VT_NT) * 8 + 0] = top_next;
            }
        }
    } else {
        // =============================================================================== CHASER
        int y = h - 1;
        bool found = true; // the last parent was located within delta_x (always, in practice)
        int x = cdesc[4], last = cdesc[5];
        for (int c = 0;; ++c) {
            const int t = c % VT_NT;
            if (c > 0 && lane == 0) cdesc[t * 8 + 4] = x; // seam column at the top of chunk c: centre of chunk c+2
            asm volatile("bar.sync 2, 64;" ::: "memory");
            const int ntr = cdesc[t * 8 + 1], wlo = cdesc[t * 8 + 2], whi = cdesc[t * 8 + 3];
            if (ntr == 0) break;
            if (!ut_mbar_wait(&mbar[t], (unsigned) ((c / VT_NT) & 1))) atomicOr(p.err, 16);
            const int *tile = tiles + t * VT_TILE;
            const int *rt = rtab + t * (VT_MAXR + 1) * 4;
            int *ob = outbuf + (c & 1) * 2 * (VT_MAXR + 1);
            const int y_top = y;
            if (lane == 0) {
                int rawadd = rt[0], ladd = rt[1];
                for (int r = 0; r < ntr; ++r, --y) {
                    ob[r] = last;
                    ob[VT_MAXR + 1 + r] = x;
                    const int rawadd_up = rt[r * 4 + 4], ladd_up = rt[r * 4 + 5];
                    const int zc = found ? last : tile[rawadd + x]; // liblqr: least[raw[y][last_x]]
                    const int l = tile[ladd + zc];
                    int nx = -1;
                    if (D1) {
                        // candidates x-1, x, x+1 (first match in that order; ids are unique, so at most one matches)
                        const int xl = max(x - 1, wlo), xr = min(x + 1, whi);
                        const int al = tile[rawadd_up + xl], a0 = tile[rawadd_up + x], ar = tile[rawadd_up + xr];
                        nx = (ar == l && x + 1 <= min(w - 1, whi)) ? x + 1 : nx;
                        nx = a0 == l ? x : nx;
                        nx = (al == l && x - 1 >= max(0, wlo)) ? x - 1 : nx;
                    } else {
                        const int x_lo = max(max(x - D, 0), wlo), x_hi = min(min(x + D, w - 1), whi);
                        for (int xx = x_hi; xx >= x_lo; --xx)
                            if (tile[rawadd_up + xx] == l) nx = xx;
                    }
                    found = nx >= 0;
                    x = found ? nx : x;
                    last = l;
                    rawadd = rawadd_up;
                    ladd = ladd_up;
                }
                if (x < wlo || x > whi) atomicOr(p.err, 2); // cannot happen: the window covers the seam's reach
            }
            y = __shfl_sync(0xffffffffu, y, 0);
            x = __shfl_sync(0xffffffffu, x, 0);
            last = __shfl_sync(0xffffffffu, last, 0);
            found = __shfl_sync(0xffffffffu, (int) found, 0) != 0;
            __syncwarp();
            // flush the chunk's seam (rows y_top, y_top-1, ...) coalesced
            for (int r = lane; r < ntr; r += 32) {
                p.vpath[y_top - r] = ob[r];
                p.vpath_x[y_top - r] = ob[VT_MAXR + 1 + r];
            }
            if (y <= 0) break;
        }
        if (lane == 0) {
            if (y > 0) atomicOr(p.err, 32); // staging was impossible before row 0 was reached
            if (y >= 0) {
                p.vpath[y] = last;
                p.vpath_x[y] = x;
            }
        }
    }
}

} // namespace b200c
