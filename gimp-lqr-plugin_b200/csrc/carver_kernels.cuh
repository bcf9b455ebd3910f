// carver_kernels.cuh -- sm_100a kernels of the seam-carving hot path.
//
// Each kernel states which step of the liblqr algorithm (cited A.n) it performs.
// Float discipline: every parity-critical float/double operation is written with the round-to-nearest
// intrinsics (__fadd_rn, __fmul_rn, __dadd_rn, ...) so that nvcc can never contract a*b+c into an FMA --
// the CPU reference path is compiled without FMA contraction and the seam choice depends on exact ties.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200c {

enum { READ_BRIGHTNESS = 0, READ_LUMA = 1 };
enum { GRAD_NORM = 0, GRAD_SUMABS = 1, GRAD_XABS = 2, GRAD_NULL = 3 };

// Parent offsets are stored as one signed byte per cell; PDX_NONE marks "the parent pixel is gone".
#define B200C_PDX_NONE (-128)
#define B200C_MAX_DELTA 120 // |parent offset| <= delta_x + 1 must fit the byte

// Device view of one carver, passed by value to every kernel.
//
// HBM layout.  Pixels, the visibility map, bias and rigidity mask are PHYSICAL
// (w0 x h0, never moved).  The seam-search state is COMPACT: row y of en / m / pdx / rig holds the values
// of the row's visible pixels in their CURRENT order, x in [0, w), at pitch `pitch` (a multiple of 16
// cells, so every row and every 16-cell window is 64-byte aligned for vector and TMA bulk access).  The
// carve kernel closes the gap in all of them, so DP rows are contiguous and need no indirection; `raw`
// (x-th visible pixel of row y -> physical id) is only used to reach pixels, bias and the visibility map.
// pdx[y][x] = x_parent - x in CURRENT coordinates: the carve kernel re-expresses it after every seam so it
// keeps naming the same parent PIXEL (liblqr stores the parent's pixel id in `least`), or PDX_NONE once
// that pixel has been carved away.
struct DevP {
    int w, h;          // current size (internal orientation)
    int w0, h0;        // allocated map size
    int w_start;       // reference width
    int raw_stride;    // pitch of the raw index table (== w_start)
    int pitch;         // pitch of the compact maps, in cells
    int channels, alpha;
    int level;
    int delta_x;
    int leftright;
    int grad_kind, read_kind, nrg_radius;
    int use_rig;       // rigidity != 0
    int raw_ident;     // the index table is the identity (raw[y][x] == y * w0 + x): pixels can be read without it
    int bd_maxseg;     // band DP: widest window (in segments) the tiled path takes (test knob B200C_BD_MAXSEG)
    const uint8_t *rgb;
    int *vs;
    int *raw;
    float *en;         // compact
    float *m;          // compact
    int8_t *pdx;       // compact
    signed char *jump; // [row blocks][pitch]: net column offset of a path across a block of rows (seam_trace.cuh)
    float *rig;        // compact copy of the rigidity mask (NULL without a mask)
    const float *bias;
    const float *rigmask;
    const float *rigmap; // centred: rigmap[dx], dx in [-delta_x, delta_x]
    int *vpath_x, *nrg_xmin, *nrg_xmax;
    int4 *fix;          // per band-DP chunk {y0, rows, elo, ehi}: the cells whose parents k_fix_parents recomputes
    int *fixn;          // number of entries of fix
    unsigned *nrg_pack; // per row: nrg_xmin | (nrg_xmax - nrg_xmin + 1) << 24, for the band DP's window planner
    unsigned long long *cells; // running count of band cells evaluated by the incremental DP
    long long *dbg;    // optional cycle counters (B200C_DBG=1), 32 slots
    int *dyn;          // seam counter of the current build session, or NULL: see seam_view()
    int *dyn_host;     // mapped host word that receives the number of COMPLETED seams of the session, or NULL
    int *far;          // count of rows whose FAR carve phase is done in this session (k_carve phase 2), or NULL
    int *tail;         // band DP -> tail kernel: {first row left to do (h: none), hull lo, hull hi} of this seam, or NULL
    int *err;          // device error word: bit 0 band left its staged window, bit 1 backtrack met a dead parent, bit 2 bulk copy timed out
};

// The kernels of the per-seam loop are launched with the SAME arguments for every seam of a session (which is what lets
// the host replay them as one CUDA graph): p.w is the width before the session's first seam, and the seam counter
// *p.dyn -- advanced by the backtrack kernel, the first of every iteration -- tells how many seams have gone since.
// post = 0: the width before this iteration's carve (backtrack); post = 1: after it (carve, energy band, DP).
// Every kernel of the seam search takes its carver's argument block by value AND an optional table of blocks in HBM: a
// batch session (b200c_batch_build_maps) advances many images with ONE launch per step, image = blockIdx.z.
__device__ __forceinline__ DevP pick_image(const DevP &p0, const DevP *tab) { return tab ? tab[blockIdx.z] : p0; }

// first thing of every iteration (one thread of the backtrack kernel): the seam counter moves on; every kernel of the
// previous iterations has finished by now (queue order), so the counter's value IS the number of completed seams --
// published to the host for the progress callbacks
// Programmatic dependent launch (the seam graph with B200C_PDL=1): first thing in every kernel of the per-seam loop --
// let the NEXT kernel of the chain be launched (its CTAs then wait here, resident), and wait until the PREVIOUS kernel
// has completed and its writes are visible.  Both are no-ops in an ordinary launch.
__device__ __forceinline__ void pdl_entry()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ void advance_seam(const DevP &p)
{
    const int i = ++*p.dyn;
    if (p.dyn_host) *reinterpret_cast<volatile int *>(p.dyn_host) = i;
}

__device__ __forceinline__ DevP seam_view(DevP p, int post, int *seam = nullptr)
{
    if (p.dyn) {
        const int i = *reinterpret_cast<volatile int *>(p.dyn);
        p.w -= i + post;
        if (seam) *seam = i;
    } else if (seam) {
        *seam = 0;
    }
    return p;
}

// ------------------------------------------------------------------------------------------------
// A.2 pixel reading: 8-bit channel / 255 in double; brightness = mean of colour channels, luma =
// Rec.709 weights; both multiplied by alpha when the image has an alpha channel.
// t255: optional table of (double) v / 255.0 for v in 0..255 (the same correctly rounded quotients, looked up instead
// of divided: a double division is a long software sequence on the GPU)
__device__ __forceinline__ double px_scalar(const DevP &p, int z, const double *t255 = nullptr)
{
    const uint8_t *q = p.rgb + (size_t) z * p.channels;
    auto n255 = [&](unsigned v) { return t255 ? t255[v] : (double) v / 255.0; };
    if (p.channels == 4) { // one 4-byte load per RGBA pixel; the arithmetic is the one below
        const uchar4 c = *reinterpret_cast<const uchar4 *>(q);
        const double r = n255(c.x), g = n255(c.y), b = n255(c.z);
        const double v4 = p.read_kind == READ_LUMA
                              ? __dadd_rn(__dadd_rn(__dmul_rn(0.2126, r), __dmul_rn(0.7152, g)), __dmul_rn(0.0722, b))
                              : __dadd_rn(__dadd_rn(r, g), b) / 3.0;
        return __dmul_rn(v4, n255(c.w));
    }
    double v;
    if (p.channels <= 2) {
        v = n255(q[0]);
    } else {
        const double r = n255(q[0]), g = n255(q[1]), b = n255(q[2]);
        if (p.read_kind == READ_LUMA)
            v = __dadd_rn(__dadd_rn(__dmul_rn(0.2126, r), __dmul_rn(0.7152, g)), __dmul_rn(0.0722, b));
        else
            v = __dadd_rn(__dadd_rn(r, g), b) / 3.0;
    }
    if (p.alpha >= 0) v = __dmul_rn(v, n255(q[p.alpha]));
    return v;
}

// A.3 energy of the pixel at current coordinates (x, y): central differences over the four nearest
// neighbours in the CURRENT image (through the raw index table), one-sided at the borders.
__device__ __forceinline__ float energy_at(const DevP &p, int x, int y, const double *t255 = nullptr)
{
    const int *row = p.raw + (size_t) y * p.raw_stride;
    const int z = row[x];
    float e = 0.f;
    if (p.grad_kind != GRAD_NULL) {
        const double b = px_scalar(p, z, t255);
        double gx, gy;
        if (y == 0)
            gy = __dsub_rn(p.h > 1 ? px_scalar(p, row[p.raw_stride + x], t255) : 0.0, b);
        else if (y < p.h - 1)
            gy = __dmul_rn(__dsub_rn(px_scalar(p, row[p.raw_stride + x], t255), px_scalar(p, row[x - p.raw_stride], t255)), 0.5);
        else
            gy = __dsub_rn(b, px_scalar(p, row[x - p.raw_stride], t255));
        if (x == 0)
            gx = __dsub_rn(p.w > 1 ? px_scalar(p, row[x + 1], t255) : 0.0, b);
        else if (x < p.w - 1)
            gx = __dmul_rn(__dsub_rn(px_scalar(p, row[x + 1], t255), px_scalar(p, row[x - 1], t255)), 0.5);
        else
            gx = __dsub_rn(b, px_scalar(p, row[x - 1], t255));
        if (p.grad_kind == GRAD_NORM)
            e = (float) sqrt(__dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy)));
        else if (p.grad_kind == GRAD_SUMABS)
            e = (float) __dmul_rn(__dadd_rn(fabs(gx), fabs(gy)), 0.5);
        else
            e = (float) fabs(gx);
    }
    float b_add = 0.f;
    if (p.bias) b_add = __fdiv_rn(p.bias[z], (float) p.w_start);
    return __fadd_rn(e, b_add);
}

// A.5 best parent of cell x of a row whose predecessor row holds the cumulative values `up` (compact, current
// coordinates): scan x+[-delta_x, delta_x] (clipped) left to right, strict '<' keeps the leftmost minimum,
// leftright==1 turns ties to the right.  Returns the candidate value; bdx = x_parent - x.  rf = the cell's
// rigidity-mask factor (1 without a mask).
__device__ __forceinline__ float best_parent(const DevP &p, const float *up, int x, float rf, int &bdx)
{
    const int x1_min = max(-x, -p.delta_x);
    const int x1_max = min(p.w - 1 - x, p.delta_x);
    int least = x1_min;
    float best;
    if (p.use_rig) {
        best = __fadd_rn(up[x + x1_min], __fmul_rn(rf, p.rigmap[x1_min]));
        for (int x1 = x1_min + 1; x1 <= x1_max; ++x1) {
            const float cand = __fadd_rn(up[x + x1], __fmul_rn(rf, p.rigmap[x1]));
            if (cand < best || (cand == best && p.leftright == 1)) {
                best = cand;
                least = x1;
            }
        }
    } else {
        best = up[x + x1_min];
        for (int x1 = x1_min + 1; x1 <= x1_max; ++x1) {
            const float cand = up[x + x1];
            if (cand < best || (cand == best && p.leftright == 1)) {
                best = cand;
                least = x1;
            }
        }
    }
    bdx = least;
    return best;
}

// liblqr's keep-old test of the incremental DP (A.8): same parent and (double) |m_old - m_new| < 1e-5.
// For floats that is |d| <= 0x3727C5AC, the largest float below the double 1e-5.
__device__ __forceinline__ bool keep_old(int pdx_old, int pdx_new, float m_old, float m_new)
{
    return pdx_old == pdx_new && fabsf(__fsub_rn(m_old, m_new)) <= __int_as_float(0x3727C5AC);
}

// ------------------------------------------------------------------------------------------------
__global__ void k_init_raw(int *raw, int w, int h)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x < w && y < h) raw[(size_t) y * w + x] = y * w + x;
}

// compact copy of the rigidity mask in current coordinates (start of a build_maps session)
__global__ void __launch_bounds__(256) k_gather_rig(const DevP p0, const DevP *tab)
{
    const DevP p = pick_image(p0, tab);
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= p.w || y >= p.h) return;
    p.rig[(size_t) y * p.pitch + x] = p.rigmask ? p.rigmask[p.raw[(size_t) y * p.raw_stride + x]] : 1.f;
}

// K1 -- A.3 full energy map (lqr_carver_build_emap).  A CTA owns a tile of EF_TW x EF_TH pixels of the CURRENT image:
// the brightness / luma of the tile and its one-pixel halo is computed ONCE per pixel (fp64, as liblqr's rcache) into
// shared memory -- RGBA pixels with one 4-byte load each, straight from the row when the index table is still the
// identity (a fresh, flattened or transposed carver: p.raw_ident), through the table otherwise -- and every thread then
// forms the central differences of its pixels from the staged values.  Same operations in the same order as
// energy_at(), so the band updates (k_energy_band) and this kernel agree bit for bit.
#define EF_TW 128
#define EF_TH 16
__global__ void __launch_bounds__(256) k_energy_full(const DevP p0, const DevP *tab)
{
    const DevP p = pick_image(p0, tab);
    __shared__ double sb[EF_TH + 2][EF_TW + 2];
    __shared__ double t255[256]; // v / 255.0, each quotient divided once per CTA
    const int x0 = blockIdx.x * EF_TW, y0 = blockIdx.y * EF_TH;
    const float inf = __int_as_float(0x7f800000);
    if (p.grad_kind != GRAD_NULL) {
        t255[threadIdx.x] = (double) threadIdx.x / 255.0;
        __syncthreads();
        for (int i = threadIdx.x; i < (EF_TH + 2) * (EF_TW + 2); i += 256) {
            const int ty = i / (EF_TW + 2), tx = i - ty * (EF_TW + 2);
            const int x = x0 + tx - 1, y = y0 + ty - 1;
            double v = 0.0;
            if (x >= 0 && x < p.w && y >= 0 && y < p.h)
                v = px_scalar(p, p.raw_ident ? y * p.w0 + x : p.raw[(size_t) y * p.raw_stride + x], t255);
            sb[ty][tx] = v;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < EF_TH * EF_TW; i += 256) {
        const int ty = i / EF_TW, tx = i - ty * EF_TW;
        const int x = x0 + tx, y = y0 + ty;
        if (x >= p.pitch || y >= p.h) continue;
        float e = inf; // columns right of the image hold +inf in en and m: a sentinel that never wins a minimum
        if (x < p.w) {
            e = 0.f;
            if (p.grad_kind != GRAD_NULL) {
                const double b = sb[ty + 1][tx + 1];
                double gx, gy;
                if (y == 0)
                    gy = __dsub_rn(p.h > 1 ? sb[ty + 2][tx + 1] : 0.0, b);
                else if (y < p.h - 1)
                    gy = __dmul_rn(__dsub_rn(sb[ty + 2][tx + 1], sb[ty][tx + 1]), 0.5);
                else
                    gy = __dsub_rn(b, sb[ty][tx + 1]);
                if (x == 0)
                    gx = __dsub_rn(p.w > 1 ? sb[ty + 1][tx + 2] : 0.0, b);
                else if (x < p.w - 1)
                    gx = __dmul_rn(__dsub_rn(sb[ty + 1][tx + 2], sb[ty + 1][tx]), 0.5);
                else
                    gx = __dsub_rn(b, sb[ty + 1][tx]);
                if (p.grad_kind == GRAD_NORM)
                    e = (float) sqrt(__dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy)));
                else if (p.grad_kind == GRAD_SUMABS)
                    e = (float) __dmul_rn(__dadd_rn(fabs(gx), fabs(gy)), 0.5);
                else
                    e = (float) fabs(gx);
            }
            if (p.bias) {
                const int z = p.raw_ident ? y * p.w0 + x : p.raw[(size_t) y * p.raw_stride + x];
                e = __fadd_rn(e, __fdiv_rn(p.bias[z], (float) p.w_start));
            }
        }
        p.en[(size_t) y * p.pitch + x] = e;
    }
}

// K1b -- A.8 energy band after a carve (lqr_carver_update_emap): eight threads per row derive the row's
// [nrg_xmin, nrg_xmax] from the seam positions of rows y-radius..y+radius and recompute that band (a few cells: the
// seam moves at most delta_x columns per row).  p.w is the width AFTER the carve; vpath_x is in pre-carve coordinates.
#define B200C_EB_ROWS 32 // rows per CTA of 256 threads
__global__ void __launch_bounds__(256) k_energy_band(const DevP pin0, const DevP *tab)
{
    pdl_entry();
    const DevP pin = pick_image(pin0, tab);
    const DevP p = seam_view(pin, 1);
    __shared__ double t255[256]; // v / 255.0, each quotient divided once per CTA (a cell reads 5 pixels x 4 channels)
    t255[threadIdx.x] = (double) threadIdx.x / 255.0;
    __syncthreads();
    const int sub = threadIdx.x & 7;
    const int y = blockIdx.x * B200C_EB_ROWS + (threadIdx.x >> 3);
    if (y >= p.h) return;
    const int r = p.nrg_radius;
    const int own = p.vpath_x[y];
    int xmin = own, xmax = own - 1;
    for (int y1 = max(y - r, 0); y1 <= min(y + r, p.h - 1); ++y1) {
        const int x = p.vpath_x[y1];
        xmin = min(xmin, x - r);
        xmax = max(xmax, x + r - 1);
    }
    xmin = max(0, xmin);
    xmax = min(p.w - 1, xmax);
    if (sub == 0) {
        p.nrg_xmin[y] = xmin;
        p.nrg_xmax[y] = xmax;
        p.nrg_pack[y] = ((unsigned) xmin & 0xffffffu) | ((unsigned) max(xmax - xmin + 1, 0) << 24);
    }
    for (int x = xmin + sub; x <= xmax; x += 8) p.en[(size_t) y * p.pitch + x] = energy_at(p, x, y, t255);
}

// K2 (generic) -- A.5 full m-map DP (lqr_carver_build_mmap): one CTA walks the rows, the row is spread
// over the threads, a block barrier separates dependent rows.  Correct for any width / delta_x; the
// cluster kernel of mmap_cluster.cuh is the fast path.
__global__ void __launch_bounds__(1024) k_mmap_full(const DevP p0, const DevP *tab)
{
    const DevP p = pick_image(p0, tab);
    for (int x = threadIdx.x; x < p.pitch; x += blockDim.x) p.m[x] = p.en[x]; // incl. the +inf sentinels
    __syncthreads();
    for (int y = 1; y < p.h; ++y) {
        const size_t o = (size_t) y * p.pitch;
        const float *up = p.m + o - p.pitch;
        for (int x = threadIdx.x; x < p.w; x += blockDim.x) {
            int bdx;
            const float best = best_parent(p, up, x, p.rig ? p.rig[o + x] : 1.f, bdx);
            p.pdx[o + x] = (int8_t) bdx;
            p.m[o + x] = __fadd_rn(p.en[o + x], best);
        }
        for (int x = p.w + threadIdx.x; x < p.pitch; x += blockDim.x) p.m[o + x] = __int_as_float(0x7f800000);
        __syncthreads();
    }
}

// K2b (generic) -- A.8 incremental DP (lqr_carver_update_mmap) with the keep-old rule.  Rows [y_from, h) are
// processed by one CTA; (lo, hi) is the hull of the cells of row y_from-1 whose value or parent changed
// (lo > hi: none).  Row y evaluates [min(lo, nrg_xmin[y]) - delta_x, max(hi, nrg_xmax[y]) + delta_x]: every
// cell with a changed parent, a changed energy or a changed set of parent candidates lies in that range, and a
// cell outside it re-evaluates to "keep" (DESIGN.md section 5), so the stored maps are the ones liblqr's
// self-trimming band produces.  s_red: 64 ints of shared scratch.
__device__ void update_rows_generic(const DevP &p, int y_from, int lo, int hi, int *s_red)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    unsigned long long cells = 0;
    for (int y = y_from; y < p.h; ++y) {
        const size_t o = (size_t) y * p.pitch;
        const int elo = max(min(lo, p.nrg_xmin[y]) - (y ? p.delta_x : 0), 0);
        const int ehi = min(max(hi, p.nrg_xmax[y]) + (y ? p.delta_x : 0), p.w - 1);
        int first = INT_MAX, last = INT_MIN;
        for (int x = elo + tid; x <= ehi; x += blockDim.x) {
            if (y == 0) { // row 0: m = en over the energy band; the band carries over to row 1 (A.8)
                p.m[x] = p.en[x];
                first = min(first, x);
                last = max(last, x);
                continue;
            }
            int bdx;
            const float new_m = __fadd_rn(p.en[o + x], best_parent(p, p.m + o - p.pitch, x, p.rig ? p.rig[o + x] : 1.f, bdx));
            if (!keep_old(p.pdx[o + x], bdx, p.m[o + x], new_m)) {
                p.m[o + x] = new_m;
                p.pdx[o + x] = (int8_t) bdx;
                first = min(first, x);
                last = max(last, x);
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
        __syncthreads(); // publishes this row's m / pdx and the per-warp extremes
        lo = INT_MAX, hi = INT_MIN;
        for (int i = 0; i < nwarp; ++i) {
            lo = min(lo, s_red[buf + i]);
            hi = max(hi, s_red[buf + 16 + i]);
        }
    }
    if (tid == 0 && p.cells) atomicAdd(p.cells, cells);
}

__global__ void __launch_bounds__(512) k_mmap_update(const DevP pin0, const DevP *tab)
{
    pdl_entry();
    const DevP pin = pick_image(pin0, tab);
    const DevP p = seam_view(pin, 1);
    __shared__ int s_red[64];
    update_rows_generic(p, 0, INT_MAX, INT_MIN, s_red);
}

// K3 (generic) -- A.6 seam: arg-min over the last row of m with the tie rule, then follow the parents upwards.
__device__ __forceinline__ bool seam_better(float ov, int ox, float v, int x, int leftright)
{
    if (ox < 0) return false;
    if (x < 0) return true;
    if (ov < v) return true;
    if (ov == v) return leftright ? (ox > x) : (ox < x);
    return false;
}

// block-wide arg-min over the last row (all threads must call); result valid in thread 0. s_v/s_x: 32 entries.
__device__ __forceinline__ int last_row_argmin(const DevP &p, float *s_v, int *s_x)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const float *row = p.m + (size_t) (p.h - 1) * p.pitch;
    float best = 536870912.f; // (float)(1 << 29)
    int bx = -1;
    for (int x = tid; x < p.w; x += blockDim.x) {
        const float v = row[x];
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
    }
    return bx < 0 ? 0 : bx;
}

__global__ void __launch_bounds__(1024) k_vpath(const DevP pin0, const DevP *tab)
{
    pdl_entry();
    const DevP pin = pick_image(pin0, tab);
    __shared__ float s_v[32];
    __shared__ int s_x[32];
    if (pin.dyn && threadIdx.x == 0) advance_seam(pin); // this iteration's seam
    __syncthreads();
    const DevP p = seam_view(pin, 0);
    int x = last_row_argmin(p, s_v, s_x);
    if (threadIdx.x != 0) return;
    for (int y = p.h - 1; y >= 0; --y) {
        p.vpath_x[y] = x;
        if (y > 0) {
            const int d = p.pdx[(size_t) y * p.pitch + x];
            if (d == B200C_PDX_NONE) {
                atomicOr(p.err, 2);
            } else {
                x = min(max(x + d, 0), p.w - 1);
            }
        }
    }
}

// K4 -- A.7 carve (lqr_carver_update_vsmap + lqr_carver_carve): mark the seam pixel in the visibility map and
// close the gap at x = s in the row's index table AND in the compact maps (en, m, rig shift left by one; pdx
// shifts and is re-expressed so that it still names the same parent pixel: a parent right of the upper row's
// seam moved one column left, a parent ON that seam is gone).  One CTA per row; p.w is the width AFTER the
// carve.  Loads of a chunk are parked in registers before the barrier, so the in-place shift never reads a
// slot another thread already overwrote.
#define B200C_CARVE_THREADS 256
#define B200C_CARVE_ITEMS 4
// phase 0: the whole row.  phases 1 / 2 split it so that the band DP of the same seam can start early (it only reads
// columns near the seam): 1 = NEAR, the visibility mark and the first chunk of B200C_CARVE_SPAN columns from the seam;
// 2 = FAR, every later chunk, launched after NEAR (it overwrites the column NEAR's last cell is read from) and counted in
// p.far when done.  Whoever moves the row's last cell also writes the +inf sentinel into the vacated column.
#define B200C_CARVE_SPAN (B200C_CARVE_THREADS * B200C_CARVE_ITEMS)
__global__ void __launch_bounds__(B200C_CARVE_THREADS) k_carve(const DevP pin0, int vs_value, int phase, const DevP *tab)
{
    pdl_entry();
    const DevP pin = pick_image(pin0, tab);
    int seam;
    const DevP p = seam_view(pin, 1, &seam);
    vs_value += seam; // the level this seam's pixels get in the visibility map
    const int y = blockIdx.x;
    const size_t o = (size_t) y * p.pitch;
    int *raw = p.raw + (size_t) y * p.raw_stride;
    float *en = p.en + o, *m = p.m + o, *rig = p.rig ? p.rig + o : nullptr;
    int8_t *pdx = p.pdx + o;
    const int s = p.vpath_x[y];
    const int sp = y > 0 ? p.vpath_x[y - 1] : INT_MIN;
    if (threadIdx.x == 0 && phase != 2) p.vs[raw[s]] = vs_value; // read before the first barrier, i.e. before any store
    const int start = max(s - p.delta_x - 1, 0);   // cells left of the seam stay put but may see their parent move
    const bool one_chunk = start + B200C_CARVE_SPAN >= p.w; // NEAR moves the whole row
    const int from = phase == 2 ? start + B200C_CARVE_SPAN : start;
    const int to = phase == 1 ? min(p.w, start + B200C_CARVE_SPAN) : p.w;
    for (int base = from; base < to; base += B200C_CARVE_SPAN) {
        int vr[B200C_CARVE_ITEMS], vd[B200C_CARVE_ITEMS];
        float ve[B200C_CARVE_ITEMS], vm[B200C_CARVE_ITEMS], vg[B200C_CARVE_ITEMS];
#pragma unroll
        for (int i = 0; i < B200C_CARVE_ITEMS; ++i) {
            const int xn = base + i * B200C_CARVE_THREADS + threadIdx.x; // new column
            const int xo = xn + (xn >= s);                               // the column it had before
            vd[i] = B200C_PDX_NONE;
            if (xn < p.w) {
                if (xn >= s) {
                    vr[i] = raw[xo];
                    ve[i] = en[xo];
                    vm[i] = m[xo];
                    if (rig) vg[i] = rig[xo];
                }
                if (y > 0) {
                    const int d = pdx[xo];
                    const int xp = xo + d; // old column of the parent in the upper row
                    if (d != B200C_PDX_NONE && xp != sp) vd[i] = (xp - (xp > sp)) - xn;
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < B200C_CARVE_ITEMS; ++i) {
            const int xn = base + i * B200C_CARVE_THREADS + threadIdx.x;
            if (xn < p.w) {
                if (xn >= s) {
                    raw[xn] = vr[i];
                    en[xn] = ve[i];
                    m[xn] = vm[i];
                    if (rig) rig[xn] = vg[i];
                }
                if (y > 0) pdx[xn] = (int8_t) vd[i];
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (phase == 0 || (phase == 1) == one_chunk) en[p.w] = m[p.w] = __int_as_float(0x7f800000); // the vacated column joins the sentinels
        if (phase != 1 && p.far) { // FAR done (a whole-row launch counts as both phases)
            __threadfence();
            atomicAdd(p.far, 1);
        }
    }
}

// The whole-row carve of the per-seam loop: the same result as k_carve phase 0, one CTA per row, in ONE pass: the part of
// the row from the seam on (from column base = (s - delta_x - 1) & ~15) is fetched into shared memory by bulk copies
// (TMA: one per map, issued by one thread, no register parking, the whole row in flight), and once they have landed the
// threads write the shifted row back with 16-byte stores.  A thread owns groups of four consecutive NEW columns
// [xn0, xn0 + 4), xn0 a multiple of 4 (the rows of en / m / rig / pdx are 16-byte aligned).  Cells left of the seam
// inside the first groups are written back as they were.  Nothing is stored before everything is loaded, so the in-place
// shift needs no ordering between the threads.  The index table's rows (stride w_start, no padding) are fetched by a
// bulk copy when they are 16-byte aligned, by ordinary loads otherwise.
__host__ __device__ inline size_t carve_row_smem(int pitch, bool rig) { return (size_t) (pitch + 16) * (rig ? 17 : 13) + 64; }
#define B200C_CARVE_ROW_SMEM_MAX (200 * 1024)

__device__ __forceinline__ unsigned cv_saddr(const void *q) { return (unsigned) __cvta_generic_to_shared(q); }
__device__ __forceinline__ void cv_bulk(void *dst_smem, const void *src, unsigned bytes, void *mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(cv_saddr(dst_smem)),
                 "l"(src), "r"(bytes), "r"(cv_saddr(mbar))
                 : "memory");
}

__global__ void __launch_bounds__(B200C_CARVE_THREADS) k_carve_row(const DevP pin0, int vs_value, int /*phase*/, const DevP *tab)
{
    pdl_entry();
    const DevP pin = pick_image(pin0, tab);
    int seam;
    const DevP p = seam_view(pin, 1, &seam);
    vs_value += seam; // the level this seam's pixels get in the visibility map
    extern __shared__ __align__(16) unsigned char cv_smem[];
    __shared__ __align__(8) unsigned long long mbar;
    const int y = blockIdx.x, tid = threadIdx.x;
    const size_t o = (size_t) y * p.pitch;
    int *raw = p.raw + (size_t) y * p.raw_stride;
    float *en = p.en + o, *m = p.m + o, *rig = p.rig ? p.rig + o : nullptr;
    int8_t *pdx = p.pdx + o;
    const int s = p.vpath_x[y];
    const int sp = y > 0 ? p.vpath_x[y - 1] : INT_MIN;
    const int start = max(s - p.delta_x - 1, 0); // cells left of the seam stay put but may see their parent move
    const int base = start & ~15;
    const int n = p.pitch - base;                // staged columns: a multiple of 16
    const int nraw = min(n, (p.w + 1 - base + 3) & ~3); // the index table's row has p.w + 1 entries
    const int cap = p.pitch + 16;
    float *se = reinterpret_cast<float *>(cv_smem), *sm = se + cap, *sg = sm + cap;
    int *sr = reinterpret_cast<int *>(rig ? sg + cap : sg);
    unsigned char *sd = reinterpret_cast<unsigned char *>(sr + cap);
    const bool raw_bulk = ((reinterpret_cast<size_t>(raw + base) & 15) == 0);
    const float inf = __int_as_float(0x7f800000);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cv_saddr(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned bytes = (unsigned) n * (rig ? 12u : 8u) + (y > 0 ? (unsigned) n : 0u) + (raw_bulk ? (unsigned) nraw * 4u : 0u);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cv_saddr(&mbar)), "r"(bytes) : "memory");
        cv_bulk(se, en + base, (unsigned) n * 4u, &mbar);
        cv_bulk(sm, m + base, (unsigned) n * 4u, &mbar);
        if (rig) cv_bulk(sg, rig + base, (unsigned) n * 4u, &mbar);
        if (y > 0) cv_bulk(sd, pdx + base, (unsigned) n, &mbar);
        if (raw_bulk) cv_bulk(sr, raw + base, (unsigned) nraw * 4u, &mbar);
        p.vs[raw[s]] = vs_value; // read before anything is stored
    }
    if (!raw_bulk)
        for (int i = tid; i < p.w + 1 - base; i += B200C_CARVE_THREADS) sr[i] = raw[base + i];
    __syncthreads(); // the barrier is initialised; the index table is staged
    {
        unsigned ok = 0;
        unsigned long long t0 = 0;
        for (unsigned tries = 0; !ok; ++tries) {
            asm volatile("{\n.reg .pred q;\nmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\nselp.u32 %0, 1, 0, q;\n}\n"
                         : "=r"(ok)
                         : "r"(cv_saddr(&mbar))
                         : "memory");
            if (!ok && (tries & 1023u) == 1023u) { // a copy that never lands is a bug, not a reason to hang the GPU
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                if (t0 == 0) t0 = t;
                else if (t - t0 > 2000000000ull) {
                    if (tid == 0) atomicOr(p.err, 8);
                    return;
                }
            }
        }
    }
    const int ngroups = (p.w - base + 3) >> 2; // groups with at least one column of the image
    for (int g = tid; g < ngroups; g += B200C_CARVE_THREADS) {
        const int k = 4 * g, xn0 = base + k;
        const float4 e4 = *reinterpret_cast<const float4 *>(se + k), m4 = *reinterpret_cast<const float4 *>(sm + k);
        const float4 g4 = rig ? *reinterpret_cast<const float4 *>(sg + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int4 r4 = *reinterpret_cast<const int4 *>(sr + k);
        const unsigned d4 = *reinterpret_cast<const unsigned *>(sd + k);
        // the first cell of the next group (past the staged columns only when no cell of this group moves)
        const float oe[5] = {e4.x, e4.y, e4.z, e4.w, se[k + 4]}, om[5] = {m4.x, m4.y, m4.z, m4.w, sm[k + 4]};
        const float og[5] = {g4.x, g4.y, g4.z, g4.w, rig ? sg[k + 4] : 0.f};
        const int orw[5] = {r4.x, r4.y, r4.z, r4.w, sr[k + 4]};
        const unsigned d5 = sd[k + 4];
        float xe[4], xm[4], xg[4];
        int xr[4];
        unsigned pk = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int xn = xn0 + i;
            const bool moved = xn >= s && xn < p.w; // takes the value of the cell right of it
            xr[i] = moved ? orw[i + 1] : orw[i];
            xg[i] = moved ? og[i + 1] : og[i];
            xe[i] = xn >= p.w ? inf : (moved ? oe[i + 1] : oe[i]); // the vacated column joins the +inf sentinels
            xm[i] = xn >= p.w ? inf : (moved ? om[i + 1] : om[i]);
            const int dsame = (int) (signed char) (d4 >> (8 * i));
            const int dnext = (int) (signed char) (i < 3 ? (d4 >> (8 * i + 8)) : d5);
            int dn = dsame; // cells left of `start` or past the image keep their byte
            if (xn >= start && xn < p.w) {
                const int xo = xn + (xn >= s), d = xn >= s ? dnext : dsame;
                const int xp = xo + d; // old column of the parent in the upper row
                dn = (d != B200C_PDX_NONE && xp != sp) ? (xp - (xp > sp)) - xn : B200C_PDX_NONE;
            }
            pk |= ((unsigned) dn & 0xffu) << (8 * i);
        }
        *reinterpret_cast<float4 *>(en + xn0) = make_float4(xe[0], xe[1], xe[2], xe[3]);
        *reinterpret_cast<float4 *>(m + xn0) = make_float4(xm[0], xm[1], xm[2], xm[3]);
        if (rig) *reinterpret_cast<float4 *>(rig + xn0) = make_float4(xg[0], xg[1], xg[2], xg[3]);
        if (y > 0) *reinterpret_cast<unsigned *>(pdx + xn0) = pk;
        if (raw_bulk && xn0 + 3 <= p.w) {
            *reinterpret_cast<int4 *>(raw + xn0) = make_int4(xr[0], xr[1], xr[2], xr[3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (xn0 + i <= p.w) raw[xn0 + i] = xr[i];
        }
    }
    if (tid == 0) {
        if ((p.w & 3) == 0) en[p.w] = m[p.w] = inf; // the vacated column starts a group of its own: no thread wrote it
        if (p.far) { // a whole-row launch counts as both phases of the split carve
            __threadfence();
            atomicAdd(p.far, 1);
        }
    }
}

// A.7 finish_vsmap: the image is one pixel wide; the survivors get the largest level.
__global__ void k_finish_vsmap(const DevP pin0, const DevP *tab)
{
    const DevP pin = pick_image(pin0, tab);
    const DevP p = seam_view(pin, 1);
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y < p.h) p.vs[p.raw[(size_t) y * p.raw_stride]] = p.w0;
}

// test hook: the compact maps scattered back to liblqr's physical layout (index = pixel id); what: 0 en, 1 m,
// 2 least (pixel id of the parent, -1 when it has none / it is gone)
__global__ void __launch_bounds__(256) k_export_physical(DevP p, int what, int *out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= p.w || y >= p.h) return;
    const int z = p.raw[(size_t) y * p.raw_stride + x];
    const size_t o = (size_t) y * p.pitch + x;
    if (what == 0) {
        out[z] = __float_as_int(p.en[o]);
    } else if (what == 1) {
        out[z] = __float_as_int(p.m[o]);
    } else {
        const int d = p.pdx[o];
        out[z] = (y > 0 && d != B200C_PDX_NONE) ? p.raw[(size_t) (y - 1) * p.raw_stride + x + d] : -1;
    }
}

// ------------------------------------------------------------------------------------------------
// Block-wide exclusive scan of one int per thread (blockDim.x <= 1024, multiple of 32).
// s_warp must hold 33 ints.  Returns the exclusive prefix; *total receives the block sum.
__device__ __forceinline__ int block_excl_scan(int v, int *s_warp, int *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int t = lane < nwarp ? s_warp[lane] : 0;
        int ti = t;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, ti, off);
            if (lane >= off) ti += u;
        }
        if (lane < nwarp) s_warp[lane] = ti - t; // exclusive warp offsets
        if (lane == 31) s_warp[32] = ti;         // grand total
    }
    __syncthreads();
    const int res = s_warp[warp] + incl - v;
    *total = s_warp[32];
    __syncthreads(); // s_warp may be reused right away
    return res;
}

__device__ __forceinline__ bool is_visible(int vs, int level) { return vs == 0 || vs >= level; }

// K6 / K7 -- A.12 read-out and A.11 flatten: per internal row, stream-compact the visible pixels
// (vs == 0 or vs >= level) of the w0-wide row into a dense w-wide row; optionally the float maps too.
#define B200C_ROW_THREADS 256
__global__ void __launch_bounds__(B200C_ROW_THREADS)
k_compact_rows(const uint8_t *rgb, const int *vs, int w0, int w, int channels, int level, uint8_t *out_rgb,
               const float *bias, float *out_bias, const float *rigmask, float *out_rigmask)
{
    __shared__ int s_warp[33];
    const int y = blockIdx.x;
    const size_t in0 = (size_t) y * w0, out0 = (size_t) y * w;
    int base = 0;
    for (int c0 = 0; c0 < w0; c0 += B200C_ROW_THREADS) {
        const int x = c0 + threadIdx.x;
        const bool vis = x < w0 && is_visible(vs[in0 + x], level);
        int total;
        const int pos = base + block_excl_scan(vis ? 1 : 0, s_warp, &total);
        if (vis && pos < w) {
            const uint8_t *src = rgb + (in0 + x) * channels;
            uint8_t *dst = out_rgb + (out0 + pos) * channels;
            for (int k = 0; k < channels; ++k) dst[k] = src[k];
            if (out_bias) out_bias[out0 + pos] = bias[in0 + x];
            if (out_rigmask) out_rigmask[out0 + pos] = rigmask[in0 + x];
        }
        base += total;
    }
}

// K10 -- A.13 visibility-map dump: at the reference width (level = depth + 1) the visible pixels of row y,
// in order, get vs == 0 ? 0 : vs - depth, written in image orientation.
__global__ void __launch_bounds__(B200C_ROW_THREADS)
k_vmap_rows(const int *vs, int w0, int w_ref, int h, int depth, int transposed, int *out)
{
    __shared__ int s_warp[33];
    const int y = blockIdx.x;
    const size_t in0 = (size_t) y * w0;
    const int level = depth + 1;
    int base = 0;
    for (int c0 = 0; c0 < w0; c0 += B200C_ROW_THREADS) {
        const int x = c0 + threadIdx.x;
        const int v = x < w0 ? vs[in0 + x] : 1;
        const bool vis = x < w0 && is_visible(v, level);
        int total;
        const int pos = base + block_excl_scan(vis ? 1 : 0, s_warp, &total);
        if (vis && pos < w_ref) {
            const size_t o = transposed ? (size_t) pos * h + y : (size_t) y * w_ref + pos;
            out[o] = v == 0 ? 0 : v - depth;
        }
        base += total;
    }
}

// K5 -- A.9 inflate: walk the w0-wide row (all pixels visible at level 1); a pixel that belongs to a seam
// found in this session (vs in [2*max_level-1, l+max_level-1]) is preceded by a new pixel whose channels
// are the integer mean of it and its left neighbour; vs is remapped; never-carved pixels refill the raw
// index table.  new_vs / raw may be NULL (attached carvers only carry pixels).
__global__ void __launch_bounds__(B200C_ROW_THREADS)
k_inflate_rows(const uint8_t *rgb, const int *vs, int w0, int w1, int channels, int l, int max_level,
               uint8_t *new_rgb, int *new_vs, const float *bias, float *new_bias, const float *rigmask,
               float *new_rigmask, int *raw, int raw_stride)
{
    __shared__ int s_warp[33];
    const int y = blockIdx.x;
    const size_t in0 = (size_t) y * w0, out0 = (size_t) y * w1;
    int base = 0, zbase = 0;
    for (int c0 = 0; c0 < w0; c0 += B200C_ROW_THREADS) {
        const int x = c0 + threadIdx.x;
        const bool in = x < w0;
        const int v = in ? vs[in0 + x] : 1;
        const bool dup = in && v != 0 && v <= l + max_level - 1 && v >= 2 * max_level - 1;
        int total, ztotal;
        int pos = base + block_excl_scan(in ? (dup ? 2 : 1) : 0, s_warp, &total);
        const int zpos = zbase + block_excl_scan(in && v == 0 ? 1 : 0, s_warp, &ztotal);
        if (in) {
            const size_t now = in0 + x;
            if (dup) {
                const size_t left = x > 0 ? now - 1 : now;
                const size_t z0 = out0 + pos;
                for (int k = 0; k < channels; ++k)
                    new_rgb[z0 * channels + k] =
                        (uint8_t) (((int) rgb[left * channels + k] + (int) rgb[now * channels + k]) / 2);
                if (new_bias) new_bias[z0] = __fmul_rn(__fadd_rn(bias[left], bias[now]), 0.5f);
                if (new_rigmask) new_rigmask[z0] = __fmul_rn(__fadd_rn(rigmask[left], rigmask[now]), 0.5f);
                if (new_vs) new_vs[z0] = l - v + max_level;
                ++pos;
            }
            const size_t z0 = out0 + pos;
            for (int k = 0; k < channels; ++k) new_rgb[z0 * channels + k] = rgb[now * channels + k];
            if (new_bias) new_bias[z0] = bias[now];
            if (new_rigmask) new_rigmask[z0] = rigmask[now];
            if (v != 0) {
                if (new_vs) new_vs[z0] = v + l - max_level + 1;
            } else if (raw) {
                raw[(size_t) y * raw_stride + zpos] = (int) z0;
            }
        }
        base += total;
        zbase += ztotal;
    }
}

// K8 -- A.11 transpose of a (h x w) array of ELEM-byte elements through a 32x32 shared-memory tile.
template <int ELEM>
struct Elem {
    uint8_t b[ELEM];
};
template <int ELEM>
__global__ void __launch_bounds__(256) k_transpose(const uint8_t *in, uint8_t *out, int w, int h)
{
    __shared__ Elem<ELEM> tile[32][33];
    const Elem<ELEM> *src = reinterpret_cast<const Elem<ELEM> *>(in);
    Elem<ELEM> *dst = reinterpret_cast<Elem<ELEM> *>(out);
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5; // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const int x = x0 + tx, y = y0 + j;
        if (x < w && y < h) tile[j][tx] = src[(size_t) y * w + x];
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int y = y0 + tx, x = x0 + j; // out is (w x h): out[x][y]
        if (x < w && y < h) dst[(size_t) x * h + y] = tile[tx][j];
    }
}

// K9 -- A.4 preservation / discard bias and rigidity mask from an 8-bit mask layer placed at (x_off, y_off).
__global__ void k_bias_add(float *bias, int w0, const uint8_t *mask, int channels, int mw, int bias_factor,
                           int x0, int y0, int x1, int y1, int nx, int ny)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= nx || y >= ny) return;
    const int has_alpha = (channels == 2 || channels >= 4);
    const int cc = channels - has_alpha;
    const size_t px = (size_t) (y - y0) * mw + (x - x0);
    int sum = 0;
    for (int k = 0; k < cc; ++k) sum += mask[px * channels + k];
    float b = (float) (__dmul_rn((double) bias_factor, (double) sum) / (double) (2 * 255 * cc));
    if (has_alpha) b = __fmul_rn(b, __fdiv_rn((float) mask[(px + 1) * channels - 1], 255.f));
    float *dst = bias + (size_t) (y + y1) * w0 + (x + x1);
    *dst = __fadd_rn(*dst, b);
}

__global__ void k_rigmask_set(float *rigmask, int w0, const uint8_t *mask, int channels, int mw, int x0, int y0,
                              int x1, int y1, int nx, int ny)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= nx || y >= ny) return;
    const int has_alpha = (channels == 2 || channels >= 4);
    const int cc = channels - has_alpha;
    const size_t px = (size_t) (y - y0) * mw + (x - x0);
    int sum = 0;
    for (int k = 0; k < cc; ++k) sum += mask[px * channels + k];
    float v = __fdiv_rn((float) sum, (float) (255 * cc));
    if (has_alpha) v = __fmul_rn(v, __fdiv_rn((float) mask[(px + 1) * channels - 1], 255.f));
    rigmask[(size_t) (y + y1) * w0 + (x + x1)] = v;
}

// energy read-back in image orientation (internal row i, column j)
__global__ void k_energy_export(DevP p, int transposed, float *out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= p.w || i >= p.h) return;
    const float v = p.en[(size_t) i * p.pitch + j];
    out[transposed ? (size_t) j * p.h + i : (size_t) i * p.w + j] = v;
}

// ------------------------------------------------------------------------------------------------
// The plug-in's own host loops next to the hot path, as kernels.
//
// write_vmap_to_layer's colouring (reference src/io_functions.c:249-279): seam order k of depth d -> value =
// (d+1-k)/(d+1) in double; colour = value*start + (1-value)*end, alpha = 0.5*(1+value); bytes = (guchar)(255*x), i.e.
// truncated.  Built with -fmad=false: the products and the sum round separately, like the reference on x86-64.
__global__ void __launch_bounds__(256) k_vmap_colour(const int *vmap, size_t n, int depth, double sr, double sg, double sb,
                                                     double er, double eg, double eb, uchar4 *out)
{
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int vs = vmap[i];
    uchar4 px = make_uchar4(0, 0, 0, 0);
    if (vs != 0) {
        const double value = __ddiv_rn((double) (depth + 1 - vs), (double) (depth + 1));
        const double rest = __dsub_rn(1.0, value);
        const double rd = __dadd_rn(__dmul_rn(value, sr), __dmul_rn(rest, er));
        const double gr = __dadd_rn(__dmul_rn(value, sg), __dmul_rn(rest, eg));
        const double bl = __dadd_rn(__dmul_rn(value, sb), __dmul_rn(rest, eb));
        const double al = __dmul_rn(0.5, __dadd_rn(1.0, value));
        px.x = (unsigned char) __double2int_rz(__dmul_rn(255.0, rd));
        px.y = (unsigned char) __double2int_rz(__dmul_rn(255.0, gr));
        px.z = (unsigned char) __double2int_rz(__dmul_rn(255.0, bl));
        px.w = (unsigned char) __double2int_rz(__dmul_rn(255.0, al));
    }
    out[i] = px;
}

// guess_new_size's reduction (reference src/layers_combo.c:347-386): per line (a mask row, or a mask column when
// vertical) the number of pixels whose intensity -- mean of the colour channels / 255, times alpha / 255 -- reaches
// 0.5 / colour channels; the maximum over the lines.  One warp per line; *result starts at 0.
__global__ void __launch_bounds__(256) k_guess_mask_size(const unsigned char *mask, int width, int bpp, int has_alpha,
                                                         int line0, int nlines, int first, int count, int vertical,
                                                         int *result)
{
    const int line = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (line >= nlines) return;
    const int c_bpp = bpp - (has_alpha ? 1 : 0);
    const double thr = __ddiv_rn(0.5, (double) c_bpp);
    int n = 0;
    for (int z = lane; z < count; z += 32) {
        const size_t at = vertical ? (size_t) (first + z) * width + (size_t) (line0 + line)
                                   : (size_t) (line0 + line) * width + (size_t) (first + z);
        const unsigned char *px = mask + at * bpp;
        double sum = 0;
        for (int k = 0; k < c_bpp; ++k) sum = __dadd_rn(sum, (double) px[k]);
        sum = __ddiv_rn(sum, (double) (255 * c_bpp));
        if (has_alpha) sum = __dmul_rn(sum, __ddiv_rn((double) px[bpp - 1], 255.0));
        n += sum >= thr ? 1 : 0;
    }
    n = __reduce_add_sync(0xffffffffu, n);
    if (lane == 0) atomicMax(result, n);
}

} // namespace b200c
