// ubench_stage.cu -- how fast ONE CTA can pull ~150 KB of L2-resident bytes into shared memory (the two staging phases of
// k_seam_chase, seam_trace.cuh): 16-byte cp.async pieces vs one bulk copy (TMA, cp.async.bulk) per row vs LDG+STS.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/ubench_stage.cu -o tools/bin/ubench_stage
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int PITCH = 3856, NROW = 68, REACH = 32, H = 2160;
constexpr int DYN = 176 * 1024;

__global__ void k_fill(signed char *a, size_t n)
{
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) a[i] = (signed char) (i % 3) - 1;
}

__device__ __forceinline__ unsigned saddr(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(void *d, const void *s) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr(d)), "l"(s) : "memory"); }
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void bulk(void *d, const void *s, unsigned bytes, void *mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(saddr(d)), "l"(s), "r"(bytes),
                 "r"(saddr(mbar))
                 : "memory");
}
__device__ __forceinline__ void mbar_init(void *m, unsigned n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(m)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect(void *m, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(m)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(void *m, unsigned parity)
{
    unsigned ok = 0;
    while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(saddr(m)), "r"(parity) : "memory");
}

__device__ __forceinline__ void tri_row(int k, int xg, int &a, int &e)
{
    a = max(xg - k * REACH, 0) & ~15;
    e = min((xg + k * REACH + 16) & ~15, PITCH);
}

// mode 0: cp.async 16 B, a warp per row; 1: bulk copy per row; 2: LDG.128 + STS.128
__global__ void __launch_bounds__(1024, 1) k_triangle(const signed char *jump, int xg, int mode, long long *cyc, int *sink)
{
    extern __shared__ __align__(128) unsigned char stage[];
    __shared__ int roff[NROW + 1];
    __shared__ unsigned long long mbar;
    const int tid = threadIdx.x;
    if (tid == 0) {
        int o = 0;
        for (int k = 0; k < NROW; ++k) {
            int a, e;
            tri_row(k, xg, a, e);
            roff[k] = o;
            o += e - a;
        }
        roff[NROW] = o;
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    if (mode == 0) {
        for (int k = tid >> 5; k < NROW; k += 32) {
            int a, e;
            tri_row(k, xg, a, e);
            const signed char *src = jump + (size_t) k * PITCH + a;
            for (int i = tid & 31; i < (e - a) >> 4; i += 32) cp16(stage + roff[k] + (i << 4), src + (i << 4));
        }
        cp_wait();
    } else if (mode == 1) {
        if (tid == 0) mbar_expect(&mbar, (unsigned) roff[NROW]);
        __syncthreads();
        if (tid < NROW) {
            int a, e;
            tri_row(tid, xg, a, e);
            bulk(stage + roff[tid], jump + (size_t) tid * PITCH + a, (unsigned) (e - a), &mbar);
        }
        mbar_wait(&mbar, 0);
    } else {
        for (int k = tid >> 5; k < NROW; k += 32) {
            int a, e;
            tri_row(k, xg, a, e);
            const int4 *src = reinterpret_cast<const int4 *>(jump + (size_t) k * PITCH + a);
            int4 *dst = reinterpret_cast<int4 *>(stage + roff[k]);
            for (int i = tid & 31; i < (e - a) >> 4; i += 32) dst[i] = __ldcg(src + i);
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) {
        cyc[0] = t1 - t0;
        cyc[1] = roff[NROW];
    }
    sink[tid] = stage[(tid * 131) % roff[NROW]];
}

// the parent tiles of the re-walk: NROW blocks x 32 rows x 80 bytes, rows PITCH apart
// mode 0: the index arithmetic of the kernel as it is; 1: a warp per block, lane = row, 5 pieces each; 2: bulk copy per (block, row)
__global__ void __launch_bounds__(1024, 1) k_tiles(const signed char *pdx, const int *ent, int mode, long long *cyc, int *sink)
{
    extern __shared__ __align__(128) unsigned char stage[];
    __shared__ unsigned long long mbar;
    const int tid = threadIdx.x, R = 32, twf = 80, pieces = twf >> 4;
    if (tid == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    if (mode == 0) {
        for (int i = tid; i < NROW * R * pieces; i += 1024) {
            const int bb = i / (R * pieces), rem = i - bb * (R * pieces), r = rem / pieces, c = (rem - r * pieces) << 4;
            const int y = H - 1 - bb * R - r;
            if (y < 1) continue;
            const int lo = min(max(ent[bb] - REACH, 0) & ~15, PITCH - twf);
            cp16(stage + ((size_t) bb * R + r) * twf + c, pdx + (size_t) y * PITCH + lo + c);
        }
        cp_wait();
    } else if (mode == 1) {
        for (int bb = tid >> 5; bb < NROW; bb += 32) {
            const int r = tid & 31, y = H - 1 - bb * R - r;
            if (y < 1) continue;
            const int lo = min(max(ent[bb] - REACH, 0) & ~15, PITCH - twf);
            const signed char *src = pdx + (size_t) y * PITCH + lo;
            unsigned char *dst = stage + ((size_t) bb * R + r) * twf;
#pragma unroll
            for (int c = 0; c < 5; ++c) cp16(dst + 16 * c, src + 16 * c);
        }
        cp_wait();
    } else {
        int n = 0;
        for (int i = tid; i < NROW * R; i += 1024) n += (H - 1 - i >= 1);
        if (tid == 0) {
            int tot = 0;
            for (int i = 0; i < NROW * R; ++i) tot += (H - 1 - i >= 1);
            mbar_expect(&mbar, (unsigned) tot * twf);
        }
        __syncthreads();
        for (int i = tid; i < NROW * R; i += 1024) {
            const int bb = i >> 5, y = H - 1 - i;
            if (y < 1) continue;
            const int lo = min(max(ent[bb] - REACH, 0) & ~15, PITCH - twf);
            bulk(stage + (size_t) i * twf, pdx + (size_t) y * PITCH + lo, twf, &mbar);
        }
        mbar_wait(&mbar, 0);
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) cyc[0] = t1 - t0;
    sink[tid] = stage[(tid * 131) % (NROW * R * twf)];
}

// the row walk of k_seam_chase: 272 threads, 8 dependent one-byte loads each, rows PITCH apart, data last touched by a
// grid-wide read (as k_seam_jumps leaves it) or written by a grid-wide kernel
__global__ void k_touch(const signed char *a, size_t n, int *sink)
{
    int acc = 0;
    for (size_t i = (blockIdx.x * (size_t) blockDim.x + threadIdx.x) * 16; i < n; i += (size_t) gridDim.x * blockDim.x * 16)
        acc += __ldcg(reinterpret_cast<const int4 *>(a + i)).x;
    if (acc == 0x12345678) sink[0] = acc;
}
// spread: walkers per warp (32: the 272 walkers fill 9 warps; 9: spread over all 32 warps; 1: one walker per warp)
__global__ void __launch_bounds__(1024, 1) k_walk(const signed char *pdx, int steps, long long *cyc, int *out, int spread = 32)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tid = spread == 32 ? (int) threadIdx.x : (lane < spread ? warp * spread + lane : 1 << 20);
    __syncthreads();
    const long long t0 = clock64();
    if (tid < 272) {
        int x = 1900 + (tid * 7) % 200, y = H - 1 - tid * 8;
        for (int s = 0; s < steps && y - s >= 1; ++s) {
            const int d = __ldcg(pdx + (size_t) (y - s) * PITCH + x);
            x = min(max(x + d, 0), PITCH - 1);
        }
        out[tid] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) cyc[0] = clock64() - t0;
}

int main()
{
    signed char *jump, *pdx;
    int *ent, *sink;
    long long *cyc;
    CHECK(cudaMalloc(&jump, (size_t) PITCH * NROW));
    CHECK(cudaMalloc(&pdx, (size_t) PITCH * H));
    CHECK(cudaMalloc(&ent, 4 * (NROW + 1)));
    CHECK(cudaMalloc(&sink, 4096));
    CHECK(cudaMallocManaged(&cyc, 64));
    int h_ent[NROW + 1];
    for (int i = 0; i <= NROW; ++i) h_ent[i] = 1900 + (i * 37) % 200;
    CHECK(cudaMemcpy(ent, h_ent, sizeof h_ent, cudaMemcpyHostToDevice));
    CHECK(cudaFuncSetAttribute(k_triangle, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN));
    CHECK(cudaFuncSetAttribute(k_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN));
    const char *tn[3] = {"cp.async 16 B, warp per row", "bulk copy per row (TMA)", "LDG.128 + STS.128"};
    for (int mode = 0; mode < 3; ++mode)
        for (int rep = 0; rep < 3; ++rep) {
            k_fill<<<148 * 4, 256>>>(jump, (size_t) PITCH * NROW);
            k_triangle<<<1, 1024, DYN>>>(jump, 1900, mode, cyc, sink);
            CHECK(cudaDeviceSynchronize());
            if (rep == 2) printf("triangle %-30s: %6lld cycles for %lld bytes\n", tn[mode], cyc[0], cyc[1]);
        }
    const char *qn[3] = {"cp.async 16 B, div/mod indexing", "cp.async 16 B, lane = row", "bulk copy per row of 80 B"};
    for (int mode = 0; mode < 3; ++mode)
        for (int rep = 0; rep < 3; ++rep) {
            k_fill<<<148 * 4, 256>>>(pdx, (size_t) PITCH * H);
            k_tiles<<<1, 1024, DYN>>>(pdx, ent, mode, cyc, sink);
            CHECK(cudaDeviceSynchronize());
            if (rep == 2) printf("tiles    %-30s: %6lld cycles for %d bytes\n", qn[mode], cyc[0], NROW * 32 * 80);
        }
    for (int mode = 0; mode < 2; ++mode)
        for (int rep = 0; rep < 3; ++rep) {
            k_fill<<<148 * 4, 256>>>(pdx, (size_t) PITCH * H);
            if (mode == 1) k_touch<<<148 * 4, 256>>>(pdx, (size_t) PITCH * H, sink);
            k_walk<<<1, 1024>>>(pdx, 8, cyc, sink);
            CHECK(cudaDeviceSynchronize());
            if (rep == 2) printf("walk 8 dependent byte loads x 272 threads, %s: %lld cycles (%lld per step)\n",
                                 mode ? "after a grid-wide read " : "after a grid-wide write", cyc[0], cyc[0] / 8);
        }
    for (int spread : {32, 9, 1}) {
        for (int rep = 0; rep < 3; ++rep) {
            k_fill<<<148 * 4, 256>>>(pdx, (size_t) PITCH * H);
            k_touch<<<148 * 4, 256>>>(pdx, (size_t) PITCH * H, sink);
            k_walk<<<1, 1024>>>(pdx, 8, cyc, sink, spread);
            CHECK(cudaDeviceSynchronize());
            if (rep == 2) printf("walk, %2d walkers per warp: %lld cycles (%lld per step)\n", spread, cyc[0], cyc[0] / 8);
        }
    }
    return 0;
}
