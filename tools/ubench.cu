// ubench.cu -- single-SM latency probes behind the design of the per-seam kernels: how long one
// dependent row step of the band DP takes for a lone warp, what a shuffle / vote / named barrier / shared-memory or L2
// pointer-chase step costs.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false tools/ubench.cu -o /tmp/ubench
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_shfl_chain(float *out, long long *cyc, int n)
{
    float v = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) v = __shfl_up_sync(0xffffffffu, v, 1) + 1.f;
    long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_vote_chain(int *out, long long *cyc, int n)
{
    int v = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        if (__any_sync(0xffffffffu, v == -1)) v += 7;
        v += 1;
    }
    long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_lds_chase(int *out, long long *cyc, int n)
{
    __shared__ signed char tab[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) tab[i] = (signed char) ((i * 7) % 3 - 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int x = 100;
        long long t0 = clock64();
        for (int i = 0; i < n; ++i) x = (x + tab[x & 4095] + 64) & 4095;
        long long t1 = clock64();
        out[0] = x;
        cyc[0] = t1 - t0;
    }
}

__global__ void k_l2_chase(const int *tab, int *out, long long *cyc, int n)
{
    if (threadIdx.x == 0) {
        int x = 0;
        long long t0 = clock64();
        for (int i = 0; i < n; ++i) x = __ldcg(tab + x);
        long long t1 = clock64();
        out[0] = x;
        cyc[0] = t1 - t0;
    }
}

// named barrier round trip: nw warps, each iteration = STS + bar.sync + LDS of the neighbour's value
__global__ void k_bar(float *out, long long *cyc, int n)
{
    __shared__ float buf[2][1024];
    float v = threadIdx.x;
    const int nthr = blockDim.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        buf[i & 1][threadIdx.x] = v;
        asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
        v = buf[i & 1][(threadIdx.x + 33) % nthr] + 1.f;
    }
    long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// The planned row body of the band DP: 4 cells per lane, values-only chain (shuffle -> 3-input min -> add), near test
// accumulated as an integer minimum, operands (en, old m) from shared memory, result to global; one vote per VR rows.
template <int VR>
__global__ void k_row_body(const float *en_g, float *m_g, long long *cyc, int rows, int pitch)
{
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    float *es = sm + (size_t) warp * 2 * 16 * 128; // [16][128] en, [16][128] old m per warp (reused every 16 rows)
    float *os = es + 16 * 128;
    for (int i = lane; i < 16 * 128; i += 32) {
        es[i] = en_g[(i * 13 + warp) % 4096];
        os[i] = en_g[(i * 7 + warp) % 4096] * 100.f;
    }
    __syncthreads();
    float mp[4] = {1.f + lane, 2.f, 3.f, 4.f};
    const float inf = __int_as_float(0x7f800000);
    const float leftfloor = lane == 0 ? inf : -inf;
    unsigned go = warp * 128 + 4 * lane;
    int redo = 0;
    long long t0 = clock64();
    for (int r0 = 0; r0 < rows; r0 += VR) {
        unsigned umin = 0xffffffffu;
#pragma unroll
        for (int k = 0; k < VR; ++k) {
            const int r = (r0 + k) & 15;
            const float4 e4 = *reinterpret_cast<const float4 *>(es + r * 128 + 4 * lane);
            const float4 o4 = *reinterpret_cast<const float4 *>(os + r * 128 + 4 * lane);
            const float l = fmaxf(__shfl_up_sync(0xffffffffu, mp[3], 1), leftfloor);
            const float rr = __shfl_down_sync(0xffffffffu, mp[0], 1);
            float nv[4];
            nv[0] = __fadd_rn(e4.x, fminf(fminf(l, mp[0]), mp[1]));
            nv[1] = __fadd_rn(e4.y, fminf(fminf(mp[0], mp[1]), mp[2]));
            nv[2] = __fadd_rn(e4.z, fminf(fminf(mp[1], mp[2]), mp[3]));
            nv[3] = __fadd_rn(e4.w, fminf(fminf(mp[2], mp[3]), rr));
            const unsigned u0 = (unsigned) __float_as_int(__fsub_rn(o4.x, nv[0])) * 2u - 2u;
            const unsigned u1 = (unsigned) __float_as_int(__fsub_rn(o4.y, nv[1])) * 2u - 2u;
            const unsigned u2 = (unsigned) __float_as_int(__fsub_rn(o4.z, nv[2])) * 2u - 2u;
            const unsigned u3 = (unsigned) __float_as_int(__fsub_rn(o4.w, nv[3])) * 2u - 2u;
            umin = min(min(umin, u0), min(min(u1, u2), u3));
            *reinterpret_cast<float4 *>(m_g + go) = make_float4(nv[0], nv[1], nv[2], nv[3]);
            go += pitch;
#pragma unroll
            for (int i = 0; i < 4; ++i) mp[i] = nv[i];
        }
        if (__any_sync(0xffffffffu, umin <= 2u * 0x3727C5ACu - 2u)) ++redo;
    }
    long long t1 = clock64();
    if (lane == 0) cyc[warp] = t1 - t0;
    if (threadIdx.x == 0) cyc[nw] = redo;
    m_g[go + 1] = mp[0] + mp[1] + mp[2] + mp[3];
}

// the same row body with the near vote taken every row but consumed one row late (off the chain)
__global__ void k_row_body_lag(const float *en_g, float *m_g, long long *cyc, int rows, int pitch)
{
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    float *es = sm + (size_t) warp * 2 * 16 * 128;
    float *os = es + 16 * 128;
    for (int i = lane; i < 16 * 128; i += 32) {
        es[i] = en_g[(i * 13 + warp) % 4096];
        os[i] = en_g[(i * 7 + warp) % 4096] * 100.f;
    }
    __syncthreads();
    float mp[4] = {1.f + lane, 2.f, 3.f, 4.f};
    const float inf = __int_as_float(0x7f800000);
    const float leftfloor = lane == 0 ? inf : -inf;
    unsigned go = warp * 128 + 4 * lane;
    int redo = 0;
    bool pend = false;
    long long t0 = clock64();
    for (int r0 = 0; r0 < rows; r0 += 16) {
#pragma unroll 4
        for (int r = 0; r < 16; ++r) {
            const float4 e4 = *reinterpret_cast<const float4 *>(es + r * 128 + 4 * lane);
            const float4 o4 = *reinterpret_cast<const float4 *>(os + r * 128 + 4 * lane);
            const float l = fmaxf(__shfl_up_sync(0xffffffffu, mp[3], 1), leftfloor);
            const float rr = __shfl_down_sync(0xffffffffu, mp[0], 1);
            float nv[4];
            nv[0] = __fadd_rn(e4.x, fminf(fminf(l, mp[0]), mp[1]));
            nv[1] = __fadd_rn(e4.y, fminf(fminf(mp[0], mp[1]), mp[2]));
            nv[2] = __fadd_rn(e4.z, fminf(fminf(mp[1], mp[2]), mp[3]));
            nv[3] = __fadd_rn(e4.w, fminf(fminf(mp[2], mp[3]), rr));
            const unsigned u0 = (unsigned) __float_as_int(__fsub_rn(o4.x, nv[0])) * 2u - 2u;
            const unsigned u1 = (unsigned) __float_as_int(__fsub_rn(o4.y, nv[1])) * 2u - 2u;
            const unsigned u2 = (unsigned) __float_as_int(__fsub_rn(o4.z, nv[2])) * 2u - 2u;
            const unsigned u3 = (unsigned) __float_as_int(__fsub_rn(o4.w, nv[3])) * 2u - 2u;
            const unsigned key = min(min(u0, u1), min(u2, u3));
            if (__any_sync(0xffffffffu, pend)) {
                ++redo;
                nv[0] += 1.f;
            }
            *reinterpret_cast<float4 *>(m_g + go) = make_float4(nv[0], nv[1], nv[2], nv[3]);
            go += pitch;
#pragma unroll
            for (int i = 0; i < 4; ++i) mp[i] = nv[i];
            pend = key <= 2u * 0x3727C5ACu - 2u;
        }
    }
    long long t1 = clock64();
    if (lane == 0) cyc[warp] = t1 - t0;
    if (threadIdx.x == 0) cyc[nw] = redo;
    m_g[go + 1] = mp[0] + mp[1] + mp[2] + mp[3];
}

// Software-pipelined variant: the near test of row r-1 is issued in the shadow of row r's shuffles (it only needs row
// r-1's values and old values), its vote is consumed at the end of row r.  PACKED: additions as add.rn.f32x2 (sm_100).
template <bool PACKED>
__global__ void k_row_body_pipe(const float *en_g, float *m_g, long long *cyc, int rows, int pitch)
{
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    float *es = sm + (size_t) warp * 2 * 16 * 128;
    float *os = es + 16 * 128;
    for (int i = lane; i < 16 * 128; i += 32) {
        es[i] = en_g[(i * 13 + warp) % 4096];
        os[i] = en_g[(i * 7 + warp) % 4096] * 100.f;
    }
    __syncthreads();
    float mp[4] = {1.f + lane, 2.f, 3.f, 4.f};
    float4 po = make_float4(0.f, 0.f, 0.f, 0.f); // old values of row r-1
    const float inf = __int_as_float(0x7f800000);
    const float leftfloor = lane == 0 ? inf : -inf;
    unsigned go = warp * 128 + 4 * lane;
    int redo = 0;
    long long t0 = clock64();
    for (int r0 = 0; r0 < rows; r0 += 16) {
#pragma unroll 4
        for (int r = 0; r < 16; ++r) {
            const float l0 = __shfl_up_sync(0xffffffffu, mp[3], 1);
            const float rr = __shfl_down_sync(0xffffffffu, mp[0], 1);
            const float4 e4 = *reinterpret_cast<const float4 *>(es + r * 128 + 4 * lane);
            const float4 o4 = *reinterpret_cast<const float4 *>(os + r * 128 + 4 * lane);
            // near test of the previous row, in the shadow of the shuffles
            unsigned u0, u1, u2, u3;
            if (PACKED) {
                const float2 d01 = __fadd2_rn(make_float2(mp[0], mp[1]), make_float2(-po.x, -po.y));
                const float2 d23 = __fadd2_rn(make_float2(mp[2], mp[3]), make_float2(-po.z, -po.w));
                u0 = __float_as_uint(d01.x) * 2u - 2u, u1 = __float_as_uint(d01.y) * 2u - 2u;
                u2 = __float_as_uint(d23.x) * 2u - 2u, u3 = __float_as_uint(d23.y) * 2u - 2u;
            } else {
                u0 = __float_as_uint(__fsub_rn(po.x, mp[0])) * 2u - 2u, u1 = __float_as_uint(__fsub_rn(po.y, mp[1])) * 2u - 2u;
                u2 = __float_as_uint(__fsub_rn(po.z, mp[2])) * 2u - 2u, u3 = __float_as_uint(__fsub_rn(po.w, mp[3])) * 2u - 2u;
            }
            const bool pend = min(min(u0, u1), min(u2, u3)) <= 2u * 0x3727C5ACu - 2u;
            const bool any = __any_sync(0xffffffffu, pend);
            const float l = fmaxf(l0, leftfloor);
            float nv[4];
            const float b0 = fminf(fminf(l, mp[0]), mp[1]), b1 = fminf(fminf(mp[0], mp[1]), mp[2]);
            const float b2 = fminf(fminf(mp[1], mp[2]), mp[3]), b3 = fminf(fminf(mp[2], mp[3]), rr);
            if (PACKED) {
                const float2 a01 = __fadd2_rn(make_float2(e4.x, e4.y), make_float2(b0, b1));
                const float2 a23 = __fadd2_rn(make_float2(e4.z, e4.w), make_float2(b2, b3));
                nv[0] = a01.x, nv[1] = a01.y, nv[2] = a23.x, nv[3] = a23.y;
            } else {
                nv[0] = __fadd_rn(e4.x, b0), nv[1] = __fadd_rn(e4.y, b1), nv[2] = __fadd_rn(e4.z, b2), nv[3] = __fadd_rn(e4.w, b3);
            }
            if (any) {
                ++redo;
                nv[0] += 1.f;
            }
            *reinterpret_cast<float4 *>(m_g + go) = make_float4(nv[0], nv[1], nv[2], nv[3]);
            go += pitch;
#pragma unroll
            for (int i = 0; i < 4; ++i) mp[i] = nv[i];
            po = o4;
        }
    }
    long long t1 = clock64();
    if (lane == 0) cyc[warp] = t1 - t0;
    if (threadIdx.x == 0) cyc[nw] = redo;
    m_g[go + 1] = mp[0] + mp[1] + mp[2] + mp[3];
}

// near key on the fma pipe: bits(d) * 2 - 2 as an integer multiply-add (the compiler turns the C expression into an
// IADD3, which shares the alu pipe with the 3-input minima)
__device__ __forceinline__ unsigned nearkey(float d)
{
    unsigned r;
    asm("mad.lo.u32 %0, %1, 2, 0xfffffffe;" : "=r"(r) : "r"(__float_as_uint(d)));
    return r;
}

// Variant: TWO rows per shuffle round.  Each lane also computes the cells one column left and right of its own four on
// the first row of a pair (from two shuffled values per side), so the second row needs no shuffle: the loop-carried
// chain is shuffle + 2 x (min3 + add) per two rows.
template <int VR, bool IMADKEY>
__global__ void k_row_body2(const float *en_g, float *m_g, long long *cyc, int rows, int pitch)
{
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    float *es = sm + (size_t) warp * 2 * 16 * 136; // en rows with one spare float4 on each side: [16][136]
    float *os = es + 16 * 136;
    for (int i = lane; i < 16 * 136; i += 32) {
        es[i] = en_g[(i * 13 + warp) % 4096];
        os[i] = en_g[(i * 7 + warp) % 4096] * 100.f;
    }
    __syncthreads();
    float mp[4] = {1.f + lane, 2.f, 3.f, 4.f};
    const float inf = __int_as_float(0x7f800000);
    const float leftfloor = lane == 0 ? inf : -inf;
    unsigned go = warp * 128 + 4 * lane;
    int redo = 0;
    long long t0 = clock64();
    for (int r0 = 0; r0 < rows; r0 += VR) {
        unsigned umin = 0xffffffffu;
#pragma unroll
        for (int k = 0; k < VR; k += 2) {
            const int r = (r0 + k) & 15;
            const float *e = es + r * 136 + 4 + 4 * lane;
            const float *o = os + r * 136 + 4 + 4 * lane;
            const float4 e4 = *reinterpret_cast<const float4 *>(e);
            const float el = e[-1], er = e[4];
            const float4 o4 = *reinterpret_cast<const float4 *>(o);
            const float4 f4 = *reinterpret_cast<const float4 *>(e + 136);
            const float4 q4 = *reinterpret_cast<const float4 *>(o + 136);
            const float l2 = fmaxf(__shfl_up_sync(0xffffffffu, mp[2], 1), leftfloor);
            const float l3 = fmaxf(__shfl_up_sync(0xffffffffu, mp[3], 1), leftfloor);
            const float r0v = __shfl_down_sync(0xffffffffu, mp[0], 1);
            const float r1v = __shfl_down_sync(0xffffffffu, mp[1], 1);
            float a[6], b[4];
            a[0] = fmaxf(__fadd_rn(el, fminf(fminf(l2, l3), mp[0])), leftfloor);
            a[1] = __fadd_rn(e4.x, fminf(fminf(l3, mp[0]), mp[1]));
            a[2] = __fadd_rn(e4.y, fminf(fminf(mp[0], mp[1]), mp[2]));
            a[3] = __fadd_rn(e4.z, fminf(fminf(mp[1], mp[2]), mp[3]));
            a[4] = __fadd_rn(e4.w, fminf(fminf(mp[2], mp[3]), r0v));
            a[5] = __fadd_rn(er, fminf(fminf(mp[3], r0v), r1v));
            b[0] = __fadd_rn(f4.x, fminf(fminf(a[0], a[1]), a[2]));
            b[1] = __fadd_rn(f4.y, fminf(fminf(a[1], a[2]), a[3]));
            b[2] = __fadd_rn(f4.z, fminf(fminf(a[2], a[3]), a[4]));
            b[3] = __fadd_rn(f4.w, fminf(fminf(a[3], a[4]), a[5]));
            unsigned u[8];
            const float d0 = __fsub_rn(o4.x, a[1]), d1 = __fsub_rn(o4.y, a[2]), d2 = __fsub_rn(o4.z, a[3]), d3 = __fsub_rn(o4.w, a[4]);
            const float d4 = __fsub_rn(q4.x, b[0]), d5 = __fsub_rn(q4.y, b[1]), d6 = __fsub_rn(q4.z, b[2]), d7 = __fsub_rn(q4.w, b[3]);
            if (IMADKEY) {
                u[0] = nearkey(d0), u[1] = nearkey(d1), u[2] = nearkey(d2), u[3] = nearkey(d3);
                u[4] = nearkey(d4), u[5] = nearkey(d5), u[6] = nearkey(d6), u[7] = nearkey(d7);
            } else {
                u[0] = __float_as_uint(d0) * 2u - 2u, u[1] = __float_as_uint(d1) * 2u - 2u, u[2] = __float_as_uint(d2) * 2u - 2u, u[3] = __float_as_uint(d3) * 2u - 2u;
                u[4] = __float_as_uint(d4) * 2u - 2u, u[5] = __float_as_uint(d5) * 2u - 2u, u[6] = __float_as_uint(d6) * 2u - 2u, u[7] = __float_as_uint(d7) * 2u - 2u;
            }
            umin = min(min(min(umin, u[0]), min(u[1], u[2])), min(min(u[3], u[4]), min(min(u[5], u[6]), u[7])));
            *reinterpret_cast<float4 *>(m_g + go) = make_float4(a[1], a[2], a[3], a[4]);
            *reinterpret_cast<float4 *>(m_g + go + pitch) = make_float4(b[0], b[1], b[2], b[3]);
            go += 2 * pitch;
#pragma unroll
            for (int i = 0; i < 4; ++i) mp[i] = b[i];
        }
        if (__any_sync(0xffffffffu, umin <= 2u * 0x3727C5ACu - 2u)) ++redo;
    }
    long long t1 = clock64();
    if (lane == 0) cyc[warp] = t1 - t0;
    if (threadIdx.x == 0) cyc[nw] = redo;
    m_g[go + 1] = mp[0] + mp[1] + mp[2] + mp[3];
}

template <int VR, bool IMADKEY>
int run_body2(const char *name, float *d_f, long long *d_c, int rows, int pitch)
{
    long long h_c[64];
    CHECK(cudaFuncSetAttribute(k_row_body2<VR, IMADKEY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int nw : {1, 3, 4, 8}) {
        const size_t smem = (size_t) nw * 2 * 16 * 136 * 4;
        for (int rep = 0; rep < 2; ++rep) k_row_body2<VR, IMADKEY><<<1, nw * 32, smem>>>(d_f, d_f + (8 << 20), d_c, rows, pitch);
        CHECK(cudaMemcpy(h_c, d_c, sizeof h_c, cudaMemcpyDeviceToHost));
        printf("%s, %2d warps: %.1f cycles/row\n", name, nw, (double) h_c[0] / rows);
    }
    return 0;
}

int main()
{
    float *d_f;
    int *d_i;
    long long *d_c, h_c[64];
    CHECK(cudaMalloc(&d_f, 64 << 20));
    CHECK(cudaMalloc(&d_i, 64 << 20));
    CHECK(cudaMalloc(&d_c, sizeof h_c));
    CHECK(cudaMemset(d_f, 0, 64 << 20));
    const int n = 4096;
    for (int rep = 0; rep < 2; ++rep) {
        k_shfl_chain<<<1, 32>>>(d_f, d_c, n);
        CHECK(cudaMemcpy(h_c, d_c, sizeof h_c, cudaMemcpyDeviceToHost));
        if (rep) printf("shfl_up + fadd chain      : %.1f cycles/step\n", (double) h_c[0] / n);
        k_vote_chain<<<1, 32>>>(d_i, d_c, n);
        CHECK(cudaMemcpy(h_c, d_c, sizeof h_c, cudaMemcpyDeviceToHost));
        if (rep) printf("vote.any + branch chain   : %.1f cycles/step\n", (double) h_c[0] / n);
        k_lds_chase<<<1, 128>>>(d_i, d_c, n);
        CHECK(cudaMemcpy(h_c, d_c, sizeof h_c, cudaMemcpyDeviceToHost));
        if (rep) printf("lds.s8 chase              : %.1f cycles/step\n", (double) h_c[0] / n);
    }
    {
        // pointer chase through 32 MB (L2-resident after the first pass), stride ~ 4 KB
        const int cnt = 8 << 20;
        int *h = (int *) malloc((size_t) cnt * 4);
        for (int i = 0; i < cnt; ++i) h[i] = (int) (((long long) i + 1031 * 1024 + 17) % cnt);
        CHECK(cudaMemcpy(d_i, h, (size_t) cnt * 4, cudaMemcpyHostToDevice));
        for (int rep = 0; rep < 2; ++rep) {
            k_l2_chase<<<1, 32>>>(d_i, (int *) d_f, d_c, 2048);
            CHECK(cudaMemcpy(h_c, d_c, sizeof h_c, cudaMemcpyDeviceToHost));
            printf("global (L2) chase pass %d   : %.1f cycles/step\n", rep, (double) h_c[0] / 2048);
        }
        free(h);
    }
    for (int nw : {2, 4, 8, 13}) {
        for (int rep = 0; rep < 2; ++rep) k_bar<<<1, nw * 32>>>(d_f, d_c, n);
        CHECK(cudaMemcpy(h_c, d_c, sizeof h_c, cudaMemcpyDeviceToHost));
        printf("sts + bar.sync + lds, %2d warps : %.1f cycles/round\n", nw, (double) h_c[0] / n);
    }
    CHECK(cudaFuncSetAttribute(k_row_body<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CHECK(cudaFuncSetAttribute(k_row_body<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CHECK(cudaFuncSetAttribute(k_row_body<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CHECK(cudaFuncSetAttribute(k_row_body<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int rows = 2048, pitch = 3856;
    for (int nw : {1, 2, 3, 4, 8, 12}) {
        const size_t smem = (size_t) nw * 2 * 16 * 128 * 4;
        for (int rep = 0; rep < 2; ++rep) k_row_body<4><<<1, nw * 32, smem>>>(d_f, d_f + (8 << 20), d_c, rows, pitch);
        CHECK(cudaMemcpy(h_c, d_c, sizeof h_c, cudaMemcpyDeviceToHost));
        printf("row body, vote/4 rows, %2d warps: %.1f cycles/row (warp 0), %.1f (last warp)\n", nw, (double) h_c[0] / rows, (double) h_c[nw - 1] / rows);
    }
    {
        const size_t smem = (size_t) 1 * 2 * 16 * 128 * 4;
        for (int rep = 0; rep < 2; ++rep) k_row_body<2><<<1, 32, smem>>>(d_f, d_f + (8 << 20), d_c, rows, pitch);
        CHECK(cudaMemcpy(h_c, d_c, sizeof h_c, cudaMemcpyDeviceToHost));
        printf("row body, vote/2 rows,  1 warp : %.1f cycles/row\n", (double) h_c[0] / rows);
        for (int rep = 0; rep < 2; ++rep) k_row_body<8><<<1, 32, smem>>>(d_f, d_f + (8 << 20), d_c, rows, pitch);
        CHECK(cudaMemcpy(h_c, d_c, sizeof h_c, cudaMemcpyDeviceToHost));
        printf("row body, vote/8 rows,  1 warp : %.1f cycles/row\n", (double) h_c[0] / rows);
        for (int rep = 0; rep < 2; ++rep) k_row_body<16><<<1, 32, smem>>>(d_f, d_f + (8 << 20), d_c, rows, pitch);
        CHECK(cudaMemcpy(h_c, d_c, sizeof h_c, cudaMemcpyDeviceToHost));
        printf("row body, vote/16 rows, 1 warp : %.1f cycles/row\n", (double) h_c[0] / rows);
    }
    for (int nw : {1, 3, 4}) {
        const size_t smem = (size_t) nw * 2 * 16 * 128 * 4;
        for (int rep = 0; rep < 2; ++rep) k_row_body_lag<<<1, nw * 32, smem>>>(d_f, d_f + (8 << 20), d_c, rows, pitch);
        CHECK(cudaMemcpy(h_c, d_c, sizeof h_c, cudaMemcpyDeviceToHost));
        printf("row body, lagged vote every row, %2d warps: %.1f cycles/row (redo %lld)\n", nw, (double) h_c[0] / rows, h_c[nw]);
    }
    for (int nw : {1, 3, 4}) {
        const size_t smem = (size_t) nw * 2 * 16 * 128 * 4;
        for (int rep = 0; rep < 2; ++rep) k_row_body_pipe<false><<<1, nw * 32, smem>>>(d_f, d_f + (8 << 20), d_c, rows, pitch);
        CHECK(cudaMemcpy(h_c, d_c, sizeof h_c, cudaMemcpyDeviceToHost));
        printf("row body, pipelined near test, %2d warps: %.1f cycles/row (redo %lld)\n", nw, (double) h_c[0] / rows, h_c[nw]);
        for (int rep = 0; rep < 2; ++rep) k_row_body_pipe<true><<<1, nw * 32, smem>>>(d_f, d_f + (8 << 20), d_c, rows, pitch);
        CHECK(cudaMemcpy(h_c, d_c, sizeof h_c, cudaMemcpyDeviceToHost));
        printf("row body, pipelined near test, f32x2, %2d warps: %.1f cycles/row (redo %lld)\n", nw, (double) h_c[0] / rows, h_c[nw]);
    }
    run_body2<8, false>("2-row steps, vote/8 ", d_f, d_c, rows, pitch);
    run_body2<8, true>("2-row steps, vote/8, imad key", d_f, d_c, rows, pitch);
    run_body2<16, true>("2-row steps, vote/16, imad key", d_f, d_c, rows, pitch);
    run_body2<4, true>("2-row steps, vote/4, imad key", d_f, d_c, rows, pitch);
    CHECK(cudaDeviceSynchronize());
    return 0;
}
