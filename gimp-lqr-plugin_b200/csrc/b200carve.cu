// b200carve.cu -- host side of the CUDA seam-carving engine: device state per carver handle, kernel
// sequencing of the per-seam loop, and the C ABI of include/b200carve.h.
//
// State model follows liblqr's: reference size (w_start,h_start), current size
// (w,h), allocated size (w0,h0), level / max_level, the raw index table (x-th visible pixel of row y),
// the visibility map vs, energy en, cumulative map m and parent map least -- all resident in HBM for the
// life of the handle.  The host only sequences kernels; no pixel arithmetic happens on the CPU.
#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "b200carve.h"
#include "carver_kernels.cuh"
#include "band_dp.cuh"
#include "band_tail.cuh"
#include "seam_path.cuh"
#include "seam_trace.cuh"
#include "mmap_full_strips.cuh"
#include "mmap_cluster.cuh"

using namespace b200c;

namespace {

thread_local std::string g_err;
std::atomic<long> g_launches{0};
std::atomic<unsigned long long> g_update_cells{0};

int fail(int code, const char *what, cudaError_t e = cudaSuccess)
{
    char buf[512];
    if (e != cudaSuccess)
        snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    else
        snprintf(buf, sizeof buf, "%s", what);
    g_err = buf;
    return code;
}

#define CU_TRY(expr)                                                                          \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return fail(e__ == cudaErrorMemoryAllocation ? B200C_NOMEM : B200C_ERROR, #expr, e__); \
    } while (0)
#define B_TRY(expr)                      \
    do {                                 \
        int r__ = (expr);                \
        if (r__ != B200C_OK) return r__; \
    } while (0)

// ---- optional per-stage timing (B200C_TIMING=1): CUDA events around every launch of a stage ----------
struct StageStat {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    double ms = 0;
    long launches = 0;
};
std::mutex g_stage_mu;

// host-side wall time per phase of the API calls, summed over all threads (b200c_hostprof_ms; diagnostics only)
std::atomic<long long> g_hostprof_ns[16];
struct HostScope {
    int i;
    std::chrono::steady_clock::time_point t0;
    explicit HostScope(int idx) : i(idx), t0(std::chrono::steady_clock::now()) {}
    ~HostScope() { g_hostprof_ns[i] += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(); }
};

std::map<std::string, StageStat> g_stages;
bool g_timing = false;

struct StageScope {
    const char *name;
    cudaStream_t s;
    cudaEvent_t a = nullptr, b = nullptr;
    StageScope(const char *n, cudaStream_t st, int launches = 1) : name(n), s(st)
    {
        g_launches += launches;
        if (!g_timing) return;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, s);
    }
    ~StageScope()
    {
        if (!g_timing) return;
        cudaEventRecord(b, s);
        std::lock_guard<std::mutex> lk(g_stage_mu);
        auto &st = g_stages[name];
        st.pending.emplace_back(a, b);
        st.launches++;
    }
};

void drain_stage(StageStat &st)
{
    for (auto &pr : st.pending) {
        cudaEventSynchronize(pr.second);
        float ms = 0;
        if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) st.ms += ms;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    st.pending.clear();
}

int g_device_tls_default = 0;
thread_local int g_device = -1;
thread_local cudaStream_t g_ext_stream = nullptr;
thread_local bool g_use_ext_stream = false;

// ---- lanes: a stream, the upload events and the instantiated per-seam graphs, pooled per process.  Creating and
// destroying these for every image goes through the driver's global lock; with a dozen carvers in flight on host
// threads that cost tens of ms per image -- more than the seams.  A carver borrows a lane for its
// lifetime; graph executables are re-targeted at the next carver's buffers with cudaGraphExecUpdate.
struct LaneGraph {
    cudaGraph_t graph = nullptr; // owns the nodes whose handles address the executable's nodes
    cudaGraphExec_t exec = nullptr;
    cudaGraphNode_t node[10] = {};
    int n = 0;
};
void lane_graph_reset(LaneGraph *g)
{
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    *g = LaneGraph();
}
struct Lane {
    int device = 0;
    bool pooled = false; // owns its stream and goes back to the pool
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaEvent_t done = nullptr; // blocking-sync event: see carver_sync()
    cudaEvent_t prog[4] = {nullptr, nullptr, nullptr, nullptr}; // progress points of a build session (build_vsmap)
    cudaEvent_t rd[4] = {nullptr, nullptr, nullptr, nullptr};   // arrival of the chunks of a read-out (b200c_carver_readout)
    int *seams_h = nullptr;     // mapped pinned word: seams of the running session the device has completed
    std::map<int, LaneGraph> graphs; // per kernel set of the per-seam loop (graph_key)
};
std::mutex g_lane_mu;
std::vector<Lane *> g_lane_free;
std::atomic<int> g_lanes_busy{0}; // carvers alive in this process
constexpr size_t kLaneKeep = 64;

void lane_destroy(Lane *l)
{
    if (!l) return;
    for (auto &kv : l->graphs) lane_graph_reset(&kv.second);
    for (cudaEvent_t e : l->ev)
        if (e) cudaEventDestroy(e);
    if (l->done) cudaEventDestroy(l->done);
    for (cudaEvent_t e : l->prog)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : l->rd)
        if (e) cudaEventDestroy(e);
    if (l->seams_h) cudaFreeHost(l->seams_h);
    if (l->pooled && l->stream) cudaStreamDestroy(l->stream);
    delete l;
}

// ext != nullptr: a private lane around the caller's stream (b200c_set_stream)
Lane *lane_acquire(int device, bool use_ext, cudaStream_t ext)
{
    ++g_lanes_busy;
    {
        // a lane around a caller's stream is kept too (with its graph executables), and reused for that stream only
        std::lock_guard<std::mutex> lk(g_lane_mu);
        for (size_t i = 0; i < g_lane_free.size(); ++i) {
            Lane *l = g_lane_free[i];
            if (l->device == device && (use_ext ? (!l->pooled && l->stream == ext) : l->pooled)) {
                g_lane_free.erase(g_lane_free.begin() + i);
                return l;
            }
        }
    }
    Lane *l = new (std::nothrow) Lane();
    if (!l) {
        --g_lanes_busy;
        return nullptr;
    }
    l->device = device;
    l->pooled = !use_ext;
    l->stream = ext;
    bool ok = use_ext || cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i) ok = cudaEventCreateWithFlags(&l->ev[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&l->done, cudaEventDisableTiming | cudaEventBlockingSync) == cudaSuccess;
    for (int i = 0; i < 4 && ok; ++i) ok = cudaEventCreateWithFlags(&l->prog[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 4 && ok; ++i) ok = cudaEventCreateWithFlags(&l->rd[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaHostAlloc((void **) &l->seams_h, 64, cudaHostAllocMapped) == cudaSuccess;
    if (ok) *l->seams_h = 0;
    if (!ok) {
        --g_lanes_busy;
        lane_destroy(l);
        return nullptr;
    }
    return l;
}

// the lane's stream must be idle
void lane_release(Lane *l)
{
    if (!l) return;
    --g_lanes_busy;
    Lane *evict = l;
    {
        std::lock_guard<std::mutex> lk(g_lane_mu);
        if (g_lane_free.size() >= kLaneKeep) // full: the oldest lane around a caller's stream makes room
            for (size_t i = 0; i < g_lane_free.size(); ++i)
                if (!g_lane_free[i]->pooled) {
                    evict = g_lane_free[i];
                    g_lane_free.erase(g_lane_free.begin() + i);
                    break;
                }
        if (g_lane_free.size() < kLaneKeep) g_lane_free.push_back(l);
        if (evict == l && g_lane_free.back() == l) evict = nullptr;
    }
    lane_destroy(evict);
}

// With many carvers in flight (a batch host: one thread per image), threads that spin in cudaStreamSynchronize get in
// the way of the ones that have work to enqueue (measured: cudaMemcpyAsync taking ms); waits then sleep on a blocking
// event instead, and the upload helpers stay out of the way.  A lone image (and its attached carvers) keeps the
// spinning wait -- lowest latency -- and the helpers.
bool host_crowded()
{
    static const int limit = [] {
        const char *e = getenv("B200C_CROWDED"); // carvers alive from which waits sleep (default: 4)
        return e && atoi(e) > 0 ? atoi(e) : 4;
    }();
    return g_lanes_busy.load(std::memory_order_relaxed) > limit;
}
bool host_shared() { return g_lanes_busy.load(std::memory_order_relaxed) > 2; }

} // namespace

struct B200Carver {
    int device = 0;
    cudaStream_t stream = nullptr;
    int w = 0, h = 0, w0 = 0, h0 = 0, w_start = 0, h_start = 0;
    int level = 1, max_level = 1;
    int channels = 0, alpha = -1, transposed = 0;
    bool active = false, nrg_active = false, nrg_uptodate = false;
    bool raw_ident = false; // the index table is the identity since init_raw (no carve, no inflate yet)
    B200Carver *root = nullptr;
    std::vector<B200Carver *> attached;

    uint8_t *rgb = nullptr;
    int *vs = nullptr; // owned by the root; attached carvers alias it
    float *bias = nullptr, *rigmask = nullptr;   // physical (pixel id)
    int *raw = nullptr;                          // index table: x-th visible pixel of row y
    float *en = nullptr, *m = nullptr, *rig = nullptr; // compact maps [h_start][pitch] (DevP)
    int8_t *pdx = nullptr;                       // compact parent offsets
    int8_t *jump = nullptr;                      // block jump tables of the backtrack (seam_trace.cuh)
    int pitch = 0;
    int *vpath_x = nullptr, *nrg_xmin = nullptr, *nrg_xmax = nullptr;
    unsigned *nrg_pack = nullptr;
    int *dyn_d = nullptr;                     // device seam counter of the running build session (DevP::dyn)
    int w_epoch = 0, vs_epoch = 0;            // width / visibility level at the session's first seam
    Lane *lane = nullptr;                     // stream, upload events and per-seam graph executables (pooled)
    bool graph_fresh[2] = {false, false};     // the lane's graph for this leftright value points at this session
    LaneGraph *graph[2] = {nullptr, nullptr};
    bool use_graph = true;                    // B200C_GRAPH=0: launch the kernels one by one
    bool use_pdl = false;                     // B200C_PDL=1: programmatic dependent launch between the nodes of the seam graph
    int4 *fix_d = nullptr;                    // band-DP chunk table for k_fix_parents (+ its count)
    int *fixn_d = nullptr;
    int *tail_d = nullptr;                    // band DP -> tail kernel hand-over (DevP::tail)
    int *far_d = nullptr;                     // rows whose FAR carve phase is done in this session (DevP::far)
    bool use_split = false;                   // B200C_SPLIT=1: the carve as NEAR + FAR launches, FAR beside the band DP (measured
                                              // SLOWER on a B200: the FAR CTAs share the band DP's SM and slow its one critical warp)
    bool use_tail = true;                     // B200C_TAIL=0: the band kernel keeps its in-CTA wide-window loop
    bool use_trace = true;                    // B200C_TRACE=0: the single-CTA staged backtrack (seam_path.cuh)
    bool use_cluster = true;                  // B200C_CLUSTER=0: the full DP as h/32 strip launches (mmap_full_strips.cuh)
    // batch session (b200c_batch_build_maps): this carver LEADS, its launches advance the mates too (image = blockIdx.z)
    std::vector<B200Carver *> mates;
    DevP *tab_d = nullptr;                    // [2][n]: per-seam argument blocks, then the full-pass ones
    BdMaps *mtab_d = nullptr;                 // [n] tensor maps
    int tab_n = 0;
    alignas(64) BdMaps maps;                  // TMA tensor maps over the compact arrays (band DP)
    int *err_d = nullptr;                     // device error word (see DevP::err)
    unsigned long long *cells_d = nullptr;    // band cells visited by the incremental DP
    long long *dbg_d = nullptr;               // role cycle counters (B200C_DBG=1)
    bool generic = false;                     // B200C_GENERIC=1: only the generic single-CTA kernels
    int bd_maxseg = 1 << 20;                  // B200C_BD_MAXSEG: test knob, forces the band DP's wide-window path

    float rigidity = 0.f;
    int delta_x = 1;
    std::vector<float> rigmap_h; // 2*delta_x+1, centred at [delta_x]
    float *rigmap_d = nullptr;

    int ef = 2, grad_kind = GRAD_XABS, read_kind = READ_BRIGHTNESS, nrg_radius = 1;
    int leftright = 0;
    unsigned lr_freq = 0;

    uint8_t *host_out = nullptr; // pinned read-out staging
    size_t host_out_cap = 0;
    cudaMemPool_t pool = nullptr; // the engine's own device memory pool (engine_pool), or NULL: the default pool
    int rd_rows = 0, rd_chunks = 0, rd_ready = 0; // chunked read-out: rows per chunk, chunks, chunks known to have arrived
};

namespace {

// The engine allocates from a memory pool of its OWN per device (freed blocks stay in it: the per-resize maps are
// reallocated at every inflate / flatten), so the host process's default pool and its release threshold are left alone.
cudaMemPool_t engine_pool(int device)
{
    static std::mutex mu;
    static std::map<int, cudaMemPool_t> pools;
    std::lock_guard<std::mutex> lk(mu);
    auto it = pools.find(device);
    if (it != pools.end()) return it->second;
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    cudaMemPool_t pool = nullptr;
    if (cudaMemPoolCreate(&pool, &props) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    } else {
        cudaGetLastError(); // no private pool: allocate from the device's default pool, untouched
        pool = nullptr;
    }
    pools.emplace(device, pool);
    return pool;
}

cudaError_t pool_alloc(const B200Carver *c, void **p, size_t bytes, cudaStream_t s)
{
    return c->pool ? cudaMallocFromPoolAsync(p, bytes, c->pool, s) : cudaMallocAsync(p, bytes, s);
}

template <class T>
int dalloc(B200Carver *c, T **p, size_t n, bool zero)
{
    *p = nullptr;
    CU_TRY(pool_alloc(c, (void **) p, n * sizeof(T), c->stream));
    if (zero) CU_TRY(cudaMemsetAsync(*p, 0, n * sizeof(T), c->stream));
    return B200C_OK;
}

template <class T>
void dfree(B200Carver *c, T *&p)
{
    if (p) cudaFreeAsync((void *) p, c->stream);
    p = nullptr;
}

int use_device(const B200Carver *c)
{
    CU_TRY(cudaSetDevice(c->device));
    return B200C_OK;
}

// waits for the carver's queue: spinning, or asleep on the lane's blocking event when the host is crowded
cudaError_t carver_sync(const B200Carver *c)
{
    const Lane *l = c->lane ? c->lane : (c->root ? c->root->lane : nullptr);
    if (l && l->done && host_crowded()) {
        const cudaError_t e = cudaEventRecord(l->done, c->stream);
        return e != cudaSuccess ? e : cudaEventSynchronize(l->done);
    }
    return cudaStreamSynchronize(c->stream);
}

int check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(B200C_ERROR, what, e);
    return B200C_OK;
}

bool fast_path(const B200Carver *c);
// the carve runs as NEAR + FAR launches, FAR beside the band DP: a lone image on the tiled band path only
bool split_carve(const B200Carver *c)
{
    return c->use_split && c->mates.empty() && fast_path(c) && c->delta_x <= 4 && c->h <= BD_HMAX && c->far_d;
}

DevP view(const B200Carver *c)
{
    DevP p;
    p.dyn = nullptr;
    p.dyn_host = nullptr;
    p.w = c->w;
    p.h = c->h;
    p.w0 = c->w0;
    p.h0 = c->h0;
    p.w_start = c->w_start;
    p.raw_stride = c->w_start;
    p.pitch = c->pitch;
    p.channels = c->channels;
    p.alpha = c->alpha;
    p.level = c->level;
    p.delta_x = c->delta_x;
    p.leftright = c->leftright;
    p.grad_kind = c->grad_kind;
    p.read_kind = c->read_kind;
    p.nrg_radius = c->nrg_radius;
    p.use_rig = c->rigidity != 0.f;
    p.raw_ident = c->raw_ident && c->w == c->w0 && c->w0 == c->w_start ? 1 : 0;
    p.bd_maxseg = c->bd_maxseg;
    p.rgb = c->rgb;
    p.vs = c->vs;
    p.raw = c->raw;
    p.en = c->en;
    p.m = c->m;
    p.pdx = c->pdx;
    p.jump = (signed char *) c->jump;
    p.rig = c->rig;
    p.bias = c->bias;
    p.rigmask = c->rigmask;
    p.rigmap = c->rigmap_d ? c->rigmap_d + c->delta_x : nullptr;
    p.vpath_x = c->vpath_x;
    p.nrg_xmin = c->nrg_xmin;
    p.nrg_xmax = c->nrg_xmax;
    p.nrg_pack = c->nrg_pack;
    p.fix = c->fix_d;
    p.fixn = c->fixn_d;
    p.tail = c->use_tail ? c->tail_d : nullptr;
    p.far = nullptr;
    p.err = c->err_d;
    p.cells = c->cells_d;
    p.dbg = c->dbg_d;
    return p;
}

// the argument block of the per-seam kernels: the same for every seam of a build session (carver_kernels.cuh seam_view)
DevP view_dyn(const B200Carver *c)
{
    DevP p = view(c);
    p.w = c->w_epoch;
    p.dyn = c->dyn_d;
    p.dyn_host = (c->lane && c->mates.empty()) ? c->lane->seams_h : nullptr;
    p.far = split_carve(c) ? c->far_d : nullptr; // UVA: the mapped word's host address is its device address
    return p;
}

// ---- batch sessions: one launch advances the leader and its mates; the kernels pick their image by blockIdx.z from a
// table of argument blocks in HBM (carver_kernels.cuh pick_image).  A lone carver passes no table.
int batch_n(const B200Carver *c) { return 1 + (int) c->mates.size(); }
const DevP *tab_dyn(const B200Carver *c) { return c->mates.empty() ? nullptr : c->tab_d; }
const DevP *tab_static(const B200Carver *c) { return c->mates.empty() ? nullptr : c->tab_d + c->tab_n; }
const BdMaps *tab_maps(const B200Carver *c) { return c->mates.empty() ? nullptr : c->mtab_d; }

template <class F>
void for_batch(B200Carver *c, F f)
{
    f(c);
    for (B200Carver *m : c->mates) f(m);
}

// (re)writes one half of the table: dyn = the per-seam blocks (view_dyn), else the full-pass blocks (view)
int upload_tab(B200Carver *c, bool dyn)
{
    if (c->mates.empty()) return B200C_OK;
    std::vector<DevP> h;
    for_batch(c, [&](B200Carver *m) { h.push_back(dyn ? view_dyn(m) : view(m)); });
    CU_TRY(cudaMemcpyAsync(c->tab_d + (dyn ? 0 : c->tab_n), h.data(), h.size() * sizeof(DevP), cudaMemcpyHostToDevice, c->stream));
    if (dyn) {
        std::vector<BdMaps> hm;
        for_batch(c, [&](B200Carver *m) { hm.push_back(m->maps); });
        CU_TRY(cudaMemcpyAsync(c->mtab_d, hm.data(), hm.size() * sizeof(BdMaps), cudaMemcpyHostToDevice, c->stream));
    }
    return B200C_OK; // pageable sources: staged by the runtime before the calls return
}

void set_width_one(B200Carver *c, int w1)
{
    c->w = w1;
    c->level = c->w0 - w1 + 1;
}

void set_width_rec(B200Carver *c, int w1)
{
    set_width_one(c, w1);
    for (B200Carver *a : c->attached) set_width_rec(a, w1);
}

void propagate_vs(B200Carver *c)
{
    for (B200Carver *a : c->attached) {
        a->vs = c->vs;
        propagate_vs(a);
    }
}

int upload_rigmap(B200Carver *c)
{
    CU_TRY(cudaMemcpyAsync(c->rigmap_d, c->rigmap_h.data(), c->rigmap_h.size() * sizeof(float),
                           cudaMemcpyHostToDevice, c->stream));
    // the host vector may be rewritten (transpose) before the copy runs: pageable copies are staged
    // synchronously by the runtime, so no extra sync is required here.
    return B200C_OK;
}

int init_raw(B200Carver *c)
{
    dim3 grid((c->w_start + 255) / 256, c->h_start);
    StageScope sc("init_raw", c->stream);
    k_init_raw<<<grid, 256, 0, c->stream>>>(c->raw, c->w_start, c->h_start);
    c->raw_ident = true;
    return check_launch("k_init_raw");
}

// ---- TMA tensor maps over the compact arrays (band_dp.cuh): 2-D, tiled, boxes of BD_BW columns x box_rows rows
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_map(CUtensorMap *tm, void *base, bool bytes, int pitch, int rows, int box_rows)
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn) f;
    });
    if (!fn) return fail(B200C_ERROR, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t esz = bytes ? 1 : 4;
    cuuint64_t dims[2] = {(cuuint64_t) pitch, (cuuint64_t) rows};
    cuuint64_t strides[1] = {(cuuint64_t) pitch * esz};
    cuuint32_t box[2] = {BD_BW, (cuuint32_t) box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(tm, bytes ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[96];
        snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed (CUresult %d)", (int) r);
        return fail(B200C_ERROR, msg);
    }
    return B200C_OK;
}

// The compact maps (DevP): en for every carver that computes energy, m / pdx for an initialised one; sized by
// the reference size, which only flatten and transpose change.  rig is made on demand by build_maps.
void free_maps(B200Carver *c)
{
    dfree(c, c->en);
    dfree(c, c->m);
    dfree(c, c->pdx);
    dfree(c, c->jump);
    dfree(c, c->rig);
    c->nrg_uptodate = false;
}

int alloc_maps(B200Carver *c)
{
    free_maps(c);
    c->pitch = (c->w_start + 4 + 15) / 16 * 16; // >= 4 columns right of the image: +inf sentinels of en / m
    const size_t n = (size_t) c->pitch * c->h_start + 64; // slack: aligned windows may end past the last row
    if (c->nrg_active) B_TRY(dalloc(c, &c->en, n, true));
    if (c->active) {
        B_TRY(dalloc(c, &c->m, n, true));
        B_TRY(dalloc(c, &c->pdx, n, true));
        B_TRY(dalloc(c, &c->jump, st_jump_bytes(c->h_start, c->delta_x <= 4 ? c->delta_x : 4, c->pitch), true));
        const int K = bd_rows(c->delta_x, c->rigidity != 0.f);
        B_TRY(encode_map(&c->maps.m, c->m, false, c->pitch, c->h_start, K + 1));
        B_TRY(encode_map(&c->maps.en, c->en, false, c->pitch, c->h_start, K));
        B_TRY(encode_map(&c->maps.pdx, c->pdx, true, c->pitch, c->h_start, K));
        B_TRY(encode_map(&c->maps.mst, c->m, false, c->pitch, c->h_start, K));
        c->maps.rig = c->maps.en;
    }
    return B200C_OK;
}

int init_energy_related(B200Carver *c)
{
    if (c->active || c->nrg_active) return fail(B200C_ERROR, "init_energy_related: already initialised");
    B_TRY(dalloc(c, &c->raw, (size_t) c->w_start * c->h_start + 16, false)); // slack: 16-byte accesses at the end of the last row
    B_TRY(init_raw(c));
    c->nrg_active = true;
    return alloc_maps(c);
}

// ---- A.3 / A.5 full passes ---------------------------------------------------------------------------
int build_emap(B200Carver *c)
{
    if (c->nrg_uptodate) return B200C_OK;
    B_TRY(upload_tab(c, false));
    dim3 grid((c->pitch + EF_TW - 1) / EF_TW, (c->h + EF_TH - 1) / EF_TH, batch_n(c));
    StageScope sc("energy_full", c->stream);
    k_energy_full<<<grid, 256, 0, c->stream>>>(view(c), tab_static(c));
    B_TRY(check_launch("k_energy_full"));
    for_batch(c, [](B200Carver *m) { m->nrg_uptodate = true; });
    return B200C_OK;
}

bool fast_path(const B200Carver *c) { return !c->generic; }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the device that is current when it is called, and one
// process may drive several GPUs (b200c_set_device): the opt-in is tracked per device.
// It is also LAZY -- only the instances a carver can launch (its delta_x, with / without rigidity; both tie rules) are
// touched: with lazy module loading every cudaFuncSetAttribute pulls that kernel's code onto the device, and a one-shot
// plug-in process (GIMP runs one per non-interactive invocation) should not pay for the instances it never uses.
template <int D, bool RIG>
void raise_smem_limits_dr(cudaError_t &err)
{
    auto set = [&err](const void *fn, size_t bytes) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes);
        if (e != cudaSuccess && err == cudaSuccess) err = e;
    };
    set((const void *) k_band_tail<D, RIG, false>, bt_smem_bytes(D, RIG));
    set((const void *) k_band_tail<D, RIG, true>, bt_smem_bytes(D, RIG));
    set((const void *) k_band_dp<D, RIG, false>, bd_smem_bytes());
    set((const void *) k_band_dp<D, RIG, true>, bd_smem_bytes());
    set((const void *) k_mmap_full_strips<D, RIG, false>, mf_smem_bytes(D, RIG));
    set((const void *) k_mmap_full_strips<D, RIG, true>, mf_smem_bytes(D, RIG));
}

int raise_smem_limits(const B200Carver *c)
{
    static std::mutex mu;
    static std::map<int, cudaError_t> done; // (device, delta_x, rigidity) -> outcome of the opt-in
    const int device = c->device, D = c->delta_x > 4 ? 4 : c->delta_x;
    const bool rig = c->rigidity != 0.f;
    const int key = device * 64 + D * 2 + (rig ? 1 : 0), key0 = device * 64 + 63;
    std::lock_guard<std::mutex> lk(mu);
    cudaError_t err = cudaSuccess;
    if (done.find(key0) == done.end()) { // the kernels without template parameters, once per device
        err = cudaSetDevice(device);
        if (err == cudaSuccess) err = cudaFuncSetAttribute((const void *) k_seam_path, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sp_smem_bytes());
        if (err == cudaSuccess) err = cudaFuncSetAttribute((const void *) k_seam_chase, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) st_chase_smem());
        if (err == cudaSuccess) err = cudaFuncSetAttribute((const void *) k_carve_row, cudaFuncAttributeMaxDynamicSharedMemorySize, B200C_CARVE_ROW_SMEM_MAX);
        done.emplace(key0, err);
    }
    if (done[key0] != cudaSuccess) return fail(B200C_ERROR, "cudaFuncSetAttribute(max dynamic shared memory)", done[key0]);
    auto it = done.find(key);
    if (it == done.end()) {
        err = cudaSetDevice(device);
        if (err == cudaSuccess) {
            switch (D * 2 + (rig ? 1 : 0)) {
                case 0: raise_smem_limits_dr<0, false>(err); break;
                case 1: raise_smem_limits_dr<0, true>(err); break;
                case 2: raise_smem_limits_dr<1, false>(err); break;
                case 3: raise_smem_limits_dr<1, true>(err); break;
                case 4: raise_smem_limits_dr<2, false>(err); break;
                case 5: raise_smem_limits_dr<2, true>(err); break;
                case 6: raise_smem_limits_dr<3, false>(err); break;
                case 7: raise_smem_limits_dr<3, true>(err); break;
                case 8: raise_smem_limits_dr<4, false>(err); break;
                default: raise_smem_limits_dr<4, true>(err); break;
            }
        }
        it = done.emplace(key, err).first;
    }
    if (it->second != cudaSuccess) return fail(B200C_ERROR, "cudaFuncSetAttribute(max dynamic shared memory)", it->second);
    return B200C_OK;
}

// fix == false: the band DP itself (k_band_dp); fix == true: the parents of the cells it evaluated (k_fix_parents)
template <int D>
const void *band_dp_fn_d(bool fix, bool rig, bool lr)
{
    if (fix) {
        if (rig && lr) return (const void *) k_fix_parents<D, true, true>;
        if (rig) return (const void *) k_fix_parents<D, true, false>;
        if (lr) return (const void *) k_fix_parents<D, false, true>;
        return (const void *) k_fix_parents<D, false, false>;
    }
    if (rig && lr) return (const void *) k_band_dp<D, true, true>;
    if (rig) return (const void *) k_band_dp<D, true, false>;
    if (lr) return (const void *) k_band_dp<D, false, true>;
    return (const void *) k_band_dp<D, false, false>;
}

template <int D>
const void *band_tail_fn_d(bool rig, bool lr)
{
    if (rig && lr) return (const void *) k_band_tail<D, true, true>;
    if (rig) return (const void *) k_band_tail<D, true, false>;
    if (lr) return (const void *) k_band_tail<D, false, true>;
    return (const void *) k_band_tail<D, false, false>;
}

const void *band_tail_fn(const B200Carver *c)
{
    const bool rig = c->rigidity != 0.f, lr = c->leftright != 0;
    switch (c->delta_x) {
        case 0: return band_tail_fn_d<0>(rig, lr);
        case 1: return band_tail_fn_d<1>(rig, lr);
        case 2: return band_tail_fn_d<2>(rig, lr);
        case 3: return band_tail_fn_d<3>(rig, lr);
        default: return band_tail_fn_d<4>(rig, lr);
    }
}

const void *band_dp_fn(const B200Carver *c, bool fix)
{
    const bool rig = c->rigidity != 0.f, lr = c->leftright != 0;
    switch (c->delta_x) {
        case 0: return band_dp_fn_d<0>(fix, rig, lr);
        case 1: return band_dp_fn_d<1>(fix, rig, lr);
        case 2: return band_dp_fn_d<2>(fix, rig, lr);
        case 3: return band_dp_fn_d<3>(fix, rig, lr);
        default: return band_dp_fn_d<4>(fix, rig, lr);
    }
}

template <int D>
void launch_mmap_full_d(B200Carver *c, int gridx, int y0, int rows)
{
    const DevP p = view(c);
    const DevP *tab = tab_static(c);
    const dim3 grid(gridx, 1, batch_n(c));
    const bool rig = c->rigidity != 0.f, lr = c->leftright != 0;
    const size_t sm = mf_smem_bytes(D, rig);
    if (rig && lr) k_mmap_full_strips<D, true, true><<<grid, MF_THREADS, sm, c->stream>>>(p, y0, rows, tab);
    else if (rig) k_mmap_full_strips<D, true, false><<<grid, MF_THREADS, sm, c->stream>>>(p, y0, rows, tab);
    else if (lr) k_mmap_full_strips<D, false, true><<<grid, MF_THREADS, sm, c->stream>>>(p, y0, rows, tab);
    else k_mmap_full_strips<D, false, false><<<grid, MF_THREADS, sm, c->stream>>>(p, y0, rows, tab);
}

// the full DP as one cluster launch per pass (mmap_cluster.cuh) + the parents of the finished map
template <int D, bool RIG>
int launch_mmap_cluster_dr(B200Carver *c, int csize, int nwarps)
{
    const void *fn = (const void *) k_mmap_full_cluster<D, RIG>;
    const size_t smem = (size_t) nwarps * mc_warp_bytes(RIG);
    {
        static std::mutex mu;
        static std::map<int, bool> done; // per device
        std::lock_guard<std::mutex> lk(mu);
        if (!done[c->device]) {
            CU_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            CU_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * mc_warp_bytes(RIG)));
            done[c->device] = true;
        }
    }
    DevP p = view(c);
    const DevP *tab = tab_static(c);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize, 1, batch_n(c));
    cfg.blockDim = dim3(nwarps * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    void *args[3] = {&p, &nwarps, &tab};
    CU_TRY(cudaLaunchKernelExC(&cfg, fn, args));
    const dim3 pg((c->w + 1023) / 1024, c->h > 1 ? c->h - 1 : 1, batch_n(c));
    if (c->leftright)
        k_parents_full<D, RIG, true><<<pg, 256, 0, c->stream>>>(p, tab);
    else
        k_parents_full<D, RIG, false><<<pg, 256, 0, c->stream>>>(p, tab);
    return check_launch("k_parents_full");
}

int launch_mmap_cluster(B200Carver *c, int csize, int nwarps)
{
    const bool rig = c->rigidity != 0.f;
    switch (c->delta_x * 2 + (rig ? 1 : 0)) {
        case 0: return launch_mmap_cluster_dr<0, false>(c, csize, nwarps);
        case 1: return launch_mmap_cluster_dr<0, true>(c, csize, nwarps);
        case 2: return launch_mmap_cluster_dr<1, false>(c, csize, nwarps);
        case 3: return launch_mmap_cluster_dr<1, true>(c, csize, nwarps);
        case 4: return launch_mmap_cluster_dr<2, false>(c, csize, nwarps);
        case 5: return launch_mmap_cluster_dr<2, true>(c, csize, nwarps);
        case 6: return launch_mmap_cluster_dr<3, false>(c, csize, nwarps);
        case 7: return launch_mmap_cluster_dr<3, true>(c, csize, nwarps);
        case 8: return launch_mmap_cluster_dr<4, false>(c, csize, nwarps);
        default: return launch_mmap_cluster_dr<4, true>(c, csize, nwarps);
    }
}

int build_mmap(B200Carver *c)
{
    B_TRY(upload_tab(c, false));
    if (fast_path(c) && c->delta_x <= 4 && c->use_cluster) {
        const int wlim = std::min((c->w + 4 + 3) & ~3, c->pitch);
        const int nseg = (wlim + MC_S - 1) / MC_S;
        int nwarps = nseg <= 64 ? 4 : (nseg + 15) / 16, csize = 1;
        while (csize * nwarps < nseg) csize *= 2;
        if (csize <= 16 && nwarps <= 8) {
            StageScope sc("mmap_full", c->stream, 2);
            return launch_mmap_cluster(c, csize, nwarps);
        }
    }
    if (fast_path(c) && c->delta_x <= 4) {
        B_TRY(raise_smem_limits(c));
        const int R = mf_rows(c->delta_x), S = 128 - 2 * mf_hk(c->delta_x);
        const int nstrips = (c->w + 4 + S - 1) / S;
        const int grid = (nstrips + MF_WARPS - 1) / MF_WARPS;
        StageScope sc("mmap_full", c->stream, (c->h + R - 1) / R);
        for (int y0 = 0; y0 < c->h; y0 += R) {
            const int rows = c->h - y0 < R ? c->h - y0 : R;
            switch (c->delta_x) {
                case 0: launch_mmap_full_d<0>(c, grid, y0, rows); break;
                case 1: launch_mmap_full_d<1>(c, grid, y0, rows); break;
                case 2: launch_mmap_full_d<2>(c, grid, y0, rows); break;
                case 3: launch_mmap_full_d<3>(c, grid, y0, rows); break;
                default: launch_mmap_full_d<4>(c, grid, y0, rows); break;
            }
        }
        return check_launch("k_mmap_full_strips");
    }
    StageScope sc("mmap_full", c->stream);
    k_mmap_full<<<dim3(1, 1, batch_n(c)), 1024, 0, c->stream>>>(view(c), tab_static(c));
    return check_launch("k_mmap_full");
}

// ---- A.9 inflate ----------------------------------------------------------------------------------------
int inflate(B200Carver *c, int l)
{
    for (B200Carver *a : c->attached) B_TRY(inflate(a, l));

    set_width_one(c, c->w0);
    const int w1 = c->w0 + l - c->max_level + 1;
    const size_t n1 = (size_t) w1 * c->h0;
    uint8_t *new_rgb = nullptr;
    int *new_vs = nullptr;
    float *new_bias = nullptr, *new_rigmask = nullptr;
    B_TRY(dalloc(c, &new_rgb, n1 * c->channels, true));
    if (!c->root) B_TRY(dalloc(c, &new_vs, n1, true));
    if (c->active) {
        if (c->bias) B_TRY(dalloc(c, &new_bias, n1, true));
        if (c->rigmask) B_TRY(dalloc(c, &new_rigmask, n1, true));
    }
    {
        StageScope sc("inflate", c->stream);
        k_inflate_rows<<<c->h0, B200C_ROW_THREADS, 0, c->stream>>>(
            c->rgb, c->vs, c->w0, w1, c->channels, l, c->max_level, new_rgb, new_vs, c->bias, new_bias, c->rigmask,
            new_rigmask, c->raw, c->w_start);
        B_TRY(check_launch("k_inflate_rows"));
    }
    dfree(c, c->rgb);
    dfree(c, c->bias);
    dfree(c, c->rigmask);
    c->nrg_uptodate = false; // the compact maps keep their size (w_start x h); the next build_maps refills them
    c->raw_ident = false;
    c->rgb = new_rgb;
    if (!c->root) {
        dfree(c, c->vs);
        c->vs = new_vs;
        propagate_vs(c);
    }
    if (c->active) {
        c->bias = new_bias;
        c->rigmask = new_rigmask;
    }
    c->w0 = w1;
    c->w = c->w_start;
    c->level = l + 1;
    c->max_level = l + 1;
    return B200C_OK;
}

// ---- A.7 per-seam loop ------------------------------------------------------------------------------------
// One iteration of the per-seam loop (A.7) is a fixed list of kernel launches: backtrack, carve, band energy [, band DP
// + parent fix-up].  Every kernel takes the session's argument block (view_dyn) and reads the seam number from device
// memory, so the list is the same for every seam of a session: it is launched kernel by kernel, or -- as nodes of a
// CUDA graph built from the same list -- replayed with one call per seam.
struct SeamLaunch {
    const char *stage; // StageScope / error name
    const void *fn;
    dim3 grid, block;
    size_t smem;
    int second; // second kernel argument after the DevP block: 0 none, 1 the session's visibility epoch + carve phase, 2 the tensor maps
    bool coop;  // cooperative launch (grid barrier inside)
    int phase = 0; // k_carve: 0 whole row, 1 NEAR, 2 FAR
    int dep = -1;  // graph: index of the launch this one waits for (-1: the one before it)
};
constexpr int kSeamLaunchMax = 10;

// the whole-row carve: the one-pass kernel that stages the row in shared memory, unless the row does not fit (images
// wider than ~15000 columns) or the plain kernels were asked for (B200C_GENERIC=1)
const void *carve_row_fn(const B200Carver *c, size_t *smem)
{
    const size_t need = carve_row_smem(c->pitch, c->rig != nullptr);
    if (fast_path(c) && need <= B200C_CARVE_ROW_SMEM_MAX) {
        *smem = need;
        return (const void *) k_carve_row;
    }
    *smem = 0;
    return (const void *) k_carve;
}

// the backtrack: jump tables over all SMs + a short chase (seam_trace.cuh); the single-CTA staged chase for delta_x > 4;
// the plain walk with B200C_GENERIC=1
int vpath_launch_list(const B200Carver *c, SeamLaunch out[kSeamLaunchMax])
{
    int n = 0;
    if (fast_path(c) && c->delta_x <= 4 && c->h <= ST_HMAX && c->w_epoch <= ST_WMAX && c->use_trace) {
        const int nblk = st_nblk(c->h, c->delta_x);
        if (nblk > 0)
            out[n++] = {"seam_jumps", (const void *) k_seam_jumps, dim3((c->w_epoch + ST_COLS - 1) / ST_COLS, nblk), dim3(ST_THREADS),
                        st_jump_smem(c->delta_x), 0, false};
        out[n++] = {"vpath", (const void *) k_seam_chase, dim3(1), dim3(ST_CHASE_THREADS), st_chase_smem(), 0, false};
    } else if (fast_path(c)) {
        out[n++] = {"vpath", (const void *) k_seam_path, dim3(1), dim3(SP_THREADS), sp_smem_bytes(), 0, false};
    } else {
        out[n++] = {"vpath", (const void *) k_vpath, dim3(1), dim3(1024), 0, 0, false};
    }
    return n;
}

int seam_launch_list(const B200Carver *c, bool with_update, SeamLaunch out[kSeamLaunchMax])
{
    const bool fast = fast_path(c), band = fast && c->delta_x <= 4 && c->h <= BD_HMAX;
    int n = vpath_launch_list(c, out);
    int near_at = -1;
    if (split_carve(c) && with_update) {
        near_at = n;
        out[n++] = {"carve", (const void *) k_carve, dim3(c->h), dim3(B200C_CARVE_THREADS), 0, 1, false, 1};
        out[n++] = {"carve_far", (const void *) k_carve, dim3(c->h), dim3(B200C_CARVE_THREADS), 0, 1, false, 2};
    } else {
        size_t smem = 0;
        const void *fn = carve_row_fn(c, &smem);
        out[n++] = {"carve", fn, dim3(c->h), dim3(B200C_CARVE_THREADS), smem, 1, false};
    }
    out[n++] = {"energy_band", (const void *) k_energy_band, dim3((c->h + B200C_EB_ROWS - 1) / B200C_EB_ROWS), dim3(256), 0, 0, false};
    if (near_at >= 0) out[n - 1].dep = near_at; // beside the FAR phase
    if (!with_update) return n;
    if (band) {
        out[n++] = {"mmap_update", band_dp_fn(c, false), dim3(1), dim3(BD_THREADS), bd_smem_bytes(), 2, false};
        if (c->use_tail) // the rows the band kernel could not tile (none, most of the time: the kernel returns at once)
            out[n++] = {"mmap_tail", band_tail_fn(c), dim3(bt_grid(c->w_epoch, c->delta_x)), dim3(BT_THREADS),
                        bt_smem_bytes(c->delta_x > 4 ? 4 : c->delta_x, c->rigidity != 0.f), 0, true};
        out[n++] = {"fix_parents", band_dp_fn(c, true), dim3((c->h + 7) / 8, c->mates.empty() ? 4 : 1), dim3(256), 0, 0, false};
    } else {
        out[n++] = {"mmap_update", (const void *) k_mmap_update, dim3(1), dim3(512), 0, 0, false};
    }
    return n;
}

// kernel parameters of one launch of the list: (DevP, table) | (DevP, int, table) | (DevP, BdMaps, table, map table)
struct SeamArgs {
    DevP p;
    int epoch, phase;
    const DevP *tab;
    const BdMaps *mtab;
    void *ptr[4];
    void **of(const B200Carver *c, const SeamLaunch &l)
    {
        int n = 0;
        ptr[n++] = &p;
        if (l.second == 1) {
            ptr[n++] = &epoch;
            phase = l.phase;
            ptr[n++] = &phase;
        }
        if (l.second == 2) ptr[n++] = const_cast<BdMaps *>(&c->maps);
        ptr[n++] = &tab;
        if (l.second == 2) ptr[n++] = &mtab;
        return ptr;
    }
};
SeamArgs seam_args(const B200Carver *c)
{
    SeamArgs a;
    a.p = view_dyn(c);
    a.epoch = c->vs_epoch;
    a.tab = tab_dyn(c);
    a.mtab = tab_maps(c);
    return a;
}
dim3 batch_grid(const B200Carver *c, dim3 g) { return dim3(g.x, g.y, batch_n(c)); }

int launch_seam_kernels(B200Carver *c, bool with_update)
{
    SeamLaunch L[kSeamLaunchMax];
    const int n = seam_launch_list(c, with_update, L);
    SeamArgs a = seam_args(c);
    for (int i = 0; i < n; ++i) {
        StageScope sc(L[i].stage, c->stream);
        void **args = a.of(c, L[i]);
        const dim3 grid = batch_grid(c, L[i].grid);
        const cudaError_t e = L[i].coop ? cudaLaunchCooperativeKernel(L[i].fn, grid, L[i].block, args, L[i].smem, c->stream)
                                        : cudaLaunchKernel(L[i].fn, grid, L[i].block, args, L[i].smem, c->stream);
        if (e != cudaSuccess) return fail(B200C_ERROR, L[i].stage, e);
    }
    return B200C_OK;
}

void drop_seam_graphs(B200Carver *c) { c->graph_fresh[0] = c->graph_fresh[1] = false; }

// which kernels one iteration consists of: the lane keeps one executable per set
int graph_key(const B200Carver *c)
{
    const bool fast = fast_path(c), band = fast && c->delta_x <= 4 && c->h <= BD_HMAX;
    return (fast ? 1 : 0) | (band ? 2 : 0) | ((c->leftright & 1) << 2) | (c->rigidity != 0.f ? 8 : 0) | (c->delta_x << 4) |
           (c->use_tail ? 1 << 12 : 0) | (c->use_trace ? 1 << 13 : 0) | (c->mates.empty() ? 0 : 1 << 14) | (split_carve(c) ? 1 << 15 : 0) | (c->use_pdl ? 1 << 16 : 0);
}

// Points the lane's graph for the current kernel set at this carver's session: the first time the nodes are added and
// the graph instantiated, later only the node parameters of the executable change.  (Stream capture is not used: with
// other host threads waiting on their own streams, cudaStreamBeginCapture was measured to block for tens of ms.)
// pdl: the edge between two ordinary kernel nodes is a PROGRAMMATIC one -- the dependent grid is launched as soon as every
// CTA of its predecessor has started (each per-seam kernel begins with griddepcontrol.launch_dependents) and its CTAs wait
// in griddepcontrol.wait until the predecessor has completed and flushed: the launch latency of the dependent kernel
// disappears from the chain.  Edges that touch the cooperative tail kernel stay ordinary.
static int seam_graph_build(B200Carver *c, LaneGraph &g, bool pdl)
{
    SeamLaunch L[kSeamLaunchMax];
    const int n = seam_launch_list(c, true, L);
    SeamArgs a = seam_args(c);
    if (g.exec && g.n != n) lane_graph_reset(&g);
    const bool fresh = g.exec == nullptr;
    if (fresh) CU_TRY(cudaGraphCreate(&g.graph, 0));
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < n && e == cudaSuccess; ++i) {
        cudaKernelNodeParams kp = {};
        kp.func = const_cast<void *>(L[i].fn);
        kp.gridDim = batch_grid(c, L[i].grid);
        kp.blockDim = L[i].block;
        kp.sharedMemBytes = (unsigned) L[i].smem;
        kp.kernelParams = a.of(c, L[i]);
        if (fresh) {
            const int dep = i ? (L[i].dep >= 0 ? L[i].dep : i - 1) : -1;
            const bool prog = pdl && dep >= 0 && !L[i].coop && !L[dep].coop;
            e = cudaGraphAddKernelNode(&g.node[i], g.graph, (dep >= 0 && !prog) ? &g.node[dep] : nullptr, (dep >= 0 && !prog) ? 1 : 0, &kp);
            if (e == cudaSuccess && L[i].coop) {
                cudaKernelNodeAttrValue v = {};
                v.cooperative = 1;
                e = cudaGraphKernelNodeSetAttribute(g.node[i], cudaKernelNodeAttributeCooperative, &v);
            }
            if (e == cudaSuccess && prog) {
                cudaGraphEdgeData ed = {};
                ed.from_port = cudaGraphKernelNodePortProgrammatic;
                ed.to_port = 0;
                ed.type = cudaGraphDependencyTypeProgrammatic;
                e = cudaGraphAddDependencies_v2(g.graph, &g.node[dep], &g.node[i], &ed, 1);
            }
        } else
            e = cudaGraphExecKernelNodeSetParams(g.exec, g.node[i], &kp);
    }
    if (e == cudaSuccess && fresh) e = cudaGraphInstantiate(&g.exec, g.graph, 0);
    if (e != cudaSuccess) {
        lane_graph_reset(&g);
        return fail(B200C_ERROR, "seam graph", e);
    }
    g.n = n;
    g_launches += n;
    return B200C_OK;
}

int seam_graph_prepare(B200Carver *c, LaneGraph **out)
{
    LaneGraph &g = c->lane->graphs[graph_key(c)];
    const bool pdl = c->use_pdl && c->mates.empty();
    int rc = seam_graph_build(c, g, pdl);
    if (rc != B200C_OK && pdl) { // no programmatic edges on this driver / for this kernel set: ordinary ones
        cudaGetLastError();
        c->use_pdl = false;
        rc = seam_graph_build(c, c->lane->graphs[graph_key(c)], false);
        if (rc == B200C_OK) {
            *out = &c->lane->graphs[graph_key(c)];
            return rc;
        }
    }
    if (rc == B200C_OK) *out = &g;
    return rc;
}

int seam_iteration(B200Carver *c, int l, int lr_switch_interval)
{
    cudaStream_t s = c->stream;
    if (fast_path(c)) B_TRY(raise_smem_limits(c));
    const bool last = c->w - 1 <= 1; // the image is about to be one pixel wide
    const bool lr_switch = !last && c->lr_freq && ((l - c->max_level + lr_switch_interval / 2) % lr_switch_interval) == 0;
    if (last) {
        SeamLaunch L[kSeamLaunchMax];
        const int nv = vpath_launch_list(c, L);
        SeamArgs a = seam_args(c);
        for (int i = 0; i < nv; ++i) {
            StageScope sc(L[i].stage, s);
            const cudaError_t e = cudaLaunchKernel(L[i].fn, batch_grid(c, L[i].grid), L[i].block, a.of(c, L[i]), L[i].smem, s);
            if (e != cudaSuccess) return fail(B200C_ERROR, L[i].stage, e);
        }
        StageScope sc2("carve", s);
        size_t smem = 0;
        const void *fn = carve_row_fn(c, &smem);
        DevP v = view_dyn(c);
        int vs0 = c->vs_epoch, ph = 0;
        const DevP *tb = tab_dyn(c);
        void *args[] = {&v, &vs0, &ph, &tb};
        const cudaError_t e = cudaLaunchKernel(fn, dim3(c->h, 1, batch_n(c)), dim3(B200C_CARVE_THREADS), args, smem, s);
        if (e != cudaSuccess) return fail(B200C_ERROR, "k_carve", e);
    } else if (lr_switch || !c->use_graph || g_timing) {
        B_TRY(launch_seam_kernels(c, !lr_switch));
    } else {
        const int lr = c->leftright & 1;
        if (!c->graph_fresh[lr]) {
            HostScope hs(7);
            B_TRY(seam_graph_prepare(c, &c->graph[lr]));
            c->graph_fresh[lr] = true;
        } else {
            g_launches += c->graph[lr]->n;
        }
        cudaGraphExec_t exec = c->graph[lr]->exec;
        HostScope hs(8);
        CU_TRY(cudaGraphLaunch(exec, s));
    }
    for_batch(c, [&](B200Carver *m) {
        m->level++;
        m->w--;
        m->nrg_uptodate = !last;
        m->raw_ident = false;
    });
    if (last) {
        StageScope sc("finish_vsmap", s);
        k_finish_vsmap<<<dim3((c->h + 255) / 256, 1, batch_n(c)), 256, 0, s>>>(view_dyn(c), tab_dyn(c));
        B_TRY(check_launch("k_finish_vsmap"));
    } else if (lr_switch) {
        for_batch(c, [](B200Carver *m) { m->leftright ^= 1; });
        B_TRY(upload_tab(c, true)); // the argument blocks carry the tie rule
        B_TRY(build_mmap(c));
    }
    return B200C_OK;
}

// the rigidity mask in current coordinates, for the DP kernels (start of a build_maps session)
int gather_rig(B200Carver *c)
{
    dfree(c, c->rig);
    if (c->rigidity == 0.f) return B200C_OK; // without a mask the factor is 1 everywhere
    B_TRY(dalloc(c, &c->rig, (size_t) c->pitch * c->h_start + 64, true));
    B_TRY(encode_map(&c->maps.rig, c->rig, false, c->pitch, c->h_start, bd_rows(c->delta_x, true)));
    dim3 grid((c->w + 255) / 256, c->h);
    StageScope sc("gather_rig", c->stream);
    k_gather_rig<<<grid, 256, 0, c->stream>>>(view(c), nullptr);
    return check_launch("k_gather_rig");
}

int build_vsmap(B200Carver *c, int depth, int update_step, b200c_progress_fn progress, void *user, bool do_inflate,
                int update_phase = 0)
{
    int lr_switch_interval = 0;
    if (depth == 0) depth = c->w_start + 1;
    if (c->lr_freq) lr_switch_interval = (depth - c->max_level - 1) / (int) c->lr_freq + 1;
    if (update_step < 1) update_step = 1;
    const int first = c->max_level;
    // session: the per-seam kernels get one argument block and count the seams themselves (DevP::dyn starts at -1)
    drop_seam_graphs(c);
    for_batch(c, [](B200Carver *m) { m->w_epoch = m->w; });
    if (!c->mates.empty()) for_batch(c, [](B200Carver *m) { m->use_tail = false; }); // one grid barrier cannot serve many images
    if (c->use_tail && fast_path(c) && c->delta_x <= 4 && c->h <= BD_HMAX) {
        // the tail kernel needs a grid barrier: every CTA of its grid must be resident at once (cooperative launch)
        B_TRY(raise_smem_limits(c));
        int coop = 0, sms = 0, per_sm = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, band_tail_fn(c), BT_THREADS,
                                                          bt_smem_bytes(c->delta_x, c->rigidity != 0.f)) != cudaSuccess)
            per_sm = 0;
        if (!coop || bt_grid(c->w_epoch, c->delta_x) > per_sm * sms) {
            cudaGetLastError();
            c->use_tail = false; // the band kernel keeps its own wide-window loop
        }
    }
    {
        cudaError_t e = cudaSuccess;
        for_batch(c, [&](B200Carver *m) {
            m->vs_epoch = first + m->max_level - 1;
            cudaError_t e1 = cudaMemsetAsync(m->dyn_d, 0xff, sizeof(int), c->stream);
            if (e1 == cudaSuccess && m->far_d) e1 = cudaMemsetAsync(m->far_d, 0, sizeof(int), c->stream);
            if (e == cudaSuccess) e = e1;
        });
        CU_TRY(e);
    }
    B_TRY(upload_tab(c, true));
    // Progress (LqrProgress update, render.c:767-779): liblqr reports "i seams done" when they ARE done.  The seam loop
    // is queued asynchronously, so every progress point is marked by an event in the queue and its callback is
    // delivered -- on this, the caller's thread -- once that event has completed, one update step behind the
    // enqueue front (the device never runs dry while the host is in the callback).  A callback that asks to cancel
    // stops the enqueueing; the seams already queued (at most one step) still complete.
    struct Point { int seam, ev; };
    Point pending[4];
    int n_pending = 0, n_points = 0;
    bool cancelled = false;
    if (c->lane && c->lane->seams_h) *c->lane->seams_h = 0;
    auto deliver = [&]() -> int { // the oldest pending point
        const Point pt = pending[0];
        for (int i = 1; i < n_pending; ++i) pending[i - 1] = pending[i];
        --n_pending;
        if (pt.ev >= 0) CU_TRY(cudaEventSynchronize(c->lane->prog[pt.ev]));
        if (progress(user, pt.seam)) cancelled = true;
        return B200C_OK;
    };
    for (int l = first; l < depth && !cancelled; ++l) {
        const int i = l - first;
        if (progress && ((i + update_phase) % update_step) == 0) {
            int ev = -1;
            if (i > 0 && c->lane) { // completes when the seams before i are done
                ev = n_points++ & 3;
                CU_TRY(cudaEventRecord(c->lane->prog[ev], c->stream));
            }
            if (n_pending == 4) B_TRY(deliver());
            pending[n_pending++] = {i, ev};
        }
        B_TRY(seam_iteration(c, l, lr_switch_interval));
        while (n_pending && !cancelled && pending[0].seam + update_step <= i + 1) B_TRY(deliver());
    }
    while (n_pending && !cancelled) B_TRY(deliver());
    if (cancelled) {
        CU_TRY(carver_sync(c)); // the queued seams complete; the carver is left mid-session, as liblqr leaves it
        return fail(B200C_CANCEL, "cancelled by progress callback");
    }
    {
        // the staged kernels check their own window invariants on the device; a violation is a hard error
        HostScope hs(9);
        int err = 0;
        std::vector<int> errs(batch_n(c), 0);
        {
            int i = 0;
            cudaError_t e = cudaSuccess;
            for_batch(c, [&](B200Carver *m) {
                const cudaError_t e1 = cudaMemcpyAsync(&errs[i++], m->err_d, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
                if (e == cudaSuccess) e = e1;
            });
            CU_TRY(e);
        }
        CU_TRY(carver_sync(c));
        for (int v : errs) err |= v;
        if (err) {
            char msg[96];
            snprintf(msg, sizeof msg, "seam loop: device invariant violated (code %d)", err);
            return fail(B200C_ERROR, msg);
        }
    }
    if (!do_inflate) return B200C_OK;
    int rc = B200C_OK;
    for_batch(c, [&](B200Carver *m) {
        if (rc != B200C_OK) return;
        rc = inflate(m, depth - 1);
        set_width_one(m, m->w_start);
        for (B200Carver *a : m->attached) set_width_rec(a, m->w_start);
    });
    return rc;
}

// ---- A.11 flatten / transpose ---------------------------------------------------------------------------------
int flatten(B200Carver *c)
{
    for (B200Carver *a : c->attached) B_TRY(flatten(a));

    free_maps(c);

    const size_t n = (size_t) c->w * c->h;
    uint8_t *new_rgb = nullptr;
    float *new_bias = nullptr, *new_rigmask = nullptr;
    B_TRY(dalloc(c, &new_rgb, n * c->channels, true));
    if (c->active && c->rigmask) B_TRY(dalloc(c, &new_rigmask, n, true));
    if (c->nrg_active && c->bias) B_TRY(dalloc(c, &new_bias, n, true));
    {
        StageScope sc("flatten", c->stream);
        k_compact_rows<<<c->h, B200C_ROW_THREADS, 0, c->stream>>>(c->rgb, c->vs, c->w0, c->w, c->channels, c->level,
                                                                  new_rgb, c->bias, new_bias, c->rigmask, new_rigmask);
        B_TRY(check_launch("k_compact_rows(flatten)"));
    }
    dfree(c, c->rgb);
    c->rgb = new_rgb;
    if (c->nrg_active) {
        dfree(c, c->bias);
        c->bias = new_bias;
    }
    if (c->active) {
        dfree(c, c->rigmask);
        c->rigmask = new_rigmask;
    }
    if (!c->root) {
        dfree(c, c->vs);
        B_TRY(dalloc(c, &c->vs, n, true));
        propagate_vs(c);
    }
    c->w0 = c->w;
    c->h0 = c->h;
    c->w_start = c->w;
    c->h_start = c->h;
    c->level = 1;
    c->max_level = 1;
    if (c->nrg_active) {
        dfree(c, c->raw);
        B_TRY(dalloc(c, &c->raw, n + 16, false)); // slack: see alloc of raw in carver_init
        B_TRY(init_raw(c));
    }
    B_TRY(alloc_maps(c));
    return B200C_OK;
}

int transpose_buffer(B200Carver *c, const void *in, void *out, int elem, int w, int h)
{
    dim3 grid((w + 31) / 32, (h + 31) / 32);
    StageScope sc("transpose", c->stream);
    const uint8_t *i8 = (const uint8_t *) in;
    uint8_t *o8 = (uint8_t *) out;
    switch (elem) {
        case 1: k_transpose<1><<<grid, 256, 0, c->stream>>>(i8, o8, w, h); break;
        case 2: k_transpose<2><<<grid, 256, 0, c->stream>>>(i8, o8, w, h); break;
        case 3: k_transpose<3><<<grid, 256, 0, c->stream>>>(i8, o8, w, h); break;
        case 4: k_transpose<4><<<grid, 256, 0, c->stream>>>(i8, o8, w, h); break;
        default: return fail(B200C_ERROR, "transpose: unsupported element size");
    }
    return check_launch("k_transpose");
}

int transpose(B200Carver *c)
{
    if (c->level > 1) B_TRY(flatten(c));
    for (B200Carver *a : c->attached) B_TRY(transpose(a));

    const size_t n = (size_t) c->w0 * c->h0;
    free_maps(c);

    uint8_t *new_rgb = nullptr;
    float *new_bias = nullptr, *new_rigmask = nullptr;
    B_TRY(dalloc(c, &new_rgb, n * c->channels, true));
    B_TRY(transpose_buffer(c, c->rgb, new_rgb, c->channels, c->w0, c->h0));
    if (c->nrg_active && c->bias) {
        B_TRY(dalloc(c, &new_bias, n, true));
        B_TRY(transpose_buffer(c, c->bias, new_bias, 4, c->w0, c->h0));
    }
    if (c->active && c->rigmask) {
        B_TRY(dalloc(c, &new_rigmask, n, true));
        B_TRY(transpose_buffer(c, c->rigmask, new_rigmask, 4, c->w0, c->h0));
    }
    dfree(c, c->rgb);
    c->rgb = new_rgb;
    if (c->nrg_active) {
        dfree(c, c->bias);
        c->bias = new_bias;
    }
    if (c->active) {
        dfree(c, c->rigmask);
        c->rigmask = new_rigmask;
    }
    if (!c->root) {
        dfree(c, c->vs);
        B_TRY(dalloc(c, &c->vs, n, true));
        propagate_vs(c);
    }

    const int d = c->w0;
    c->w0 = c->h0;
    c->h0 = d;
    c->w = c->w0;
    c->h = c->h0;
    c->w_start = c->w0;
    c->h_start = c->h0;
    c->level = 1;
    c->max_level = 1;

    if (c->nrg_active) {
        dfree(c, c->raw);
        B_TRY(dalloc(c, &c->raw, n + 16, false)); // slack: see alloc of raw in carver_init
        B_TRY(init_raw(c));
    }
    B_TRY(alloc_maps(c));
    if (c->active) {
        dfree(c, c->vpath_x);
        dfree(c, c->nrg_xmin);
        dfree(c, c->nrg_xmax);
        dfree(c, c->nrg_pack);
        dfree(c, c->fix_d);
        B_TRY(dalloc(c, &c->vpath_x, (size_t) c->h, true));
        B_TRY(dalloc(c, &c->nrg_xmin, (size_t) c->h, true));
        B_TRY(dalloc(c, &c->nrg_xmax, (size_t) c->h, true));
        B_TRY(dalloc(c, &c->nrg_pack, (size_t) c->h, true));
        B_TRY(dalloc(c, &c->fix_d, (size_t) c->h / 8 + 4, true));
        for (int x = -c->delta_x; x <= c->delta_x; ++x) {
            float &v = c->rigmap_h[x + c->delta_x];
            v = v * c->w0 / c->h0;
        }
        B_TRY(upload_rigmap(c));
    }
    c->transposed = c->transposed ? 0 : 1;
    return B200C_OK;
}

bool not_at_reference(const B200Carver *c)
{
    return c->w != c->w0 || c->w_start != c->w0 || c->h != c->h0 || c->h_start != c->h0;
}

// ---- pinned host staging, pooled per process: page-locking tens of MB costs more than moving them, and the plug-in
// creates a fresh carver for every layer (render.c:222,894), so the buffers outlive the handles.
struct PinnedBuf {
    uint8_t *p;
    size_t cap;
};
std::mutex g_pin_mu;
std::vector<PinnedBuf> g_pin_free;
constexpr size_t kPinKeepBytes = 2ull << 30; // pinned bytes kept for reuse (cudaFreeHost synchronises the device:
                                             // with many carvers in flight a free per image serialises them all)
size_t g_pin_bytes = 0;

int pinned_acquire(size_t bytes, uint8_t **out, size_t *cap)
{
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        int best = -1;
        for (int i = 0; i < (int) g_pin_free.size(); ++i)
            if (g_pin_free[i].cap >= bytes && (best < 0 || g_pin_free[i].cap < g_pin_free[best].cap)) best = i;
        if (best >= 0) {
            *out = g_pin_free[best].p;
            *cap = g_pin_free[best].cap;
            g_pin_bytes -= *cap;
            g_pin_free.erase(g_pin_free.begin() + best);
            return B200C_OK;
        }
    }
    *out = nullptr;
    CU_TRY(cudaHostAlloc((void **) out, bytes, cudaHostAllocDefault));
    *cap = bytes;
    return B200C_OK;
}

void pinned_release(uint8_t *p, size_t cap)
{
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        if (g_pin_bytes + cap <= kPinKeepBytes) {
            g_pin_free.push_back({p, cap});
            g_pin_bytes += cap;
            return;
        }
    }
    cudaFreeHost(p);
}

int ensure_host_out(B200Carver *c, size_t bytes)
{
    if (bytes <= c->host_out_cap) return B200C_OK;
    if (c->host_out) CU_TRY(carver_sync(c)); // chunks of an abandoned read-out may still be on their way into it
    pinned_release(c->host_out, c->host_out_cap);
    c->host_out = nullptr;
    c->host_out_cap = 0;
    return pinned_acquire(bytes, &c->host_out, &c->host_out_cap);
}

// A few helper threads for the host side of large uploads: one core copies pageable memory at ~5 GB/s, well below
// what the DMA engine moves from pinned memory, so the staging copy of a chunk is split over the helpers.
class CopyHelpers {
  public:
    static CopyHelpers &get()
    {
        static CopyHelpers *h = new CopyHelpers; // never destroyed: its threads outlive static destruction
        return *h;
    }
    // memcpy(dst, src, n) split over the helpers and the calling thread; returns when done
    void copy(uint8_t *dst, const uint8_t *src, size_t n)
    {
        const int parts = (int) workers_.size() + 1;
        if (workers_.empty() || n < (1u << 20)) {
            memcpy(dst, src, n);
            return;
        }
        const size_t slice = (n / parts + 4095) & ~(size_t) 4095;
        {
            std::lock_guard<std::mutex> lk(mu_);
            dst_ = dst, src_ = src, n_ = n, slice_ = slice;
            pending_ = (int) workers_.size();
            ++epoch_;
        }
        cv_.notify_all();
        do_slice(0);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending_ == 0; });
    }

  private:
    CopyHelpers()
    {
        unsigned hc = std::thread::hardware_concurrency();
        int n = hc >= 8 ? 3 : (hc >= 4 ? 2 : (hc >= 2 ? 1 : 0));
        const char *e = getenv("B200C_COPY_THREADS");
        if (e) n = atoi(e) < 0 ? 0 : (atoi(e) > 7 ? 7 : atoi(e));
        for (int i = 0; i < n; ++i) workers_.emplace_back([this, i] { loop(i + 1); }), workers_.back().detach();
    }
    void do_slice(int part)
    {
        const size_t a = (size_t) part * slice_;
        if (a < n_) memcpy(dst_ + a, src_ + a, n_ - a < slice_ ? n_ - a : slice_);
    }
    void loop(int part)
    {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return epoch_ != seen; });
                seen = epoch_;
            }
            do_slice(part);
            std::lock_guard<std::mutex> lk(mu_);
            if (--pending_ == 0) done_.notify_one();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    uint8_t *dst_ = nullptr;
    const uint8_t *src_ = nullptr;
    size_t n_ = 0, slice_ = 0;
    int pending_ = 0;
    unsigned long long epoch_ = 0;
};
std::mutex g_upload_mu; // one large upload at a time uses the helpers

// pageable host memory -> device through two pinned chunks: the CPU copy of chunk i+1 overlaps the DMA of chunk i
int upload_pageable(B200Carver *c, void *dst, const void *src, size_t bytes)
{
    constexpr size_t kChunk = 8u << 20;
    if (bytes <= (1u << 20)) {
        CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(carver_sync(c)); // the caller may free `src` right after we return
        return B200C_OK;
    }
    uint8_t *stage[2] = {nullptr, nullptr};
    size_t cap[2] = {0, 0};
    cudaEvent_t *ev = (c->lane ? c->lane : c->root->lane)->ev; // attached carvers run on their root's lane
    int rc = B200C_OK;
    for (int i = 0; i < 2 && rc == B200C_OK; ++i) {
        HostScope hs(2);
        rc = pinned_acquire(kChunk, &stage[i], &cap[i]);
    }
    size_t off = 0;
    for (int i = 0; rc == B200C_OK && off < bytes; ++i, off += kChunk) {
        const int b = i & 1;
        const size_t n = bytes - off < kChunk ? bytes - off : kChunk;
        if (i >= 2) {
            HostScope hs(6);
            if (cudaEventSynchronize(ev[b]) != cudaSuccess) rc = fail(B200C_ERROR, "upload: event", cudaGetLastError());
        }
        {
            std::unique_lock<std::mutex> lk(g_upload_mu, std::defer_lock);
            const bool helpers = !host_shared(); // several carvers in flight: their own threads are the parallelism
            if (helpers) {
                HostScope hw(3);
                lk.lock();
            }
            HostScope hc(4);
            if (helpers)
                CopyHelpers::get().copy(stage[b], (const uint8_t *) src + off, n);
            else
                memcpy(stage[b], (const uint8_t *) src + off, n);
        }
        HostScope hs5(5);
        if (rc == B200C_OK && (cudaMemcpyAsync((uint8_t *) dst + off, stage[b], n, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
                               cudaEventRecord(ev[b], c->stream) != cudaSuccess))
            rc = fail(B200C_ERROR, "upload: cudaMemcpyAsync", cudaGetLastError());
    }
    {
        HostScope hs(6);
        if (carver_sync(c) != cudaSuccess && rc == B200C_OK) rc = fail(B200C_ERROR, "upload: sync", cudaGetLastError());
    }
    HostScope hs2(2);
    for (int i = 0; i < 2; ++i) pinned_release(stage[i], cap[i]);
    return rc;
}

B200Carver *carver_new_common(int width, int height, int channels)
{
    if (width < 1 || height < 1 || channels < 1 || channels > 4) {
        fail(B200C_ERROR, "carver_new: bad geometry (channels must be 1..4)");
        return nullptr;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1) {
        fail(B200C_ERROR, "no CUDA device: the B200 engine has no CPU fallback", e);
        return nullptr;
    }
    static std::once_flag once;
    std::call_once(once, [] {
        const char *t = getenv("B200C_TIMING");
        if (t) g_timing = atoi(t) != 0;
        const char *bs = getenv("B200C_BLOCKING_SYNC");
        if (bs && atoi(bs)) cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync);
    });
    B200Carver *c = new (std::nothrow) B200Carver();
    if (!c) {
        fail(B200C_NOMEM, "carver_new: host allocation");
        return nullptr;
    }
    {
        const char *g = getenv("B200C_GENERIC");
        c->generic = g && atoi(g) != 0;
        const char *gr = getenv("B200C_GRAPH");
        if (gr) c->use_graph = atoi(gr) != 0;
        const char *pd = getenv("B200C_PDL");
        if (pd) c->use_pdl = atoi(pd) != 0;
        const char *tl = getenv("B200C_TAIL");
        if (tl) c->use_tail = atoi(tl) != 0;
        const char *sp = getenv("B200C_SPLIT");
        if (sp) c->use_split = atoi(sp) != 0;
        const char *cl = getenv("B200C_CLUSTER");
        if (cl) c->use_cluster = atoi(cl) != 0;
        const char *tr = getenv("B200C_TRACE");
        if (tr) c->use_trace = atoi(tr) != 0;
        const char *ms = getenv("B200C_BD_MAXSEG");
        if (ms && atoi(ms) > 0) c->bd_maxseg = atoi(ms);

    }
    c->device = g_device >= 0 ? g_device : g_device_tls_default;
    if (cudaSetDevice(c->device) != cudaSuccess || !(c->lane = lane_acquire(c->device, g_use_ext_stream, g_ext_stream))) {
        fail(B200C_ERROR, "carver_new: cannot create stream", cudaGetLastError());
        delete c;
        return nullptr;
    }
    c->stream = c->lane->stream;
    c->pool = engine_pool(c->device);
    c->w = c->w0 = c->w_start = width;
    c->h = c->h0 = c->h_start = height;
    c->channels = channels;
    c->alpha = (channels == 2 || channels == 4) ? channels - 1 : -1;
    return c;
}

} // namespace

// A batch host keeps one stream per image in flight; the driver maps streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware
// queues (8 by default) and streams that share a queue serialise -- measured: 16 images in flight ran their per-seam
// chains two by two.  Ask for 32 queues unless the host process chose a value; this runs at load time, before the
// first CUDA call of this library creates the context (no effect if the process initialised CUDA earlier).
__attribute__((constructor)) static void b200c_on_load() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

// =================================================================================================== C ABI
extern "C" {

int b200c_abi_version(void) { return B200C_ABI_VERSION; }
double b200c_hostprof_ms(int idx) { return idx >= 0 && idx < 16 ? g_hostprof_ns[idx].load() * 1e-6 : 0.0; }
const char *b200c_last_error(void) { return g_err.c_str(); }

int b200c_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int b200c_device_cc(int device)
{
    int major = 0, minor = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess ||
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return major * 10 + minor;
}

int b200c_set_device(int device)
{
    int n = b200c_device_count();
    if (device < 0 || device >= n) return fail(B200C_ERROR, "b200c_set_device: no such device");
    g_device = device;
    return B200C_OK;
}

B200Carver *b200c_carver_new(const unsigned char *rgb, int width, int height, int channels)
{
    if (!rgb) {
        fail(B200C_ERROR, "carver_new: NULL image");
        return nullptr;
    }
    B200Carver *c;
    {
        HostScope hs(0);
        c = carver_new_common(width, height, channels);
    }
    if (!c) return nullptr;
    const size_t n = (size_t) width * height;
    bool ok;
    {
        HostScope hs(1);
        ok = dalloc(c, &c->rgb, n * channels, false) == B200C_OK && dalloc(c, &c->vs, n, true) == B200C_OK;
    }
    if (!ok ||
        upload_pageable(c, c->rgb, rgb, n * channels) != B200C_OK) { // synchronous: the caller may free `rgb` right after
        fail(B200C_NOMEM, "carver_new: device allocation / upload failed", cudaGetLastError());
        b200c_carver_destroy(c);
        return nullptr;
    }
    return c;
}

B200Carver *b200c_carver_new_device(const void *d_rgb, int width, int height, int channels)
{
    if (!d_rgb) {
        fail(B200C_ERROR, "carver_new_device: NULL image");
        return nullptr;
    }
    B200Carver *c = carver_new_common(width, height, channels);
    if (!c) return nullptr;
    const size_t n = (size_t) width * height;
    if (dalloc(c, &c->rgb, n * channels, false) != B200C_OK || dalloc(c, &c->vs, n, true) != B200C_OK ||
        cudaMemcpyAsync(c->rgb, d_rgb, n * channels, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess) {
        fail(B200C_NOMEM, "carver_new_device: device allocation / copy failed", cudaGetLastError());
        b200c_carver_destroy(c);
        return nullptr;
    }
    return c;
}

void b200c_carver_destroy(B200Carver *c)
{
    if (!c) return;
    HostScope hs(11);
    cudaSetDevice(c->device);
    for (B200Carver *a : c->attached) b200c_carver_destroy(a);
    if (c->stream) carver_sync(c);
    dfree(c, c->rgb);
    if (!c->root) dfree(c, c->vs);
    free_maps(c);
    dfree(c, c->raw);
    dfree(c, c->bias);
    dfree(c, c->rigmask);
    dfree(c, c->vpath_x);
    dfree(c, c->nrg_xmin);
    dfree(c, c->nrg_xmax);
    dfree(c, c->nrg_pack);
    dfree(c, c->fix_d);
    dfree(c, c->fixn_d);
    dfree(c, c->tail_d);
    dfree(c, c->far_d);
    dfree(c, c->err_d);
    dfree(c, c->dyn_d);
    drop_seam_graphs(c);
    if (c->cells_d) {
        unsigned long long n = 0;
        if (cudaMemcpyAsync(&n, c->cells_d, sizeof n, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess &&
            carver_sync(c) == cudaSuccess)
            g_update_cells += n;
    }
    dfree(c, c->cells_d);
    if (c->dbg_d) {
        long long v[32];
        if (cudaMemcpyAsync(v, c->dbg_d, sizeof v, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess &&
            carver_sync(c) == cudaSuccess) {
            fprintf(stderr, "b200c dbg:");
            for (int i = 0; i < 32; ++i) fprintf(stderr, " %lld", v[i]);
            fprintf(stderr, "\n");
        }
    }
    dfree(c, c->dbg_d);
    dfree(c, c->rigmap_d);
    pinned_release(c->host_out, c->host_out_cap);
    if (c->stream) carver_sync(c);
    lane_release(c->lane); // attached carvers gave theirs back when they joined the root's queue
    delete c;
}

int b200c_carver_init(B200Carver *c, int delta_x, float rigidity)
{
    if (!c || delta_x < 0) return fail(B200C_ERROR, "carver_init: bad arguments");
    if (c->active) return fail(B200C_ERROR, "carver_init: already active");
    B_TRY(use_device(c));
    if (delta_x > B200C_MAX_DELTA) return fail(B200C_ERROR, "carver_init: delta_x above the engine's limit (120)");
    c->delta_x = delta_x; // the tensor maps made by alloc_maps depend on both
    c->rigidity = rigidity;
    if (!c->nrg_active) B_TRY(init_energy_related(c));
    c->active = true;
    B_TRY(alloc_maps(c));
    B_TRY(dalloc(c, &c->vpath_x, (size_t) c->h, true));
    B_TRY(dalloc(c, &c->nrg_xmin, (size_t) c->h, true));
    B_TRY(dalloc(c, &c->nrg_xmax, (size_t) c->h, true));
    B_TRY(dalloc(c, &c->nrg_pack, (size_t) c->h, true));
    B_TRY(dalloc(c, &c->fix_d, (size_t) c->h / 8 + 4, true));
    B_TRY(dalloc(c, &c->fixn_d, 1, true));
    B_TRY(dalloc(c, &c->tail_d, 4, true));
    B_TRY(dalloc(c, &c->far_d, 1, true));
    B_TRY(dalloc(c, &c->err_d, 1, true));
    B_TRY(dalloc(c, &c->cells_d, 1, true));
    B_TRY(dalloc(c, &c->dyn_d, 1, true));
    if (getenv("B200C_DBG")) B_TRY(dalloc(c, &c->dbg_d, 32, true));
    c->delta_x = delta_x;
    c->rigidity = rigidity;
    c->rigmap_h.assign(2 * delta_x + 1, 0.f);
    for (int x = -delta_x; x <= delta_x; ++x)
        c->rigmap_h[x + delta_x] = c->rigidity * powf(fabsf((float) x), 1.5f) / c->h; // A.1 / A.4
    B_TRY(dalloc(c, &c->rigmap_d, c->rigmap_h.size(), false));
    B_TRY(upload_rigmap(c));
    return B200C_OK;
}

int b200c_carver_attach(B200Carver *root, B200Carver *aux)
{
    if (!root || !aux) return fail(B200C_ERROR, "carver_attach: NULL");
    if (root->w0 != aux->w0 || root->h0 != aux->h0) return fail(B200C_ERROR, "carver_attach: size mismatch");
    if (root->device != aux->device) return fail(B200C_ERROR, "carver_attach: carvers live on different devices");
    B_TRY(use_device(root));
    // the aux carver's own queue must be idle before it starts sharing the root's maps and stream
    CU_TRY(carver_sync(aux));
    dfree(aux, aux->vs);
    CU_TRY(carver_sync(aux));
    lane_release(aux->lane);
    aux->lane = nullptr;
    aux->stream = root->stream; // one queue per carver family keeps every structural op ordered
    aux->vs = root->vs;
    aux->root = root;
    root->attached.push_back(aux);
    return B200C_OK;
}

int b200c_carver_set_energy_function(B200Carver *c, int ef)
{
    if (!c) return fail(B200C_ERROR, "set_energy_function: NULL");
    int grad, rd, rad = 1;
    switch (ef) {
        case 0: grad = GRAD_NORM; rd = READ_BRIGHTNESS; break;
        case 1: grad = GRAD_SUMABS; rd = READ_BRIGHTNESS; break;
        case 2: grad = GRAD_XABS; rd = READ_BRIGHTNESS; break;
        case 3: grad = GRAD_NORM; rd = READ_LUMA; break;
        case 4: grad = GRAD_SUMABS; rd = READ_LUMA; break;
        case 5: grad = GRAD_XABS; rd = READ_LUMA; break;
        case 6: grad = GRAD_NULL; rd = READ_BRIGHTNESS; rad = 0; break;
        default: return fail(B200C_ERROR, "set_energy_function: unknown builtin");
    }
    c->ef = ef;
    c->grad_kind = grad;
    c->read_kind = rd;
    c->nrg_radius = rad;
    c->nrg_uptodate = false;
    return B200C_OK;
}

int b200c_carver_set_side_switch_frequency(B200Carver *c, unsigned int f)
{
    if (!c) return fail(B200C_ERROR, "set_side_switch_frequency: NULL");
    c->lr_freq = f;
    return B200C_OK;
}

static int mask_common(B200Carver *c, const unsigned char *rgb, int channels, int width, int height, int x_off,
                       int y_off, bool is_bias, int bias_factor)
{
    if (!c || !rgb || channels < 1 || width < 1 || height < 1) return fail(B200C_ERROR, "mask: bad arguments");
    B_TRY(use_device(c));
    if (!is_bias && !c->active) return fail(B200C_ERROR, "rigmask: carver not initialised");
    if (not_at_reference(c)) B_TRY(flatten(c));
    if (is_bias && bias_factor == 0) return B200C_OK;
    const size_t n = (size_t) c->w0 * c->h0;
    if (is_bias && !c->bias) B_TRY(dalloc(c, &c->bias, n, true));
    if (!is_bias && !c->rigmask) B_TRY(dalloc(c, &c->rigmask, n, true));
    const int was_transposed = c->transposed;
    if (was_transposed) B_TRY(transpose(c));

    const int x0 = x_off < 0 ? x_off : 0, y0 = y_off < 0 ? y_off : 0;
    const int x1 = x_off > 0 ? x_off : 0, y1 = y_off > 0 ? y_off : 0;
    const int x2 = c->w < width + x_off ? c->w : width + x_off;
    const int y2 = c->h < height + y_off ? c->h : height + y_off;
    const int nx = x2 - x1, ny = y2 - y1;
    if (nx > 0 && ny > 0) {
        uint8_t *d_mask = nullptr;
        const size_t bytes = (size_t) width * height * channels;
        B_TRY(dalloc(c, &d_mask, bytes, false));
        CU_TRY(cudaMemcpyAsync(d_mask, rgb, bytes, cudaMemcpyHostToDevice, c->stream));
        dim3 grid((nx + 255) / 256, ny);
        {
            StageScope sc("mask", c->stream);
            if (is_bias)
                k_bias_add<<<grid, 256, 0, c->stream>>>(c->bias, c->w0, d_mask, channels, width, bias_factor, x0, y0,
                                                        x1, y1, nx, ny);
            else
                k_rigmask_set<<<grid, 256, 0, c->stream>>>(c->rigmask, c->w0, d_mask, channels, width, x0, y0, x1, y1,
                                                           nx, ny);
            B_TRY(check_launch("mask kernel"));
        }
        dfree(c, d_mask);
        CU_TRY(carver_sync(c)); // the caller frees the mask right after (io_functions.c:97,128)
    }
    if (is_bias) c->nrg_uptodate = false;
    if (was_transposed != c->transposed) B_TRY(transpose(c));
    return B200C_OK;
}

int b200c_carver_bias_add_rgb_area(B200Carver *c, const unsigned char *rgb, int bias_factor, int channels, int width,
                                   int height, int x_off, int y_off)
{
    return mask_common(c, rgb, channels, width, height, x_off, y_off, true, bias_factor);
}

int b200c_carver_rigmask_add_rgb_area(B200Carver *c, const unsigned char *rgb, int channels, int width, int height,
                                      int x_off, int y_off)
{
    return mask_common(c, rgb, channels, width, height, x_off, y_off, false, 0);
}

int b200c_carver_build_maps(B200Carver *c, int depth, int update_step, b200c_progress_fn progress, void *user)
{
    return b200c_carver_build_maps_phase(c, depth, update_step, 0, progress, user);
}

int b200c_carver_seams_done(const B200Carver *c)
{
    if (!c || !c->lane || !c->lane->seams_h) return -1;
    return *reinterpret_cast<volatile int *>(c->lane->seams_h);
}

int b200c_carver_build_maps_phase(B200Carver *c, int depth, int update_step, int update_phase, b200c_progress_fn progress,
                                  void *user)
{
    if (!c) return fail(B200C_ERROR, "build_maps: NULL");
    if (depth <= c->max_level) return B200C_OK;
    if (!c->active) return fail(B200C_ERROR, "build_maps: carver not initialised");
    if (c->root) return fail(B200C_ERROR, "build_maps: attached carvers cannot be resized directly");
    B_TRY(use_device(c));
    set_width_one(c, c->w_start - c->max_level + 1);
    B_TRY(build_emap(c));
    B_TRY(gather_rig(c));
    B_TRY(build_mmap(c));
    B_TRY(build_vsmap(c, depth, update_step, progress, user, true, update_phase < 0 ? 0 : update_phase));
    return B200C_OK;
}

// A batch of independent images (the reference's own batch use, batch/batch-gimp-lqr.scm:19-66): the
// same build_maps session for every carver, advanced in lockstep by ONE launch per step on the first carver's queue (image
// = blockIdx.z, argument blocks in a table in HBM), so the row-serial chains of the images run side by side on the SMs
// and one host thread drives them all.  The carvers must agree in geometry and knobs (same size, delta_x, level, side
// switching, no rigidity); anything else is carved one by one.
int b200c_batch_build_maps(B200Carver **cs, int n, int depth)
{
    if (!cs || n < 1) return fail(B200C_ERROR, "batch_build_maps: bad arguments");
    B200Carver *L = cs[0];
    bool same = true;
    for (int i = 0; i < n; ++i) {
        const B200Carver *m = cs[i];
        if (!m) return fail(B200C_ERROR, "batch_build_maps: NULL carver");
        same = same && m->active && !m->root && m->device == L->device && m->w == L->w && m->h == L->h && m->w0 == L->w0 &&
               m->h0 == L->h0 && m->w_start == L->w_start && m->h_start == L->h_start && m->level == L->level &&
               m->max_level == L->max_level && m->delta_x == L->delta_x && m->rigidity == 0.f && m->leftright == L->leftright &&
               m->lr_freq == L->lr_freq && m->generic == L->generic && m->use_trace == L->use_trace &&
               m->use_graph == L->use_graph && m->bd_maxseg == L->bd_maxseg && m->mates.empty();
    }
    if (!same || n == 1) {
        for (int i = 0; i < n; ++i) B_TRY(b200c_carver_build_maps(cs[i], depth, 1, nullptr, nullptr));
        return B200C_OK;
    }
    if (depth <= L->max_level) return B200C_OK;
    B_TRY(use_device(L));
    // every mate's own queue must be idle before its buffers are used from the leader's
    std::vector<cudaStream_t> own(n);
    for (int i = 0; i < n; ++i) {
        CU_TRY(carver_sync(cs[i]));
        own[i] = cs[i]->stream;
    }
    int rc = B200C_OK;
    const bool tail0 = L->use_tail;
    L->mates.assign(cs + 1, cs + n);
    L->tab_n = n;
    for (int i = 0; i < n; ++i) {
        cs[i]->stream = L->stream;
        for (B200Carver *a : cs[i]->attached) a->stream = L->stream;
    }
    if (pool_alloc(L, (void **) &L->tab_d, 2 * (size_t) n * sizeof(DevP), L->stream) != cudaSuccess ||
        pool_alloc(L, (void **) &L->mtab_d, (size_t) n * sizeof(BdMaps), L->stream) != cudaSuccess)
        rc = fail(B200C_NOMEM, "batch_build_maps: table allocation", cudaGetLastError());
    if (rc == B200C_OK) {
        for_batch(L, [](B200Carver *m) { set_width_one(m, m->w_start - m->max_level + 1); });
        rc = build_emap(L);
    }
    if (rc == B200C_OK) rc = build_mmap(L);
    if (rc == B200C_OK) rc = build_vsmap(L, depth, 1, nullptr, nullptr, true);
    const cudaError_t es = carver_sync(L);
    if (es != cudaSuccess && rc == B200C_OK) rc = fail(B200C_ERROR, "batch_build_maps: sync", es);
    if (L->tab_d) cudaFreeAsync(L->tab_d, L->stream);
    if (L->mtab_d) cudaFreeAsync(L->mtab_d, L->stream);
    L->tab_d = nullptr, L->mtab_d = nullptr, L->tab_n = 0;
    for (int i = 0; i < n; ++i) {
        cs[i]->stream = own[i];
        for (B200Carver *a : cs[i]->attached) a->stream = own[i];
        cs[i]->use_tail = tail0;
    }
    L->mates.clear();
    drop_seam_graphs(L);
    return rc;
}

int b200c_carver_set_width(B200Carver *c, int w1)
{
    if (!c) return fail(B200C_ERROR, "set_width: NULL");
    set_width_rec(c, w1);
    return B200C_OK;
}

int b200c_carver_flatten(B200Carver *c)
{
    if (!c) return fail(B200C_ERROR, "flatten: NULL");
    B_TRY(use_device(c));
    return flatten(c);
}

int b200c_carver_transpose(B200Carver *c)
{
    if (!c) return fail(B200C_ERROR, "transpose: NULL");
    B_TRY(use_device(c));
    return transpose(c);
}

int b200c_carver_get(const B200Carver *c, int field)
{
    if (!c) return -1;
    switch (field) {
        case B200C_W: return c->w;
        case B200C_H: return c->h;
        case B200C_W0: return c->w0;
        case B200C_H0: return c->h0;
        case B200C_W_START: return c->w_start;
        case B200C_H_START: return c->h_start;
        case B200C_LEVEL: return c->level;
        case B200C_MAX_LEVEL: return c->max_level;
        case B200C_TRANSPOSED: return c->transposed;
        case B200C_CHANNELS: return c->channels;
        case B200C_ACTIVE: return c->active ? 1 : 0;
        case B200C_LEFTRIGHT: return c->leftright;
        case B200C_DEVICE: return c->device;
        default: return -1;
    }
}

static int readout_to(B200Carver *c, uint8_t *d_out)
{
    StageScope sc("readout", c->stream);
    k_compact_rows<<<c->h, B200C_ROW_THREADS, 0, c->stream>>>(c->rgb, c->vs, c->w0, c->w, c->channels, c->level, d_out,
                                                              nullptr, nullptr, nullptr, nullptr);
    return check_launch("k_compact_rows(readout)");
}

static int readout_impl(B200Carver *c, const unsigned char **host_pixels, bool chunked)
{
    if (!c || !host_pixels) return fail(B200C_ERROR, "readout: NULL");
    HostScope hs(10);
    B_TRY(use_device(c));
    const size_t row_bytes = (size_t) c->w * c->channels, bytes = row_bytes * c->h;
    B_TRY(ensure_host_out(c, bytes));
    uint8_t *d_out = nullptr;
    B_TRY(dalloc(c, &d_out, bytes, false));
    B_TRY(readout_to(c, d_out));
    // Large images come back in up to four chunks of rows: the call returns when the FIRST has arrived, the caller's
    // scan (b200c_carver_readout_rows) waits for the others as it reaches them, so the copy of the later rows overlaps
    // the caller's handling of the earlier ones.  A crowded host (many carvers in flight) takes one copy and sleeps.
    Lane *lane = c->lane ? c->lane : c->root->lane;
    const int chunks = (chunked && bytes >= (4u << 20) && c->h >= 4 && !host_crowded()) ? 4 : 1;
    c->rd_rows = (c->h + chunks - 1) / chunks;
    c->rd_chunks = (c->h + c->rd_rows - 1) / c->rd_rows;
    c->rd_ready = 0;
    for (int i = 0; i < c->rd_chunks; ++i) {
        const size_t off = (size_t) i * c->rd_rows * row_bytes;
        const size_t n = (off + (size_t) c->rd_rows * row_bytes <= bytes) ? (size_t) c->rd_rows * row_bytes : bytes - off;
        CU_TRY(cudaMemcpyAsync(c->host_out + off, d_out + off, n, cudaMemcpyDeviceToHost, c->stream));
        if (c->rd_chunks > 1) CU_TRY(cudaEventRecord(lane->rd[i], c->stream));
    }
    dfree(c, d_out);
    if (c->rd_chunks > 1) {
        CU_TRY(cudaEventSynchronize(lane->rd[0]));
        c->rd_ready = 1;
    } else {
        CU_TRY(carver_sync(c));
        c->rd_ready = c->rd_chunks;
    }
    *host_pixels = c->host_out;
    return B200C_OK;
}

int b200c_carver_readout(B200Carver *c, const unsigned char **host_pixels) { return readout_impl(c, host_pixels, false); }
int b200c_carver_readout_begin(B200Carver *c, const unsigned char **host_pixels) { return readout_impl(c, host_pixels, true); }

// rows of the last read-out that have arrived in the host buffer, after waiting for the chunk that holds `row`
int b200c_carver_readout_rows(B200Carver *c, int row)
{
    if (!c) return -1;
    if (c->rd_ready < c->rd_chunks && row >= c->rd_ready * c->rd_rows) {
        if (use_device(c) != B200C_OK) return -1;
        Lane *lane = c->lane ? c->lane : c->root->lane;
        while (c->rd_ready < c->rd_chunks && row >= c->rd_ready * c->rd_rows) {
            if (cudaEventSynchronize(lane->rd[c->rd_ready]) != cudaSuccess) {
                fail(B200C_ERROR, "readout_rows: event", cudaGetLastError());
                return -1;
            }
            c->rd_ready++;
        }
    }
    const int rows = c->rd_ready * c->rd_rows;
    return rows < c->h ? rows : c->h;
}

int b200c_carver_readout_device(B200Carver *c, void *d_out)
{
    if (!c || !d_out) return fail(B200C_ERROR, "readout_device: NULL");
    B_TRY(use_device(c));
    return readout_to(c, (uint8_t *) d_out);
}

int b200c_carver_vmap(B200Carver *c, int *out_host)
{
    if (!c || !out_host) return fail(B200C_ERROR, "vmap: NULL");
    B_TRY(use_device(c));
    const int depth = c->w0 - c->w_start;
    const size_t n = (size_t) c->w_start * c->h;
    int *d_out = nullptr;
    B_TRY(dalloc(c, &d_out, n, true));
    {
        StageScope sc("vmap", c->stream);
        k_vmap_rows<<<c->h, B200C_ROW_THREADS, 0, c->stream>>>(c->vs, c->w0, c->w_start, c->h, depth, c->transposed,
                                                               d_out);
        B_TRY(check_launch("k_vmap_rows"));
    }
    CU_TRY(cudaMemcpyAsync(out_host, d_out, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    dfree(c, d_out);
    CU_TRY(carver_sync(c));
    return B200C_OK;
}

// ---- the plug-in's own loops next to the hot path; no carver involved: they borrow a lane
namespace {
struct ScratchLane {
    Lane *lane = nullptr;
    int open()
    {
        const int device = g_device >= 0 ? g_device : g_device_tls_default;
        if (cudaSetDevice(device) != cudaSuccess) return fail(B200C_ERROR, "cudaSetDevice", cudaGetLastError());
        lane = lane_acquire(device, false, nullptr);
        return lane ? B200C_OK : fail(B200C_ERROR, "cannot create stream", cudaGetLastError());
    }
    ~ScratchLane()
    {
        if (!lane) return;
        cudaStreamSynchronize(lane->stream);
        lane_release(lane);
    }
};
} // namespace

int b200c_vmap_colour(const int *vmap, int w, int h, int depth, const double colour_start[3], const double colour_end[3],
                      unsigned char *out_rgba)
{
    if (!vmap || !colour_start || !colour_end || !out_rgba || w < 1 || h < 1 || depth < 0)
        return fail(B200C_ERROR, "vmap_colour: bad arguments");
    ScratchLane sl;
    B_TRY(sl.open());
    cudaStream_t s = sl.lane->stream;
    const size_t n = (size_t) w * h;
    int *d_in = nullptr;
    uchar4 *d_out = nullptr;
    cudaError_t e = cudaMallocAsync((void **) &d_in, n * sizeof(int), s);
    if (e == cudaSuccess) e = cudaMallocAsync((void **) &d_out, n * sizeof(uchar4), s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, vmap, n * sizeof(int), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        StageScope sc("vmap_colour", s);
        k_vmap_colour<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(d_in, n, depth, colour_start[0], colour_start[1],
                                                                  colour_start[2], colour_end[0], colour_end[1],
                                                                  colour_end[2], d_out);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_rgba, d_out, n * sizeof(uchar4), cudaMemcpyDeviceToHost, s);
    if (d_in) cudaFreeAsync(d_in, s);
    if (d_out) cudaFreeAsync(d_out, s);
    const cudaError_t e2 = cudaStreamSynchronize(s);
    if (e != cudaSuccess || e2 != cudaSuccess) return fail(B200C_ERROR, "vmap_colour", e != cudaSuccess ? e : e2);
    return B200C_OK;
}

int b200c_guess_new_size(const unsigned char *mask, int width, int height, int bpp, int has_alpha, int x_off, int y_off,
                         int old_width, int old_height, int direction, int *new_size)
{
    if (!mask || !new_size || width < 1 || height < 1 || bpp < 1 || bpp > 4 || bpp - (has_alpha ? 1 : 0) < 1 ||
        (direction != 0 && direction != 1))
        return fail(B200C_ERROR, "guess_new_size: bad arguments");
    // the part of the mask that lies over the layer (layers_combo.c:324-338)
    const int x_lo = std::max(0, x_off), x_hi = std::min(old_width, width + x_off);
    const int y_lo = std::max(0, y_off), y_hi = std::min(old_height, height + y_off);
    const int old_size = direction == 0 ? old_width : old_height;
    const int nlines = direction == 0 ? y_hi - y_lo : x_hi - x_lo;
    const int count = direction == 0 ? x_hi - x_lo : y_hi - y_lo;
    *new_size = old_size;
    if (nlines <= 0 || count <= 0) return B200C_OK; // no overlap: nothing to discard
    // first line / first pixel of a line in mask coordinates
    const int line0 = direction == 0 ? y_lo - y_off : x_lo - x_off;
    const int first = direction == 0 ? std::max(0, -x_off) : std::max(0, -y_off);
    ScratchLane sl;
    B_TRY(sl.open());
    cudaStream_t s = sl.lane->stream;
    const size_t bytes = (size_t) width * height * bpp;
    unsigned char *d_mask = nullptr;
    int *d_res = nullptr;
    int res = 0;
    cudaError_t e = cudaMallocAsync((void **) &d_mask, bytes, s);
    if (e == cudaSuccess) e = cudaMallocAsync((void **) &d_res, sizeof(int), s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_mask, mask, bytes, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_res, 0, sizeof(int), s);
    if (e == cudaSuccess) {
        StageScope sc("guess_new_size", s);
        k_guess_mask_size<<<(nlines + 7) / 8, 256, 0, s>>>(d_mask, width, bpp, has_alpha, line0, nlines, first, count,
                                                          direction, d_res);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&res, d_res, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (d_mask) cudaFreeAsync(d_mask, s);
    if (d_res) cudaFreeAsync(d_res, s);
    const cudaError_t e2 = cudaStreamSynchronize(s);
    if (e != cudaSuccess || e2 != cudaSuccess) return fail(B200C_ERROR, "guess_new_size", e != cudaSuccess ? e : e2);
    *new_size = old_size - res;
    return B200C_OK;
}

int b200c_carver_true_energy(B200Carver *c, float *out_host)
{
    if (!c || !out_host) return fail(B200C_ERROR, "true_energy: NULL");
    B_TRY(use_device(c));
    if (!c->nrg_active) B_TRY(init_energy_related(c));
    B_TRY(build_emap(c));
    const size_t n = (size_t) c->w * c->h;
    float *d_out = nullptr;
    B_TRY(dalloc(c, &d_out, n, false));
    dim3 grid((c->w + 255) / 256, c->h);
    {
        StageScope sc("energy_export", c->stream);
        k_energy_export<<<grid, 256, 0, c->stream>>>(view(c), c->transposed, d_out);
        B_TRY(check_launch("k_energy_export"));
    }
    CU_TRY(cudaMemcpyAsync(out_host, d_out, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    dfree(c, d_out);
    CU_TRY(carver_sync(c));
    return B200C_OK;
}

int b200c_set_stream(void *stream)
{
    g_ext_stream = (cudaStream_t) stream;
    g_use_ext_stream = stream != nullptr;
    return B200C_OK;
}

void b200c_set_timing(int on) { g_timing = on != 0; }

unsigned long long b200c_update_cells(void) { return g_update_cells.load(); }

int b200c_carver_sync(B200Carver *c)
{
    if (!c) return fail(B200C_ERROR, "sync: NULL");
    B_TRY(use_device(c));
    CU_TRY(carver_sync(c));
    return B200C_OK;
}

// ---- test / profiling hooks -----------------------------------------------------------------------------------
int b200c_debug_build(B200Carver *c, int n_seams)
{
    if (!c || !c->active || c->root) return fail(B200C_ERROR, "debug_build: need an initialised root carver");
    B_TRY(use_device(c));
    set_width_one(c, c->w_start - c->max_level + 1);
    B_TRY(build_emap(c));
    B_TRY(gather_rig(c));
    B_TRY(build_mmap(c));
    if (n_seams > 0) B_TRY(build_vsmap(c, c->max_level + n_seams, 1, nullptr, nullptr, false));
    CU_TRY(carver_sync(c));
    return B200C_OK;
}

long b200c_debug_fetch(B200Carver *c, int what, void *out, long cap)
{
    if (!c || !out) return -1;
    if (use_device(c) != B200C_OK) return -1;
    const void *src = nullptr;
    long n = (long) c->w0 * c->h0;
    switch (what) {
        case B200C_DBG_EN:
        case B200C_DBG_M:
        case B200C_DBG_LEAST: {
            // the compact maps, scattered back to the physical (pixel id) layout liblqr keeps them in
            if (!c->raw || !c->en || (what != B200C_DBG_EN && !c->m)) return 0;
            int *tmp = nullptr;
            if (dalloc(c, &tmp, (size_t) n, true) != B200C_OK) return -1;
            dim3 grid((c->w + 255) / 256, c->h);
            k_export_physical<<<grid, 256, 0, c->stream>>>(view(c), what - B200C_DBG_EN, tmp);
            if (n > cap) n = cap;
            const bool ok = cudaMemcpyAsync(out, tmp, (size_t) n * 4, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess &&
                            carver_sync(c) == cudaSuccess;
            dfree(c, tmp);
            return ok ? n : -1;
        }
        case B200C_DBG_RAW: src = c->raw; n = (long) c->w_start * c->h_start; break;
        case B200C_DBG_VS: src = c->vs; break;
        case B200C_DBG_VPATH_X: src = c->vpath_x; n = c->h; break;
        case B200C_DBG_BIAS: src = c->bias; break;
        case B200C_DBG_RIGMASK: src = c->rigmask; break;
        default: return -1;
    }
    if (!src) return 0;
    if (n > cap) n = cap;
    if (carver_sync(c) != cudaSuccess) return -1;
    if (cudaMemcpy(out, src, (size_t) n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return n;
}

long b200c_launch_count(void) { return g_launches.load(); }

double b200c_stage_ms(const char *stage, long *launches)
{
    std::lock_guard<std::mutex> lk(g_stage_mu);
    auto it = g_stages.find(stage ? stage : "");
    if (it == g_stages.end()) {
        if (launches) *launches = 0;
        return 0.0;
    }
    drain_stage(it->second);
    if (launches) *launches = it->second.launches;
    return it->second.ms;
}

void b200c_stage_reset(void)
{
    std::lock_guard<std::mutex> lk(g_stage_mu);
    for (auto &kv : g_stages) drain_stage(kv.second);
    g_stages.clear();
}

} // extern "C"
