// ptg_capi.cu -- the extern "C" boundary declared in include/ptg_b200.h.
//
// Argument validation, method selection, per-device cached state (SM count, streams,
// staging ring for the host-buffer entry points) and the H2D -> deposit -> D2H
// pipeline.  No CUDA or C++ type crosses the boundary; nothing here throws.
#include <cmath>
#include <cstddef>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/ptg_b200.h"
#include "ptg_internal.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define PTG_CUDA(call)                                                                  \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess)                                                         \
            return fail(e__ == cudaErrorMemoryAllocation ? PTG_ENOMEM : PTG_ECUDA,      \
                        "%s failed: %s", #call, cudaGetErrorString(e__));               \
    } while (0)

constexpr int kMaxDevices = 16;
constexpr int kSlots = 3;                         // staging ring depth
constexpr int64_t kChunk = int64_t(1) << 22;      // particles per pipeline chunk (32 MiB per array)
constexpr int64_t kStageChunk = int64_t(1) << 21; // ... when staging through pinned memory (16 MiB per array)

struct DeviceState {
    std::atomic<bool> probed{false};
    bool usable = false;
    std::mutex mu;                                // serialises the host-buffer entry points of THIS device
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    cudaMemPool_t pool = nullptr;                 // the library's private stream-ordered pool on this device
    // host-entry staging
    cudaStream_t stream[kSlots] = {};
    double *slot[kSlots] = {};                    // each 4 * kChunk doubles
    double *grid = nullptr;
    int64_t grid_cells = 0;
    bool staged = false;
    // pinned host ring for pageable / strided sources (filled by worker threads, then DMA'd)
    double *pinned[kSlots] = {};
    cudaEvent_t h2d_done[kSlots] = {};
    bool pinned_ready = false;
    // AUTO: the candidate kernels that the probe's verdict switches off run on side streams next to the one that works
    cudaStream_t aux[3] = {};
    cudaEvent_t ev[64] = {};
    std::atomic<unsigned> ev_next{0};
    std::atomic<bool> aux_ready{false};
    // ring of probe verdicts for the TILED / AUTO methods (one entry per call in flight)
    std::atomic<ptg::TileConfig *> cfg_ring{nullptr};
    std::atomic<unsigned> cfg_next{0};
};

DeviceState g_dev[kMaxDevices];
std::mutex g_mu;          // first-use probing of a device only

// a CUDA event that is destroyed on every exit path
struct Event {
    cudaEvent_t e = nullptr;
    cudaError_t create() { return cudaEventCreateWithFlags(&e, cudaEventDisableTiming); }
    ~Event() { if (e) cudaEventDestroy(e); }
    Event() = default;
    Event(const Event &) = delete;
    Event &operator=(const Event &) = delete;
};

}  // namespace

namespace ptg {
cudaError_t scratch_alloc(void **ptr, size_t bytes, cudaStream_t st) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    cudaMemPool_t pool = (dev >= 0 && dev < kMaxDevices) ? g_dev[dev].pool : nullptr;
    return pool ? cudaMallocFromPoolAsync(ptr, bytes, pool, st) : cudaMallocAsync(ptr, bytes, st);
}
}  // namespace ptg

namespace {

int probe(int device, DeviceState **out) {
    if (device < 0) {
        cudaError_t e = cudaGetDevice(&device);
        if (e != cudaSuccess) return fail(PTG_ENODEVICE, "no CUDA device: %s", cudaGetErrorString(e));
    }
    if (device >= kMaxDevices) return fail(PTG_EINVAL, "device index %d too large", device);
    DeviceState &d = g_dev[device];
    if (!d.probed.load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!d.probed.load(std::memory_order_relaxed)) {
            cudaDeviceProp p;
            cudaError_t e = cudaGetDeviceProperties(&p, device);
            if (e != cudaSuccess) {
                (void)cudaGetLastError();
                return fail(PTG_ENODEVICE, "no usable CUDA device %d: %s", device, cudaGetErrorString(e));
            }
            d.sm_count = p.multiProcessorCount;
            d.cc_major = p.major;
            d.cc_minor = p.minor;
            d.usable = (p.major == 10 && p.minor == 0);   // the library holds sm_100a code only (not sm_103)
            // stream-ordered scratch (staggered staging grid, SORTED workspace) comes from a memory pool of
            // the library's own: freed blocks stay cached there instead of going back to the driver at every
            // synchronisation, and the device's default pool -- which torch's cudaMallocAsync backend or
            // anything else in the process may use -- keeps its settings
            cudaMemPoolProps props;
            memset(&props, 0, sizeof props);
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = device;
            if (cudaMemPoolCreate(&d.pool, &props) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                (void)cudaMemPoolSetAttribute(d.pool, cudaMemPoolAttrReleaseThreshold, &keep);
            } else {
                d.pool = nullptr;                         // fall back to the default pool, untouched
            }
            (void)cudaGetLastError();
            d.probed.store(true, std::memory_order_release);
        }
    }
    if (!d.usable)
        return fail(PTG_ENODEVICE, "device %d is sm_%d%d; libptg_b200 contains sm_100a code only", device,
                    d.cc_major, d.cc_minor);
    *out = &d;
    return PTG_OK;
}

int current_device_state(DeviceState **out) { return probe(-1, out); }

int validate(int ndim, int shape, int64_t n, const double *x, const double *y, const double *z,
             const double *lo, const double *inv, const int *bins, const double *grid) {
    if (ndim < 1 || ndim > 3) return fail(PTG_EINVAL, "ndim must be 1, 2 or 3 (got %d)", ndim);
    if (shape < 0 || shape > 3) return fail(PTG_EINVAL, "unknown particle shape %d", shape);
    if (n < 0) return fail(PTG_EINVAL, "negative particle count");
    if (!lo || !inv || !bins || !grid) return fail(PTG_EINVAL, "lo/inv/bins/grid must not be NULL");
    if (n > 0 && (!x || (ndim >= 2 && !y) || (ndim >= 3 && !z)))
        return fail(PTG_EINVAL, "coordinate pointer is NULL");
    for (int a = 0; a < ndim; ++a)
        if (bins[a] < 1) return fail(PTG_EINVAL, "bins[%d] = %d must be >= 1", a, bins[a]);
    const int64_t cells = ptg_padded_cells(ndim, shape, bins);
    if (cells <= 0 || cells >= (int64_t(1) << 31))
        return fail(PTG_EUNSUPPORTED, "padded grid has %lld cells; supported: < 2^31", (long long)cells);
    return PTG_OK;
}

void fill_params(ptg::DepositParams &P, int ndim, int shape, int64_t n, const double *x, const double *y,
                 const double *z, const double *w, const int64_t *strides, const double *lo,
                 const double *inv, const int *bins, double *grid) {
    memset(&P, 0, sizeof P);
    P.p[0] = x; P.p[1] = y; P.p[2] = z; P.p[3] = w;
    for (int k = 0; k < 4; ++k) P.stride[k] = strides ? strides[k] : 1;
    P.n = n;
    for (int a = 0; a < 3; ++a) {
        const bool on = a < ndim;
        P.lo[a] = on ? lo[a] : 0.0;
        P.inv[a] = on ? inv[a] : 1.0;
        P.bins[a] = on ? bins[a] : 1;
        P.fbins[a] = (double)P.bins[a];
        P.fbins_p1[a] = (double)P.bins[a] + 1.0;
        P.dims[a] = on ? bins[a] + 2 * shape : 1;
    }
    P.grid = grid;
    P.slab_lo = -(1 << 30);
    P.slab_hi = 1 << 30;
    P.stag = nullptr;
    P.stag_pitch = 0;
    P.stag_copy = 0;
    P.fuse_kind = 0;            // (memset above already cleared q / qstride / f_* / akind / pos)
    P.f_weight = 1.0;
}

struct FuseSpec {
    int kind;
    const double *q[3];
    int64_t qstride[3];
    double mass, c, qe;
    // axis expressions and the atomic arrays / constants they may read
    int akind[3] = {0, 0, 0};
    const double *pos[3] = {nullptr, nullptr, nullptr};
    int64_t pos_stride[3] = {1, 1, 1};
    double charge = 0.0, time = 0.0, amu = 1.66053906660e-27, weight_const = 1.0;
};

// which atomic arrays scalar `kind` reads: bit 0..2 = x, y, z; bit 3 = the momenta
unsigned scalar_needs(int kind) {
    if (kind >= 1 && kind <= 6) return 8u;
    switch (kind) {
        case PTG_S_X: case PTG_S_X_UM: return 1u;
        case PTG_S_Y: case PTG_S_Y_UM: return 2u;
        case PTG_S_Z: case PTG_S_Z_UM: return 4u;
        case PTG_S_R_XY: return 3u;
        case PTG_S_R_YZ: return 6u;
        case PTG_S_R_ZX: return 5u;
        case PTG_S_R_XYZ: return 7u;
        case PTG_S_WEIGHT: case PTG_S_MASS: case PTG_S_CHARGE: case PTG_S_TIME: case PTG_S_MASS_U: case PTG_S_CHARGE_E:
        case PTG_S_ERUHE: return 0u;
        default: return 8u;
    }
}

struct MaskSpec {
    const unsigned char *mask;
    int64_t stride;
};

// Probe verdicts live in a ring: an AUTO/TILED call takes the next entry and its kernels read it in
// stream order.  4096 entries = the number of such calls that may be in flight per device at once
// (the host pipeline below never has more than 8).
constexpr unsigned kCfgRing = 4096;

std::mutex g_ring_mu;     // taken while a device mutex may be held, never the other way round

int next_tile_config(DeviceState *d, ptg::TileConfig **out) {
    ptg::TileConfig *ring = d->cfg_ring.load(std::memory_order_acquire);
    if (!ring) {
        std::lock_guard<std::mutex> lk(g_ring_mu);
        ring = d->cfg_ring.load(std::memory_order_relaxed);
        if (!ring) {
            PTG_CUDA(cudaMalloc(&ring, kCfgRing * sizeof(ptg::TileConfig)));
            cudaError_t e = cudaMemset(ring, 0, kCfgRing * sizeof(ptg::TileConfig));   // tickets start at 0
            if (e != cudaSuccess) { cudaFree(ring); PTG_CUDA(e); }
            d->cfg_ring.store(ring, std::memory_order_release);
        }
    }
    *out = ring + (d->cfg_next.fetch_add(1) % kCfgRing);
    return PTG_OK;
}

int ensure_aux(DeviceState *d) {
    if (d->aux_ready.load(std::memory_order_acquire)) return PTG_OK;
    std::lock_guard<std::mutex> lk(g_ring_mu);
    if (d->aux_ready.load(std::memory_order_relaxed)) return PTG_OK;
    for (int k = 0; k < 3; ++k)
        if (!d->aux[k]) PTG_CUDA(cudaStreamCreateWithFlags(&d->aux[k], cudaStreamNonBlocking));
    for (int k = 0; k < 64; ++k)
        if (!d->ev[k]) PTG_CUDA(cudaEventCreateWithFlags(&d->ev[k], cudaEventDisableTiming));
    d->aux_ready.store(true, std::memory_order_release);
    return PTG_OK;
}

cudaEvent_t next_event(DeviceState *d) { return d->ev[d->ev_next.fetch_add(1) % 64]; }

int deposit_device(DeviceState *d, int ndim, int shape, int64_t n, const double *x, const double *y,
                   const double *z, const double *w, const int64_t *strides, const double *lo,
                   const double *inv, const int *bins, double *grid, unsigned flags, cudaStream_t st,
                   void *workspace = nullptr, size_t workspace_bytes = 0, const FuseSpec *fuse = nullptr,
                   const MaskSpec *mask = nullptr) {
    ptg::DepositParams P;
    fill_params(P, ndim, shape, n, x, y, z, w, strides, lo, inv, bins, grid);
    if (mask && mask->mask) {
        P.mask = mask->mask;
        P.mstride = mask->stride;
    }
    if (fuse) {
        P.fuse_kind = fuse->kind;
        for (int k = 0; k < 3; ++k) { P.q[k] = fuse->q[k]; P.qstride[k] = fuse->qstride[k]; }
        P.f_mass = fuse->mass;
        P.f_mc2 = (fuse->mass * fuse->c) * (fuse->mass * fuse->c);
        P.f_c2 = fuse->c * fuse->c;
        P.f_qe = fuse->qe;
        P.f_c = fuse->c;
        P.f_charge = fuse->charge;
        P.f_time = fuse->time;
        P.f_amu = fuse->amu;
        P.f_weight = fuse->weight_const;
        for (int a = 0; a < 3; ++a) {
            P.akind[a] = a < ndim ? fuse->akind[a] : 0;
            P.any_akind |= P.akind[a];
            P.pos[a] = fuse->pos[a];
            P.pos_stride[a] = fuse->pos_stride[a];
        }
    }
    unsigned method = flags & PTG_METHOD_MASK;
    if (method == PTG_METHOD_AUTO && !(flags & PTG_EXACT_ACCUM) && ptg::oneshot_applies(P, d->sm_count)) {
        // small call on a small grid: everything (clearing the grid included) in one launch
        ptg::TileConfig *cfg;
        int rc = next_tile_config(d, &cfg);
        if (rc) return rc;
        PTG_CUDA(ptg::deposit_oneshot(ndim, shape, P, cfg, (flags & PTG_ZERO_GRID) != 0, d->sm_count, st));
        return PTG_OK;
    }
    if (flags & PTG_ZERO_GRID)
        PTG_CUDA(cudaMemsetAsync(grid, 0, (size_t)ptg_padded_cells(ndim, shape, bins) * sizeof(double), st));
    if (method == PTG_METHOD_AUTO && n < (int64_t(1) << 21)) {
        // small calls are bound by launch latency, not by the GPU: skip the probe and the gated
        // launches.  A grid that fits the shared-memory tile is privatised, anything else goes
        // through the L2 reductions.
        method = ptg_padded_cells(ndim, shape, bins) <= ptg::kTileCells ? PTG_METHOD_TILED : PTG_METHOD_ROWS;
    }
    switch (method) {
        case PTG_METHOD_GLOBAL:
            PTG_CUDA(ptg::deposit_global(ndim, shape, P, false, nullptr, d->sm_count, st));
            break;
        case PTG_METHOD_ROWS:
            PTG_CUDA(ptg::deposit_global(ndim, shape, P, shape > 0, nullptr, d->sm_count, st));
            break;
        case PTG_METHOD_TILED:
        case PTG_METHOD_AUTO: {
            // probe -> {tiled | runs | runs_wide | rows}: all candidates are launched back to back, the
            // probe's verdict in device memory lets exactly one of them do the work and the others
            // return at once; no host round trip, the call stays asynchronous.
            if (n <= 0) break;
            ptg::TileConfig *cfg;
            int rc = next_tile_config(d, &cfg);
            if (rc) return rc;
            const bool force = method == PTG_METHOD_TILED;
            if (force) {
                PTG_CUDA(ptg::deposit_tiled(ndim, shape, P, cfg, true, (flags & PTG_EXACT_ACCUM) != 0, d->sm_count, st));
                break;
            }
            // The probe runs on the caller's stream; the four candidates fork from it -- the tile kernel stays on the
            // caller's stream, RUNS / RUNS_WIDE / ROWS go to side streams -- and join again: the three that the verdict
            // switches off return at once NEXT TO the one that works instead of queueing up behind it (their launch
            // gaps were ~25 us of every AUTO call, 4 % of a 1e8-particle step on ordered input).
            if ((rc = ensure_aux(d))) return rc;
            cudaEvent_t probed = next_event(d);
            PTG_CUDA(ptg::deposit_tiled(ndim, shape, P, cfg, false, (flags & PTG_EXACT_ACCUM) != 0, d->sm_count, st, nullptr,
                                        probed));
            const bool wide = ptg::runs_wide_available(ndim, shape);
            for (int k = 0; k < 3; ++k) {
                if (k == 1 && !wide) continue;
                PTG_CUDA(cudaStreamWaitEvent(d->aux[k], probed, 0));
                cudaError_t e = k == 0 ? ptg::deposit_runs(ndim, shape, P, cfg, false, d->sm_count, d->aux[k])
                              : k == 1 ? ptg::deposit_runs(ndim, shape, P, cfg, true, d->sm_count, d->aux[k])
                                       : ptg::deposit_global(ndim, shape, P, shape > 0, cfg, d->sm_count, d->aux[k]);
                cudaEvent_t done = next_event(d);
                cudaError_t e2 = cudaEventRecord(done, d->aux[k]);
                cudaError_t e3 = cudaStreamWaitEvent(st, done, 0);       // always rejoin, also after an error
                PTG_CUDA(e);
                PTG_CUDA(e2);
                PTG_CUDA(e3);
            }
            break;
        }
        case PTG_METHOD_FIXED: {
            if (n <= 0) break;
            ptg::TileConfig *cfg;
            int rc = next_tile_config(d, &cfg);
            if (rc) return rc;
            PTG_CUDA(ptg::deposit_fixed(ndim, shape, P, cfg, d->sm_count, st));
            break;
        }
        case PTG_METHOD_RUNS:
            PTG_CUDA(ptg::deposit_runs(ndim, shape, P, nullptr, false, d->sm_count, st));
            break;
        case PTG_METHOD_RUNS_WIDE:
            if (!ptg::runs_wide_available(ndim, shape))
                return fail(PTG_EUNSUPPORTED, "method RUNS_WIDE has no instance for ndim=%d shape=%d", ndim, shape);
            PTG_CUDA(ptg::deposit_runs(ndim, shape, P, nullptr, true, d->sm_count, st));
            break;
        case PTG_METHOD_SORTED: {
            if (n <= 0) break;
            if (n >= (int64_t(1) << 32)) return fail(PTG_EUNSUPPORTED, "method SORTED handles < 2^32 particles per call");
            const size_t need = ptg::sorted_workspace_bytes(n, ptg_padded_cells(ndim, shape, bins));
            if (workspace && workspace_bytes >= need) {
                PTG_CUDA(ptg::deposit_sorted(ndim, shape, P, workspace, workspace_bytes, d->sm_count, st));
            } else if (workspace) {
                return fail(PTG_ENOMEM, "workspace of %zu bytes given, method SORTED needs %zu", workspace_bytes, need);
            } else {
                // no workspace supplied: stream-ordered allocation, released after the last kernel
                void *ws = nullptr;
                PTG_CUDA(ptg::scratch_alloc(&ws, need, st));
                cudaError_t e = ptg::deposit_sorted(ndim, shape, P, ws, need, d->sm_count, st);
                cudaError_t e2 = cudaFreeAsync(ws, st);
                PTG_CUDA(e);
                PTG_CUDA(e2);
            }
            break;
        }
        default:
            return fail(PTG_EUNSUPPORTED, "deposition method 0x%x is not available in this build", method);
    }
    return PTG_OK;
}

int ensure_staging(DeviceState *d, int64_t cells) {
    if (!d->staged) {
        for (int k = 0; k < kSlots; ++k) {      // (a call that failed half-way is resumed, not repeated)
            if (!d->stream[k]) PTG_CUDA(cudaStreamCreateWithFlags(&d->stream[k], cudaStreamNonBlocking));
            if (!d->slot[k]) PTG_CUDA(cudaMalloc(&d->slot[k], (size_t)4 * kChunk * sizeof(double)));
        }
        d->staged = true;
    }
    if (cells > d->grid_cells) {
        if (d->grid) PTG_CUDA(cudaFree(d->grid));
        d->grid = nullptr;
        d->grid_cells = 0;
        PTG_CUDA(cudaMalloc(&d->grid, (size_t)cells * sizeof(double)));
        d->grid_cells = cells;
    }
    return PTG_OK;
}

// host -> device copy of m elements with an element stride on the host side
cudaError_t h2d_strided(double *dst, const double *src, int64_t stride, int64_t m, cudaStream_t st) {
    if (stride == 1) return cudaMemcpyAsync(dst, src, (size_t)m * 8, cudaMemcpyHostToDevice, st);
    return cudaMemcpy2DAsync(dst, 8, src, (size_t)stride * 8, 8, (size_t)m, cudaMemcpyHostToDevice, st);
}

int ensure_pinned(DeviceState *d) {
    if (d->pinned_ready) return PTG_OK;
    for (int k = 0; k < kSlots; ++k) {
        if (!d->pinned[k])
            PTG_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&d->pinned[k]), (size_t)4 * kStageChunk * sizeof(double),
                                   cudaHostAllocDefault));
        if (!d->h2d_done[k]) PTG_CUDA(cudaEventCreateWithFlags(&d->h2d_done[k], cudaEventDisableTiming));
    }
    d->pinned_ready = true;
    return PTG_OK;
}

// true when cudaMemcpyAsync can DMA straight out of `p` (page-locked or registered host memory)
bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int staging_threads() {
    static const int threads = [] {
        const char *e = getenv("PTG_HOST_THREADS");
        const int hw = (int)std::thread::hardware_concurrency();
        int t = e ? atoi(e) : (hw > 16 ? 16 : hw);   // measured on the 16-thread box: 4/8/16 -> 1.14/1.38/1.73 Gp/s
        return t < 1 ? 1 : (t > 32 ? 32 : t);
    }();
    return threads;
}

// dst[k][i] = src[k][i * stride[k]] for i in [0, m), all arrays k, split over worker threads
void gather_to_pinned(double *const dst[4], const double *const src[4], const int64_t stride[4], int64_t m) {
    const int T = staging_threads();
    auto work = [&](int t) {
        const int64_t b = m * t / T, e = m * (t + 1) / T;
        for (int k = 0; k < 4; ++k) {
            if (!src[k]) continue;
            if (stride[k] == 1) memcpy(dst[k] + b, src[k] + b, (size_t)(e - b) * sizeof(double));
            else {
                const double *s = src[k];
                double *o = dst[k];
                const int64_t st = stride[k];
                for (int64_t i = b; i < e; ++i) o[i] = s[i * st];
            }
        }
    };
    if (T == 1 || m < (int64_t(1) << 16)) { for (int t = 0; t < T; ++t) work(t); return; }
    std::vector<std::thread> pool;
    pool.reserve(T - 1);
    for (int t = 1; t < T; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto &th : pool) th.join();
}

struct DeviceGuard {
    int prev = -1;
    bool active = false;
    int enter(int device) {
        if (device < 0) return PTG_OK;
        if (cudaGetDevice(&prev) != cudaSuccess) return fail(PTG_ENODEVICE, "no CUDA device");
        if (prev != device) {
            if (cudaSetDevice(device) != cudaSuccess) return fail(PTG_ENODEVICE, "cannot select device %d", device);
            active = true;
        }
        return PTG_OK;
    }
    ~DeviceGuard() { if (active) cudaSetDevice(prev); }
};

}  // namespace

extern "C" {

int ptg_abi_version(void) { return PTG_ABI_VERSION; }

const char *ptg_last_error(void) { return g_err; }

const char *ptg_strerror(int code) {
    switch (code) {
        case PTG_OK: return "ok";
        case PTG_EINVAL: return "invalid argument";
        case PTG_ECUDA: return "CUDA error";
        case PTG_ENOMEM: return "out of memory / workspace too small";
        case PTG_ENODEVICE: return "no usable sm_100 GPU";
        case PTG_EUNSUPPORTED: return "unsupported request";
        default: return "unknown error code";
    }
}

int ptg_device_info(int device, char *name, int name_len, int *sm_count, int64_t *l2_bytes,
                    int64_t *hbm_bytes, int *cc_major, int *cc_minor) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        (void)cudaGetLastError();
        return fail(PTG_ENODEVICE, "no CUDA device present");
    }
    if (device < 0) cudaGetDevice(&device);
    cudaDeviceProp p;
    cudaError_t e = cudaGetDeviceProperties(&p, device);
    if (e != cudaSuccess) return fail(PTG_ENODEVICE, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (name && name_len > 0) { strncpy(name, p.name, name_len - 1); name[name_len - 1] = 0; }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (l2_bytes) *l2_bytes = p.l2CacheSize;
    if (hbm_bytes) *hbm_bytes = (int64_t)p.totalGlobalMem;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return PTG_OK;
}

int ptg_axis_setup(double lo_in, double hi_in, int bins, double *lo, double *hi, double *inv) {
    if (!lo || !hi || !inv || bins < 1) return fail(PTG_EINVAL, "ptg_axis_setup: bad arguments");
    double a = lo_in, b = hi_in;
    // np.spacing(b): distance to the next representable number away from zero (pyx:54)
    const double s = std::isfinite(b) ? std::nextafter(b, b < 0 ? -INFINITY : INFINITY) - b : NAN;
    if (std::fabs(b - a) < s) { b += s; a -= s; }         // pyx:55-57
    *lo = a;
    *hi = b;
    *inv = 1.0 / (b - a) * bins;                           // pyx:60, evaluated left to right
    return PTG_OK;
}

int64_t ptg_padded_cells(int ndim, int shape, const int *bins) {
    if (ndim < 1 || ndim > 3 || shape < 0 || shape > 3 || !bins) return 0;
    int64_t c = 1;
    for (int a = 0; a < ndim; ++a) {
        if (bins[a] < 1) return 0;
        const int64_t dim = (int64_t)bins[a] + 2 * shape;
        if (c > (int64_t(1) << 40) / dim) return int64_t(1) << 41;   // saturate, caller rejects
        c *= dim;
    }
    return c;
}

size_t ptg_workspace_bytes(int ndim, int shape, int64_t n, const int *bins, unsigned flags) {
    if ((flags & PTG_METHOD_MASK) != PTG_METHOD_SORTED || n <= 0) return 0;
    const int64_t cells = ptg_padded_cells(ndim, shape, bins);
    if (cells <= 0) return 0;
    return ptg::sorted_workspace_bytes(n, cells);
}

int ptg_deposit_f64(int ndim, int shape, int64_t n, const double *x, const double *y, const double *z,
                    const double *w, const int64_t *strides, const double *lo, const double *inv,
                    const int *bins, double *grid, unsigned flags, void *workspace, size_t workspace_bytes,
                    void *stream) {
    int rc = validate(ndim, shape, n, x, y, z, lo, inv, bins, grid);
    if (rc) return rc;
    DeviceState *d;
    if ((rc = current_device_state(&d))) return rc;
    return deposit_device(d, ndim, shape, n, x, y, z, w, strides, lo, inv, bins, grid, flags,
                          (cudaStream_t)stream, workspace, workspace_bytes);
}

int ptg_deposit_fused_f64(int ndim, int shape, int64_t n, const double *x, const double *y, const double *z,
                          const double *w, const int64_t *strides, const double *lo, const double *inv,
                          const int *bins, double *grid, unsigned flags, int kind, const double *px,
                          const double *py, const double *pz, const int64_t *pstrides, double mass, double c,
                          double elementary_charge, void *stream) {
    int rc = validate(ndim, shape, n, x, y, z, lo, inv, bins, grid);
    if (rc) return rc;
    if (kind < PTG_W_GAMMA_M1 || kind > PTG_W_P) return fail(PTG_EINVAL, "unknown weight expression kind %d", kind);
    if (n > 0 && (!px || !py || !pz)) return fail(PTG_EINVAL, "momentum pointer is NULL");
    DeviceState *d;
    if ((rc = current_device_state(&d))) return rc;
    FuseSpec f{};
    f.kind = kind;
    f.q[0] = px; f.q[1] = py; f.q[2] = pz;
    for (int k = 0; k < 3; ++k) f.qstride[k] = pstrides ? pstrides[k] : 1;
    f.mass = mass; f.c = c; f.qe = elementary_charge;
    return deposit_device(d, ndim, shape, n, x, y, z, w, strides, lo, inv, bins, grid, flags, (cudaStream_t)stream,
                          nullptr, 0, &f);
}

}  // extern "C"

namespace {

// the argument block as this build knows it: callers built against the round-1 header pass a shorter struct
constexpr size_t kArgsV1 = offsetof(ptg_deposit_args, axis_kind);

int read_args(const ptg_deposit_args *in, ptg_deposit_args *a) {
    if (!in || (in->struct_size != kArgsV1 && in->struct_size < sizeof(ptg_deposit_args)))
        return fail(PTG_EINVAL, "bad argument block (struct_size)");
    memset(a, 0, sizeof *a);
    memcpy(a, in, in->struct_size < sizeof *a ? in->struct_size : sizeof *a);
    return PTG_OK;
}

// weight / axis expression set-up shared by ptg_deposit_ex_f64 and ptg_scalar_f64; *fp = nullptr when nothing is fused
int fuse_from_args(const ptg_deposit_args *a, FuseSpec *f, const FuseSpec **fp, int extra_kind) {
    *fp = nullptr;
    unsigned needs = 0;
    bool any = false;
    const int kinds[5] = {a->fuse_kind, a->axis_kind[0], a->axis_kind[1], a->axis_kind[2], extra_kind};
    for (int k = 0; k < 5; ++k) {
        if (kinds[k] == 0) continue;
        if (kinds[k] < 1 || kinds[k] > PTG_S_LAST) return fail(PTG_EINVAL, "unknown scalar expression kind %d", kinds[k]);
        needs |= scalar_needs(kinds[k]);
        any = true;
    }
    if (!any) return PTG_OK;
    if (a->n > 0) {
        if ((needs & 8u) && (!a->px || !a->py || !a->pz)) return fail(PTG_EINVAL, "momentum pointer is NULL");
        if ((needs & 1u) && !a->pos_x) return fail(PTG_EINVAL, "pos_x is NULL but a requested scalar reads x");
        if ((needs & 2u) && !a->pos_y) return fail(PTG_EINVAL, "pos_y is NULL but a requested scalar reads y");
        if ((needs & 4u) && !a->pos_z) return fail(PTG_EINVAL, "pos_z is NULL but a requested scalar reads z");
    }
    f->kind = a->fuse_kind;
    f->q[0] = a->px; f->q[1] = a->py; f->q[2] = a->pz;
    for (int k = 0; k < 3; ++k) f->qstride[k] = a->pstrides ? a->pstrides[k] : 1;
    f->mass = a->mass;
    f->c = a->c != 0.0 ? a->c : 299792458.0;
    f->qe = a->elementary_charge != 0.0 ? a->elementary_charge : 1.602176634e-19;
    for (int k = 0; k < 3; ++k) f->akind[k] = a->axis_kind[k];
    f->pos[0] = a->pos_x; f->pos[1] = a->pos_y; f->pos[2] = a->pos_z;
    for (int k = 0; k < 3; ++k) f->pos_stride[k] = a->pos_strides ? a->pos_strides[k] : 1;
    f->charge = a->charge;
    f->time = a->time;
    if (a->atomic_mass != 0.0) f->amu = a->atomic_mass;
    f->weight_const = a->weight_const != 0.0 ? a->weight_const : 1.0;
    *fp = f;
    return PTG_OK;
}

}  // namespace

extern "C" {

int ptg_deposit_ex_f64(const ptg_deposit_args *in) {
    ptg_deposit_args args;
    int rc = read_args(in, &args);
    if (rc) return rc;
    const ptg_deposit_args *a = &args;
    // a coordinate slot whose axis is an expression needs no array: give validate() something non-NULL
    const double *vx = a->axis_kind[0] ? (const double *)a->grid : a->x;
    const double *vy = a->axis_kind[1] ? (const double *)a->grid : a->y;
    const double *vz = a->axis_kind[2] ? (const double *)a->grid : a->z;
    rc = validate(a->ndim, a->shape, a->n, vx, vy, vz, a->lo, a->inv, a->bins, a->grid);
    if (rc) return rc;
    DeviceState *d;
    if ((rc = current_device_state(&d))) return rc;
    FuseSpec f;
    const FuseSpec *fp = nullptr;
    if ((rc = fuse_from_args(a, &f, &fp, 0))) return rc;
    MaskSpec m{a->mask, a->mask_stride ? a->mask_stride : 1};
    return deposit_device(d, a->ndim, a->shape, a->n, a->x, a->y, a->z, a->w, a->strides, a->lo, a->inv, a->bins,
                          a->grid, a->flags, (cudaStream_t)a->stream, a->workspace, a->workspace_bytes, fp, &m);
}

int ptg_scalar_f64(const ptg_deposit_args *in, int kind, double *out) {
    ptg_deposit_args args;
    int rc = read_args(in, &args);
    if (rc) return rc;
    if (args.n < 0 || (args.n > 0 && !out)) return fail(PTG_EINVAL, "ptg_scalar_f64: bad arguments");
    DeviceState *d;
    if ((rc = current_device_state(&d))) return rc;
    FuseSpec f;
    const FuseSpec *fp = nullptr;
    args.fuse_kind = 0;
    args.axis_kind[0] = args.axis_kind[1] = args.axis_kind[2] = 0;
    if ((rc = fuse_from_args(&args, &f, &fp, kind))) return rc;
    if (!fp) return fail(PTG_EINVAL, "unknown scalar expression kind %d", kind);
    ptg::DepositParams P;
    const double one = 1.0;
    const int bins1[3] = {1, 1, 1};
    fill_params(P, 1, 0, args.n, nullptr, nullptr, nullptr, args.w, args.strides, &one, &one, bins1, nullptr);
    P.f_mass = f.mass;
    P.f_mc2 = (f.mass * f.c) * (f.mass * f.c);
    P.f_c2 = f.c * f.c;
    P.f_qe = f.qe; P.f_c = f.c; P.f_charge = f.charge; P.f_time = f.time; P.f_amu = f.amu; P.f_weight = f.weight_const;
    for (int k = 0; k < 3; ++k) { P.q[k] = f.q[k]; P.qstride[k] = f.qstride[k]; P.pos[k] = f.pos[k]; P.pos_stride[k] = f.pos_stride[k]; }
    PTG_CUDA(ptg::scalar_f64(P, kind, out, d->sm_count, (cudaStream_t)args.stream));
    return PTG_OK;
}

int ptg_minmax_f64(int64_t n, const double *x, int64_t stride, double *out2, void *stream) {
    if (n < 0 || !out2 || (n > 0 && !x)) return fail(PTG_EINVAL, "ptg_minmax_f64: bad arguments");
    DeviceState *d;
    int rc = current_device_state(&d);
    if (rc) return rc;
    PTG_CUDA(ptg::minmax_f64(n, x, stride, out2, d->sm_count, (cudaStream_t)stream));
    return PTG_OK;
}

int ptg_scale_f64(double *grid, int64_t count, double factor, void *stream) {
    if (!grid || count < 0) return fail(PTG_EINVAL, "ptg_scale_f64: bad arguments");
    DeviceState *d;
    int rc = current_device_state(&d);
    if (rc) return rc;
    PTG_CUDA(ptg::scale_f64(grid, count, factor, d->sm_count, (cudaStream_t)stream));
    return PTG_OK;
}

int ptg_divide_f64(double *grid, int64_t count, double divisor, void *stream) {
    if (!grid || count < 0) return fail(PTG_EINVAL, "ptg_divide_f64: bad arguments");
    DeviceState *d;
    int rc = current_device_state(&d);
    if (rc) return rc;
    PTG_CUDA(ptg::divide_f64(grid, count, divisor, d->sm_count, (cudaStream_t)stream));
    return PTG_OK;
}

int ptg_weights_f64(int kind, int64_t n, const double *px, const double *py, const double *pz,
                    const double *weight, const int64_t *strides, double mass, double c, double elementary_charge,
                    double *out, void *stream) {
    if (kind < PTG_W_GAMMA_M1 || kind > PTG_W_P) return fail(PTG_EINVAL, "unknown weight expression kind %d", kind);
    if (n < 0 || (n > 0 && (!px || !py || !pz || !out))) return fail(PTG_EINVAL, "ptg_weights_f64: bad arguments");
    DeviceState *d;
    int rc = current_device_state(&d);
    if (rc) return rc;
    PTG_CUDA(ptg::weights_f64(kind, n, px, py, pz, weight, strides, mass, c, elementary_charge, out, d->sm_count,
                              (cudaStream_t)stream));
    return PTG_OK;
}

int ptg_zero_f64(double *grid, int64_t count, void *stream) {
    if (!grid || count < 0) return fail(PTG_EINVAL, "ptg_zero_f64: bad arguments");
    DeviceState *d;
    int rc = current_device_state(&d);
    if (rc) return rc;
    PTG_CUDA(cudaMemsetAsync(grid, 0, (size_t)count * sizeof(double), (cudaStream_t)stream));
    return PTG_OK;
}

}  // extern "C"

namespace {

// The host-buffer pipeline behind ptg_histogram_host_f64 (result copied back to the host grid `grid`) and
// ptg_deposit_host_f64 (result left in / added to the caller's DEVICE grid `dgrid_user`).
int host_pipeline(int ndim, int shape, int64_t n, const double *x, const double *y,
                  const double *z, const double *w, const int64_t *strides, const double *lo,
                  const double *inv, const int *bins, double *grid, double *dgrid_user, unsigned flags, int device) {
    int rc = validate(ndim, shape, n, x, y, z, lo, inv, bins, dgrid_user ? dgrid_user : grid);
    if (rc) return rc;
    DeviceGuard guard;
    if ((rc = guard.enter(device))) return rc;
    DeviceState *d;
    if ((rc = current_device_state(&d))) return rc;
    std::lock_guard<std::mutex> lk(d->mu);
    const int64_t cells = ptg_padded_cells(ndim, shape, bins);
    if ((rc = ensure_staging(d, dgrid_user ? 1 : cells))) return rc;
    double *const dgrid = dgrid_user ? dgrid_user : d->grid;
    // Sources the copy engine can read directly (page-locked, unit stride) are DMA'd from where they
    // lie.  Anything else -- ordinary pageable numpy memory, strided or reversed views, the columns
    // of an (N, D) array -- is first gathered into a pinned ring by worker threads, chunk k+1 while
    // chunk k is on the bus; cudaMemcpyAsync from pageable memory runs at ~11 GB/s on this box.
    bool direct = true;
    {
        const double *all[4] = {x, ndim >= 2 ? y : nullptr, ndim >= 3 ? z : nullptr, w};
        for (int k = 0; k < 4 && n > 0; ++k) {
            if (!all[k]) continue;
            if ((strides ? strides[k] : 1) != 1 || !is_pinned(all[k])) direct = false;
        }
    }
    if (!direct && (rc = ensure_pinned(d))) return rc;

    if (!dgrid_user || (flags & PTG_ZERO_GRID))
        PTG_CUDA(cudaMemsetAsync(dgrid, 0, (size_t)cells * sizeof(double), d->stream[0]));
    Event zeroed;
    PTG_CUDA(zeroed.create());
    PTG_CUDA(cudaEventRecord(zeroed.e, d->stream[0]));
    for (int k = 1; k < kSlots; ++k) PTG_CUDA(cudaStreamWaitEvent(d->stream[k], zeroed.e, 0));

    const double *src[4] = {x, ndim >= 2 ? y : nullptr, ndim >= 3 ? z : nullptr, w};
    const unsigned dflags = flags & ~PTG_ZERO_GRID;
    // the deterministic method must add the chunks to the grid in a fixed order: one stream
    const bool ordered = (flags & PTG_METHOD_MASK) == PTG_METHOD_SORTED || (flags & PTG_METHOD_MASK) == PTG_METHOD_FIXED;
    int slot = 0;
    const int64_t chunk = direct ? kChunk : kStageChunk;
    bool slot_used[kSlots] = {};
    // bound the work queued ahead of the GPU: chunk c is only enqueued once chunk c - 8 has finished
    constexpr int kInFlight = 8;
    Event fin[kInFlight];
    for (int k = 0; k < kInFlight; ++k) PTG_CUDA(fin[k].create());
    int64_t cidx = 0;
    for (int64_t off = 0; off < n; off += chunk, slot = (slot + 1) % kSlots, ++cidx) {
        const int64_t m = n - off < chunk ? n - off : chunk;
        cudaStream_t st = ordered ? d->stream[0] : d->stream[slot];
        if (cidx >= kInFlight) PTG_CUDA(cudaEventSynchronize(fin[cidx % kInFlight].e));
        double *dptr[4] = {nullptr, nullptr, nullptr, nullptr};
        if (direct) {
            for (int k = 0; k < 4; ++k) {
                if (!src[k]) continue;
                dptr[k] = d->slot[slot] + k * kChunk;
                PTG_CUDA(h2d_strided(dptr[k], src[k] + off, 1, m, st));
            }
        } else {
            if (slot_used[slot]) PTG_CUDA(cudaEventSynchronize(d->h2d_done[slot]));   // pinned slot drained
            double *pdst[4] = {nullptr, nullptr, nullptr, nullptr};
            const double *psrc[4] = {nullptr, nullptr, nullptr, nullptr};
            int64_t pstr[4] = {1, 1, 1, 1};
            for (int k = 0; k < 4; ++k) {
                if (!src[k]) continue;
                pstr[k] = strides ? strides[k] : 1;
                psrc[k] = src[k] + off * pstr[k];
                pdst[k] = d->pinned[slot] + k * kStageChunk;
                dptr[k] = d->slot[slot] + k * kChunk;
            }
            gather_to_pinned(pdst, psrc, pstr, m);
            for (int k = 0; k < 4; ++k)
                if (src[k]) PTG_CUDA(cudaMemcpyAsync(dptr[k], pdst[k], (size_t)m * 8, cudaMemcpyHostToDevice, st));
            PTG_CUDA(cudaEventRecord(d->h2d_done[slot], st));
            slot_used[slot] = true;
        }
        rc = deposit_device(d, ndim, shape, m, dptr[0], dptr[1], dptr[2], dptr[3], nullptr, lo, inv, bins,
                            dgrid, dflags, st);
        if (rc) return rc;
        PTG_CUDA(cudaEventRecord(fin[cidx % kInFlight].e, st));
    }
    // join the ring on stream 0, then bring the padded grid home
    for (int k = 1; k < kSlots; ++k) {
        Event done;
        PTG_CUDA(done.create());
        PTG_CUDA(cudaEventRecord(done.e, d->stream[k]));
        PTG_CUDA(cudaStreamWaitEvent(d->stream[0], done.e, 0));
    }
    if (!dgrid_user)
        PTG_CUDA(cudaMemcpyAsync(grid, dgrid, (size_t)cells * sizeof(double), cudaMemcpyDeviceToHost, d->stream[0]));
    PTG_CUDA(cudaStreamSynchronize(d->stream[0]));
    return PTG_OK;
}

}  // namespace

extern "C" {

int ptg_histogram_host_f64(int ndim, int shape, int64_t n, const double *x, const double *y,
                           const double *z, const double *w, const int64_t *strides, const double *lo,
                           const double *inv, const int *bins, double *grid, unsigned flags, int device) {
    return host_pipeline(ndim, shape, n, x, y, z, w, strides, lo, inv, bins, grid, nullptr, flags, device);
}

int ptg_deposit_host_f64(int ndim, int shape, int64_t n, const double *x, const double *y,
                         const double *z, const double *w, const int64_t *strides, const double *lo,
                         const double *inv, const int *bins, double *grid_device, unsigned flags, int device) {
    if (!grid_device) return fail(PTG_EINVAL, "ptg_deposit_host_f64: grid_device is NULL");
    return host_pipeline(ndim, shape, n, x, y, z, w, strides, lo, inv, bins, nullptr, grid_device, flags, device);
}

int ptg_minmax_host_f64(int64_t n, const double *x, int64_t stride, double *out2, int device) {
    if (n < 0 || !out2 || (n > 0 && !x)) return fail(PTG_EINVAL, "ptg_minmax_host_f64: bad arguments");
    DeviceGuard guard;
    int rc = guard.enter(device);
    if (rc) return rc;
    DeviceState *d;
    if ((rc = current_device_state(&d))) return rc;
    std::lock_guard<std::mutex> lk(d->mu);
    if ((rc = ensure_staging(d, 1))) return rc;
    // page-locked unit-stride memory is DMA'd in place, anything else goes through the pinned ring
    const bool direct = stride == 1 && n > 0 && is_pinned(x);
    if (!direct && (rc = ensure_pinned(d))) return rc;
    // one array only: a chunk may fill a whole staging slot (4 x the per-array share of a deposit chunk)
    const int64_t chunk = direct ? 4 * kChunk : 4 * kStageChunk;
    const int64_t nchunks = n > 0 ? (n + chunk - 1) / chunk : 0;
    if (nchunks == 0) { out2[0] = NAN; out2[1] = NAN; return PTG_OK; }
    // every chunk is reduced on the device into its own (min, max) pair; the pairs come home in ONE copy
    double *partials = nullptr;
    PTG_CUDA(ptg::scratch_alloc(reinterpret_cast<void **>(&partials), (size_t)nchunks * 16, d->stream[0]));
    struct Scratch { double *p; cudaStream_t st; ~Scratch() { cudaFreeAsync(p, st); } } scratch{partials, d->stream[0]};
    PTG_CUDA(cudaStreamSynchronize(d->stream[0]));         // the other ring streams may use it right away
    bool slot_used[kSlots] = {};
    int slot = 0;
    int64_t c = 0;
    for (int64_t off = 0; off < n; off += chunk, slot = (slot + 1) % kSlots, ++c) {
        const int64_t m = n - off < chunk ? n - off : chunk;
        cudaStream_t st = d->stream[slot];
        if (direct) {
            PTG_CUDA(h2d_strided(d->slot[slot], x + off, 1, m, st));
        } else {
            if (slot_used[slot]) PTG_CUDA(cudaEventSynchronize(d->h2d_done[slot]));   // pinned slot drained
            double *pdst[4] = {d->pinned[slot], nullptr, nullptr, nullptr};
            const double *psrc[4] = {x + off * stride, nullptr, nullptr, nullptr};
            const int64_t pstr[4] = {stride, 1, 1, 1};
            gather_to_pinned(pdst, psrc, pstr, m);
            PTG_CUDA(cudaMemcpyAsync(d->slot[slot], pdst[0], (size_t)m * 8, cudaMemcpyHostToDevice, st));
            PTG_CUDA(cudaEventRecord(d->h2d_done[slot], st));
            slot_used[slot] = true;
        }
        PTG_CUDA(ptg::minmax_f64(m, d->slot[slot], 1, partials + 2 * c, d->sm_count, st));
    }
    for (int k = 0; k < kSlots; ++k) PTG_CUDA(cudaStreamSynchronize(d->stream[k]));
    std::vector<double> host((size_t)nchunks * 2);
    PTG_CUDA(cudaMemcpy(host.data(), partials, (size_t)nchunks * 16, cudaMemcpyDeviceToHost));
    double lo = INFINITY, hi = -INFINITY;
    bool nan = false;
    for (int64_t k = 0; k < nchunks; ++k) {
        const double a = host[2 * k], b = host[2 * k + 1];
        if (a != a || b != b) nan = true;
        lo = a < lo ? a : lo;
        hi = b > hi ? b : hi;
    }
    out2[0] = nan ? NAN : lo;
    out2[1] = nan ? NAN : hi;
    return PTG_OK;
}

int ptg_upload_host_f64(int64_t n, const double *src, int64_t stride, double *dst_device, int device) {
    if (n < 0 || (n > 0 && (!src || !dst_device))) return fail(PTG_EINVAL, "ptg_upload_host_f64: bad arguments");
    DeviceGuard guard;
    int rc = guard.enter(device);
    if (rc) return rc;
    DeviceState *d;
    if ((rc = current_device_state(&d))) return rc;
    std::lock_guard<std::mutex> lk(d->mu);
    if ((rc = ensure_staging(d, 1))) return rc;
    const bool direct = stride == 1 && n > 0 && is_pinned(src);
    if (direct) {
        PTG_CUDA(cudaMemcpyAsync(dst_device, src, (size_t)n * 8, cudaMemcpyHostToDevice, d->stream[0]));
        PTG_CUDA(cudaStreamSynchronize(d->stream[0]));
        return PTG_OK;
    }
    if ((rc = ensure_pinned(d))) return rc;
    const int64_t chunk = 4 * kStageChunk;                 // a whole pinned slot per chunk
    bool slot_used[kSlots] = {};
    int slot = 0;
    for (int64_t off = 0; off < n; off += chunk, slot = (slot + 1) % kSlots) {
        const int64_t m = n - off < chunk ? n - off : chunk;
        cudaStream_t st = d->stream[slot];
        if (slot_used[slot]) PTG_CUDA(cudaEventSynchronize(d->h2d_done[slot]));
        double *pdst[4] = {d->pinned[slot], nullptr, nullptr, nullptr};
        const double *psrc[4] = {src + off * stride, nullptr, nullptr, nullptr};
        const int64_t pstr[4] = {stride, 1, 1, 1};
        gather_to_pinned(pdst, psrc, pstr, m);
        PTG_CUDA(cudaMemcpyAsync(dst_device + off, pdst[0], (size_t)m * 8, cudaMemcpyHostToDevice, st));
        PTG_CUDA(cudaEventRecord(d->h2d_done[slot], st));
        slot_used[slot] = true;
    }
    for (int k = 0; k < kSlots; ++k) PTG_CUDA(cudaStreamSynchronize(d->stream[k]));
    return PTG_OK;
}

int ptg_release(void) {
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) { (void)cudaGetLastError(); return PTG_OK; }
    for (int dev = 0; dev < kMaxDevices; ++dev) {
        DeviceState &d = g_dev[dev];
        std::lock_guard<std::mutex> lk(d.mu);
        std::lock_guard<std::mutex> lk_ring(g_ring_mu);
        if (!d.staged && !d.stream[0] && !d.grid && !d.cfg_ring.load() && !d.pool && !d.aux_ready.load()) continue;
        cudaSetDevice(dev);
        cudaDeviceSynchronize();                  // nothing of this device may still read what is freed below
        for (int k = 0; k < kSlots; ++k) {
            if (d.slot[k]) cudaFree(d.slot[k]);
            if (d.stream[k]) cudaStreamDestroy(d.stream[k]);
            if (d.pinned[k]) cudaFreeHost(d.pinned[k]);
            if (d.h2d_done[k]) cudaEventDestroy(d.h2d_done[k]);
            d.pinned[k] = nullptr;
            d.h2d_done[k] = nullptr;
            d.slot[k] = nullptr;
            d.stream[k] = nullptr;
        }
        if (d.grid) cudaFree(d.grid);
        for (int k = 0; k < 3; ++k) { if (d.aux[k]) cudaStreamDestroy(d.aux[k]); d.aux[k] = nullptr; }
        for (int k = 0; k < 64; ++k) { if (d.ev[k]) cudaEventDestroy(d.ev[k]); d.ev[k] = nullptr; }
        d.aux_ready.store(false);
        if (d.pool) (void)cudaMemPoolTrimTo(d.pool, 0);      // hand the cached scratch back to the driver
        if (d.cfg_ring.load()) cudaFree(d.cfg_ring.load());
        d.cfg_ring.store(nullptr);
        d.grid = nullptr; d.grid_cells = 0; d.staged = false; d.pinned_ready = false;
    }
    cudaSetDevice(prev);
    return PTG_OK;
}

int ptg_debug_last_probe(int *mode, int *ext, float *est_hit, float *run_frac, float *near_frac) {
    DeviceState *d;
    int rc = current_device_state(&d);
    if (rc) return rc;
    if (!d->cfg_ring.load() || d->cfg_next.load() == 0) return fail(PTG_EINVAL, "no AUTO/TILED call has been made on this device yet");
    ptg::TileConfig c;
    PTG_CUDA(cudaDeviceSynchronize());
    PTG_CUDA(cudaMemcpy(&c, d->cfg_ring.load() + ((d->cfg_next.load() - 1) % kCfgRing), sizeof c, cudaMemcpyDeviceToHost));
    if (mode) *mode = c.mode;
    if (ext) for (int a = 0; a < 3; ++a) ext[a] = c.ext[a];
    if (est_hit) *est_hit = c.est_hit;
    if (run_frac) *run_frac = c.run_frac;
    if (near_frac) *near_frac = c.near_frac;
    return PTG_OK;
}

int ptg_selftest_radix_sort(int64_t n, int bits, unsigned *keys, unsigned *perm, void *stream) {
    if (n < 0 || bits < 1 || bits > 32 || n >= (int64_t(1) << 32) || (n > 0 && (!keys || !perm)))
        return fail(PTG_EINVAL, "ptg_selftest_radix_sort: bad arguments");
    DeviceState *d;
    int rc = current_device_state(&d);
    if (rc) return rc;
    PTG_CUDA(ptg::selftest_radix_sort(n, bits, keys, perm, (cudaStream_t)stream));
    return PTG_OK;
}

int ptg_selftest_div6(int64_t n, const double *in, double *out, void *stream) {
    if (n < 0 || (n > 0 && (!in || !out))) return fail(PTG_EINVAL, "ptg_selftest_div6: bad arguments");
    DeviceState *d;
    int rc = current_device_state(&d);
    if (rc) return rc;
    PTG_CUDA(ptg::selftest_div6(n, in, out, (cudaStream_t)stream));
    return PTG_OK;
}

}  // extern "C"
