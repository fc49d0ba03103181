// ptg_deposit_global.cu -- deposit kernels that accumulate straight into the padded
// grid in global memory with native fp64 reductions (RED.E.ADD.F64, resolved by the
// L2 atomic units; the 8-137 MB grids of the target configurations are L2-resident
// or nearly so on B200's 126 MB L2).
//
// Replaces the loops of /root/reference/postpic/particles/_particlestogrid.pyx
// :70-129 (1-D), :181-276 (2-D), :341-463 (3-D).
//
//   method GLOBAL : one thread = one particle, one RED per stencil cell.  A warp-level
//                   RED instruction then touches 32 unrelated 32-byte sectors.
//   method ROWS   : the warp first evaluates 32 stencils, parks the (S+1)^(ndim-1)
//                   rows of each in shared memory, then re-reads them so that the
//                   lanes of one RED instruction cover whole stencil rows: S+1 adjacent
//                   lanes hit S+1 adjacent doubles (one or two sectors) and the LSU
//                   coalesces them into one L2 request per sector.
#include <cstdlib>

#include "ptg_internal.cuh"

namespace ptg {

#ifndef PTG_ROWS_THREADS
#define PTG_ROWS_THREADS 256
#endif
#ifndef PTG_GLOBAL_UNROLL
#define PTG_GLOBAL_UNROLL 4
#endif
constexpr int kThreads = PTG_ROWS_THREADS;
constexpr int kUnroll = PTG_GLOBAL_UNROLL;      // independent particle loads in flight per thread

// ---------------------------------------------------------------------------------
// one particle -> REDs (method GLOBAL)
// ---------------------------------------------------------------------------------
template <int ND, int S, bool W>
__device__ __forceinline__ void deposit_particle_global(const DepositParams &P, double x, double y,
                                                        double z, double w) {
    constexpr int NW = StencilWidth<S>::value;
    constexpr int NC = ND >= 2 ? NW : 1, NDD = ND >= 3 ? NW : 1;
    int b[3];
    double wx[NW], wy[NW], wz[NW];
    if (!particle_stencil<ND, S>(P, x, y, z, b, wx, wy, wz) || b[0] < P.slab_lo || b[0] >= P.slab_hi) return;
    const int64_t g1 = ND >= 2 ? P.dims[1] : 1, g2 = ND >= 3 ? P.dims[2] : 1;
    double *g = P.grid + ((int64_t)b[0] * g1 + b[1]) * g2 + b[2];
#pragma unroll
    for (int a = 0; a < NW; ++a)
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int d = 0; d < NDD; ++d)
                red_add(g + (a * g1 + c) * g2 + d, stencil_value<ND, S, W>(wx, wy, wz, a, c, d, w));
}

template <int ND, int S, bool W, bool F>
__global__ void __launch_bounds__(kThreads) k_deposit_global(const DepositParams P, const TileConfig *gate) {
    if (gate != nullptr && gate->mode != kModeRows) return;
    const int64_t nthreads = (int64_t)gridDim.x * kThreads;
    int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const double *px = P.p[0], *py = P.p[1], *pz = P.p[2];
    const int64_t sx = P.stride[0], sy = P.stride[1], sz = P.stride[2];
    // main loop: kUnroll particles per thread with all loads issued before any use
    for (; i + (kUnroll - 1) * nthreads < P.n; i += kUnroll * nthreads) {
        double x[kUnroll], y[kUnroll], z[kUnroll], w[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t k = i + u * nthreads;
            x[u] = particle_coord<F>(P, 0, k);
            y[u] = ND >= 2 ? particle_coord<F>(P, 1, k) : 0.0;
            z[u] = ND >= 3 ? particle_coord<F>(P, 2, k) : 0.0;
            w[u] = W ? particle_weight<F>(P, k) : 1.0;
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (particle_selected(P, i + u * nthreads)) deposit_particle_global<ND, S, W>(P, x[u], y[u], z[u], w[u]);
    }
    for (; i < P.n; i += nthreads) {
        const double x = particle_coord<F>(P, 0, i);
        const double y = ND >= 2 ? particle_coord<F>(P, 1, i) : 0.0;
        const double z = ND >= 3 ? particle_coord<F>(P, 2, i) : 0.0;
        const double w = W ? particle_weight<F>(P, i) : 1.0;
        if (particle_selected(P, i)) deposit_particle_global<ND, S, W>(P, x, y, z, w);
    }
}

// ---------------------------------------------------------------------------------
// method ROWS: sector-coalesced reductions
// ---------------------------------------------------------------------------------
// Lanes per row group (next power of two >= NW) and the layout of the per-warp staging area:
// value k of the particle held by lane p lives at stage[k * KS + p].  Writes (fixed k, p = lane)
// are contiguous, hence conflict-free.  The transposed read has lane (g, j) fetch value
// row*NW + j of particle pg*PPI + g; with KS == 4 (mod 16) for G = 4 -- resp. 8 (mod 16) for
// G = 2 -- the 16 lanes of a 64-bit phase fall on 16 different bank pairs: (4j + g) resp. (8j + g).
template <int ND, int S> struct RowsCfg {
    static constexpr int NW = S + 1;
    static constexpr int G = NW == 2 ? 2 : 4;
    static constexpr int ROWS = ND == 1 ? 1 : NW;          // rows per slab
    static constexpr int SLABS = ND == 3 ? NW : 1;
    static constexpr int K = ROWS * NW;                     // doubles staged per particle per slab
    static constexpr int KS = NW == 2 ? 40 : 36;            // stride between consecutive values k
    static constexpr int PPI = 32 / G;                      // particles per RED instruction
    static constexpr int ITERS = ROWS * G;
};

template <int ND, int S, bool W, bool STAG, bool F, bool KEEP = false>
__global__ void __launch_bounds__(kThreads) k_deposit_rows(const DepositParams P, const TileConfig *gate) {
    // KEEP (the slab passes of a grid larger than L2): mark the accumulator lines evict-last in L2 (red_add_hint,
    // ptg_stencil.cuh).  A compile-time switch: as a run-time one it cost the L2-resident headline case 6 per cent.
    constexpr bool keep = KEEP;
    const unsigned long long keep_pol = KEEP ? l2_keep_policy() : 0ull;
    if (gate != nullptr && gate->mode != kModeRows) return;
    using C = RowsCfg<ND, S>;
    constexpr int NW = C::NW;
    __shared__ __align__(16) double stage_all[(kThreads / 32) * C::K * C::KS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *stage = stage_all + warp * C::K * C::KS;
    const int g = lane / C::G, j = lane % C::G;

    const double *px = P.p[0], *py = P.p[1], *pz = P.p[2];
    const int64_t sx = P.stride[0], sy = P.stride[1], sz = P.stride[2];
    // geometry of the array the reductions go to: the padded grid itself, or the staggered staging grid
    const int pitch = STAG ? P.stag_pitch : P.dims[ND - 1];
    const int row_stride = ND == 1 ? 0 : pitch;                      // between rows of a slab
    const int slab_stride = ND == 3 ? P.dims[1] * pitch : 0;
    double *const target = STAG ? P.stag : P.grid;

    // Every CTA owns a contiguous span of the particle arrays: on ordered input the CTAs then work
    // on different parts of the grid at any moment instead of all hammering the same few rows.
    // (warp-uniform trip count so that the warp-level exchanges stay convergent)
    const int64_t span = ((P.n + gridDim.x - 1) / gridDim.x + kThreads - 1) / kThreads * kThreads;
    const int64_t begin = (int64_t)blockIdx.x * span;
    const int64_t end = begin + span < P.n ? begin + span : P.n;
    for (int64_t i0 = begin + warp * 32; i0 < end; i0 += kThreads) {
        const int64_t i = i0 + lane;
        double wx[NW], wy[NW], wz[NW];
        int bx = 0, by = 0, bz = 0;
        double w = 1.0;
        bool ok = i < end && particle_selected(P, i);
        if (ok) {
            const double x = particle_coord<F>(P, 0, i);
            const double y = ND >= 2 ? particle_coord<F>(P, 1, i) : 0.0;
            const double z = ND >= 3 ? particle_coord<F>(P, 2, i) : 0.0;
            if (W) w = particle_weight<F>(P, i);
            int b[3];
            ok = particle_stencil<ND, S>(P, x, y, z, b, wx, wy, wz) && b[0] >= P.slab_lo && b[0] < P.slab_hi;
            bx = b[0]; by = b[1]; bz = b[2];
        }
        int lin0 = -1;                                    // padded linear index of the first cell
        if (ok) {
            const int col = ND == 1 ? bx : ND == 2 ? by : bz;          // first cell of the rows on the last axis
            const int rowlin = ND == 1 ? 0 : ND == 2 ? bx : bx * P.dims[1] + by;
            lin0 = rowlin * pitch + col;
            if (STAG) {
                // rows of 3 straddle a sector when col mod 4 is 2 or 3, rows of 2 when it is 3
                const bool second = NW == 3 ? (col & 2) != 0 : (col & 3) == 3;
                if (second) lin0 += (int)P.stag_copy + 2;
            }
        }

#pragma unroll
        for (int slab = 0; slab < C::SLABS; ++slab) {
            if (ok) {
                double *mine = stage + lane;
                if (ND == 1) {
#pragma unroll
                    for (int a = 0; a < NW; ++a) mine[a * C::KS] = W ? __dmul_rn(wx[a], w) : wx[a];
                } else if (ND == 2) {
#pragma unroll
                    for (int a = 0; a < NW; ++a)
#pragma unroll
                        for (int b = 0; b < NW; ++b) {
                            const double c = __dmul_rn(wx[a], wy[b]);
                            mine[(a * NW + b) * C::KS] = W ? __dmul_rn(c, w) : c;
                        }
                } else {
#pragma unroll
                    for (int a = 0; a < NW; ++a) {
                        const double cxy = __dmul_rn(wx[slab], wy[a]);
#pragma unroll
                        for (int b = 0; b < NW; ++b) {
                            const double c = __dmul_rn(cxy, wz[b]);
                            mine[(a * NW + b) * C::KS] = W ? __dmul_rn(c, w) : c;
                        }
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int it = 0; it < C::ITERS; ++it) {
                const int row = it / C::G, pg = it % C::G;
                const int p = pg * C::PPI + g;
                const int l0 = __shfl_sync(0xffffffffu, lin0, p);
                if (j < NW && l0 >= 0) {
                    const double v = stage[(row * NW + j) * C::KS + p];
                    red_add_hint(target + (int64_t)l0 + (int64_t)slab * slab_stride + row * row_stride + j, v, keep, keep_pol);
                }
            }
            __syncwarp();
        }
    }
}

// The same sector-coalesced reductions without the shared-memory transposition: G adjacent lanes
// share ONE particle -- each of them loads it (a broadcast within the request), evaluates its stencil
// and keeps only the last-axis weight of its own cell j -- so the lanes of a RED instruction cover
// whole stencil rows by construction.  The arithmetic is done G times over, on the FP64 pipe that
// this path leaves idle; what it saves is the l1tex bandwidth of the staging stores and loads,
// which share a pipeline with the reductions (ncu: 54 shared-memory wavefronts per 32 particles next
// to 96 sector-cycles of REDs).  No warp-level exchange at all, hence no convergence constraints.
constexpr int kQuadUnroll = 2;
template <int ND, int S, bool W, bool STAG, bool F, bool KEEP = false>
__global__ void __launch_bounds__(kThreads) k_deposit_quads(const DepositParams P, const TileConfig *gate) {
    // KEEP (the slab passes of a grid larger than L2): mark the accumulator lines evict-last in L2 (red_add_hint,
    // ptg_stencil.cuh).  A compile-time switch: as a run-time one it cost the L2-resident headline case 6 per cent.
    constexpr bool keep = KEEP;
    const unsigned long long keep_pol = KEEP ? l2_keep_policy() : 0ull;
    if (gate != nullptr && gate->mode != kModeRows) return;
    using C = RowsCfg<ND, S>;
    constexpr int NW = C::NW, G = C::G, U = kQuadUnroll;
    constexpr int kGroups = kThreads / G;                             // particles per CTA per step
    const int j = threadIdx.x % G, group = threadIdx.x / G;

    const double *px = P.p[0], *py = P.p[1], *pz = P.p[2];
    const int64_t sx = P.stride[0], sy = P.stride[1], sz = P.stride[2];
    const int pitch = STAG ? P.stag_pitch : P.dims[ND - 1];
    const int64_t row_stride = ND == 1 ? 0 : pitch;
    const int64_t slab_stride = ND == 3 ? (int64_t)P.dims[1] * pitch : 0;
    double *const target = STAG ? P.stag : P.grid;

    const int64_t span = ((P.n + gridDim.x - 1) / gridDim.x + kGroups - 1) / kGroups * kGroups;
    const int64_t begin = (int64_t)blockIdx.x * span;
    const int64_t end = begin + span < P.n ? begin + span : P.n;
    for (int64_t i0 = begin + group; i0 < end; i0 += (int64_t)kGroups * U) {
        double x[U], y[U], z[U], w[U];
        bool live[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + (int64_t)u * kGroups;
            live[u] = i < end && particle_selected(P, i);
            x[u] = y[u] = z[u] = 0.0;
            w[u] = 1.0;
            if (live[u]) {
                x[u] = particle_coord<F>(P, 0, i);
                if (ND >= 2) y[u] = particle_coord<F>(P, 1, i);
                if (ND >= 3) z[u] = particle_coord<F>(P, 2, i);
                if (W) w[u] = particle_weight<F>(P, i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            double wx[NW], wy[NW], wz[NW];
            int b[3];
            const bool ok = live[u] && particle_stencil<ND, S>(P, x[u], y[u], z[u], b, wx, wy, wz) &&
                            b[0] >= P.slab_lo && b[0] < P.slab_hi;
            if (!ok || j >= NW) continue;
            const int col = ND == 1 ? b[0] : ND == 2 ? b[1] : b[2];
            const int rowlin = ND == 1 ? 0 : ND == 2 ? b[0] : b[0] * P.dims[1] + b[1];
            int64_t lin = (int64_t)rowlin * pitch + col + j;
            if (STAG) {
                const bool second = NW == 3 ? (col & 2) != 0 : (col & 3) == 3;
                if (second) lin += P.stag_copy + 2;
            }
            // the last-axis weight of this lane's cell
            const double (&wl)[NW] = ND == 1 ? wx : ND == 2 ? wy : wz;
            double wj = wl[0];
#pragma unroll
            for (int k = 1; k < NW; ++k) wj = j == k ? wl[k] : wj;
            if (ND == 1) {
                red_add_hint(target + lin, W ? __dmul_rn(wj, w[u]) : wj, keep, keep_pol);
            } else if (ND == 2) {
#pragma unroll
                for (int a = 0; a < NW; ++a) {
                    const double c = __dmul_rn(wx[a], wj);
                    red_add_hint(target + lin + a * row_stride, W ? __dmul_rn(c, w[u]) : c, keep, keep_pol);
                }
            } else {
#pragma unroll
                for (int a = 0; a < NW; ++a)
#pragma unroll
                    for (int bb = 0; bb < NW; ++bb) {
                        const double c = __dmul_rn(__dmul_rn(wx[a], wy[bb]), wj);
                        red_add_hint(target + lin + a * slab_stride + bb * row_stride, W ? __dmul_rn(c, w[u]) : c, keep, keep_pol);
                    }
            }
        }
    }
}

// grid[r][j] += stag0[r][j] + stag1[r][j + 2]   (rows = product of all but the last axis)
__global__ void __launch_bounds__(256) k_merge_staggered(const DepositParams P, int64_t rows, int last,
                                                          const TileConfig *gate) {
    if (gate != nullptr && gate->mode != kModeRows) return;
    const int64_t total = rows * last, step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
        const int64_t r = i / last;
        const int j = (int)(i - r * last);
        const double a = P.stag[r * P.stag_pitch + j], b = P.stag[P.stag_copy + r * P.stag_pitch + j + 2];
        // RED, not a plain store: other streams (the host pipeline's chunks) may be merging into the same grid
        if (a != 0.0 || b != 0.0) red_add(P.grid + i, __dadd_rn(a, b));
    }
}

__global__ void __launch_bounds__(256) k_zero_gated(double4 *p, int64_t n32, const TileConfig *gate) {
    if (gate->mode != kModeRows) return;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += step) p[i] = make_double4(0.0, 0.0, 0.0, 0.0);
}

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
template <int ND, int S, bool W>
static cudaError_t launch_global(const DepositParams &P, const TileConfig *gate, int sm_count, cudaStream_t st) {
    if (P.n <= 0) return cudaSuccess;
    int64_t want = (P.n + (int64_t)kThreads * kUnroll - 1) / ((int64_t)kThreads * kUnroll);
    int64_t cap = (int64_t)sm_count * 8;                  // 8 x 256 threads = full occupancy
    int blocks = (int)(want < cap ? want : cap);
    if (blocks < 1) blocks = 1;
    if (W && fused_mode(P)) k_deposit_global<ND, S, W, W><<<blocks, kThreads, 0, st>>>(P, gate);
    else k_deposit_global<ND, S, W, false><<<blocks, kThreads, 0, st>>>(P, gate);
    return cudaGetLastError();
}

template <int ND, int S, bool W>
static cudaError_t launch_rows(const DepositParams &P, const TileConfig *gate, int sm_count, cudaStream_t st) {
    if (P.n <= 0) return cudaSuccess;
    int64_t want = (P.n + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)sm_count * 8;
    int blocks = (int)(want < cap ? want : cap);
    if (blocks < 1) blocks = 1;
    // lane-group kernel for rows of 2 or 3 cells in 2-D / 3-D (measured against the staged kernel on cfg2-like
    // input: CIC 75.5 vs 68.1, TSC 50.5 vs 49.6, 3-D CIC 23.3 vs 20.9 Gp/s; spline3 28.0 vs 29.6, so rows
    // of 4 stay staged, and so do the fused-weight variants whose sqrt/divide chain would run G times over).
    // PTG_ROWS_IMPL=q|s forces one of them (A/B runs).
    static const int forced = [] { const char *e = getenv("PTG_ROWS_IMPL"); return !e ? 0 : e[0] == 'q' ? 1 : 2; }();
    const bool quads = forced ? forced == 1 : (ND >= 2 && (S == 1 || S == 2) && !(W && fused_mode(P)));
    if (P.stag != nullptr) {
        if constexpr (S == 1 || S == 2) {
            if (quads) {
                if (W && fused_mode(P)) k_deposit_quads<ND, S, W, true, W><<<blocks, kThreads, 0, st>>>(P, gate);
                else k_deposit_quads<ND, S, W, true, false><<<blocks, kThreads, 0, st>>>(P, gate);
            } else if (W && fused_mode(P)) k_deposit_rows<ND, S, W, true, W><<<blocks, kThreads, 0, st>>>(P, gate);
            else k_deposit_rows<ND, S, W, true, false><<<blocks, kThreads, 0, st>>>(P, gate);
            const int last = P.dims[ND - 1];
            const int64_t rows = (int64_t)P.dims[0] * P.dims[1] * P.dims[2] / last;
            int64_t mb = (rows * last + 255) / 256;
            k_merge_staggered<<<(int)(mb < cap ? mb : cap), 256, 0, st>>>(P, rows, last, gate);
            return cudaGetLastError();
        }
    }
    if (P.slab_lo > -(1 << 30) || P.slab_hi < (1 << 30)) {        // slab pass: the KEEP instances
        if (quads) {
            if (W && fused_mode(P)) k_deposit_quads<ND, S, W, false, W, true><<<blocks, kThreads, 0, st>>>(P, gate);
            else k_deposit_quads<ND, S, W, false, false, true><<<blocks, kThreads, 0, st>>>(P, gate);
        } else if (W && fused_mode(P)) k_deposit_rows<ND, S, W, false, W, true><<<blocks, kThreads, 0, st>>>(P, gate);
        else k_deposit_rows<ND, S, W, false, false, true><<<blocks, kThreads, 0, st>>>(P, gate);
        return cudaGetLastError();
    }
    if (quads) {
        if (W && fused_mode(P)) k_deposit_quads<ND, S, W, false, W><<<blocks, kThreads, 0, st>>>(P, gate);
        else k_deposit_quads<ND, S, W, false, false><<<blocks, kThreads, 0, st>>>(P, gate);
    } else if (W && fused_mode(P)) k_deposit_rows<ND, S, W, false, W><<<blocks, kThreads, 0, st>>>(P, gate);
    else k_deposit_rows<ND, S, W, false, false><<<blocks, kThreads, 0, st>>>(P, gate);
    return cudaGetLastError();
}

template <int ND, int S>
static cudaError_t dispatch_w(const DepositParams &P, bool rows, const TileConfig *gate, int sm_count, cudaStream_t st) {
    const bool weighted = P.p[3] != nullptr || fused_mode(P);
    if (rows) {
        if constexpr (S >= 1) {
            return weighted ? launch_rows<ND, S, true>(P, gate, sm_count, st) : launch_rows<ND, S, false>(P, gate, sm_count, st);
        }
    }
    return weighted ? launch_global<ND, S, true>(P, gate, sm_count, st) : launch_global<ND, S, false>(P, gate, sm_count, st);
}

template <int ND>
static cudaError_t dispatch_s(int shape, const DepositParams &P, bool rows, const TileConfig *gate, int sm_count, cudaStream_t st) {
    switch (shape) {
        case 0: return dispatch_w<ND, 0>(P, false, gate, sm_count, st);
        case 1: return dispatch_w<ND, 1>(P, rows, gate, sm_count, st);
        case 2: return dispatch_w<ND, 2>(P, rows, gate, sm_count, st);
        default: return dispatch_w<ND, 3>(P, rows, gate, sm_count, st);
    }
}

static cudaError_t deposit_global_pass(int ndim, int shape, const DepositParams &P, bool rows, const TileConfig *gate,
                                       int sm_count, cudaStream_t st) {
    switch (ndim) {
        case 1: return dispatch_s<1>(shape, P, rows, gate, sm_count, st);
        case 2: return dispatch_s<2>(shape, P, rows, gate, sm_count, st);
        default: return dispatch_s<3>(shape, P, rows, gate, sm_count, st);
    }
}

// Grids that do not fit in L2 (e.g. 258^3 doubles = 137 MB vs 126 MB) make most reductions miss:
// measured 58 G sector-REDs/s against 190 G/s on an L2-resident grid.  Such grids are deposited
// in slab passes: every pass streams all particles again but only deposits those whose base
// cell on axis 0 falls in the pass's slab, so the reductions of a pass hit an L2-resident
// slice of <= 72 MiB (sweep in profiles/README.md: 24/36/48/72/100/200 MiB -> 9.7/14.4/18.9/21.8/
// 21.8/11.1 Gp/s for 3-D CIC on 258^3).  Re-reading 32 B/particle per pass is cheaper than the misses.
static int64_t slab_bytes() {
    static int64_t v = 0;
    if (v == 0) {
        const char *e = getenv("PTG_SLAB_MB");            // tuning knob; default 72 MiB (measured best for 258^3)
        int64_t mb = e ? atoll(e) : 72;
        if (mb < 1) mb = 1;
        v = mb << 20;
    }
    return v;
}

cudaError_t deposit_global(int ndim, int shape, const DepositParams &P0, bool rows, const TileConfig *gate,
                           int sm_count, cudaStream_t st) {
    DepositParams P = P0;
    const int64_t row_bytes = (int64_t)P.dims[1] * P.dims[2] * 8;          // one axis-0 slice
    const int64_t grid_bytes = row_bytes * P.dims[0];
    int passes = 1;
    const int64_t kSlabBytes = slab_bytes();
    if (grid_bytes > 96 * (int64_t(1) << 20) && P.n >= (int64_t(1) << 22)) passes = (int)((grid_bytes + kSlabBytes - 1) / kSlabBytes);
    if (passes > P.dims[0]) passes = P.dims[0];
    if (passes <= 1) {
        P.slab_lo = -(1 << 30);
        P.slab_hi = 1 << 30;
        // staggered staging grid: stencil rows of 2 or 3 cells, both copies L2-resident, enough
        // particles to pay for clearing and merging it
        static const bool no_stagger = getenv("PTG_NO_STAGGER") != nullptr;
        const int last = P.dims[ndim - 1];
        const int pitch = (last + 2 + 3) / 4 * 4;
        const int64_t nrows = (int64_t)P.dims[0] * P.dims[1] * P.dims[2] / last;
        const int64_t copy = nrows * pitch + 4;
        const size_t bytes = (size_t)(2 * copy) * sizeof(double);
        if (rows && !no_stagger && (shape == 1 || shape == 2) && bytes <= ((size_t)64 << 20) && 2 * copy < (int64_t(1) << 31) &&
            P.n >= (int64_t(1) << 20)) {
            void *buf = nullptr;
            cudaError_t e = scratch_alloc(&buf, bytes, st);           // stream-ordered, pooled: no sync
            if (e != cudaSuccess) return e;
            if (gate != nullptr) {
                // AUTO: clear the staging grid only if method ROWS is the one that runs (a 17 MB memset on the
                // critical path of every AUTO call otherwise, and 17 MB of zeros pushed through the L2)
                k_zero_gated<<<sm_count * 4, 256, 0, st>>>(static_cast<double4 *>(buf), (int64_t)(bytes / 32), gate);
                e = cudaGetLastError();
            } else {
                e = cudaMemsetAsync(buf, 0, bytes, st);
            }
            if (e != cudaSuccess) return e;
            P.stag = static_cast<double *>(buf);
            P.stag_pitch = pitch;
            P.stag_copy = copy / 4 * 4;
            e = deposit_global_pass(ndim, shape, P, rows, gate, sm_count, st);
            cudaError_t e2 = cudaFreeAsync(buf, st);
            return e != cudaSuccess ? e : e2;
        }
        return deposit_global_pass(ndim, shape, P, rows, gate, sm_count, st);
    }
    for (int s = 0; s < passes; ++s) {
        P.slab_lo = s == 0 ? -(1 << 30) : (int)((int64_t)P.dims[0] * s / passes);
        P.slab_hi = s == passes - 1 ? (1 << 30) : (int)((int64_t)P.dims[0] * (s + 1) / passes);
        cudaError_t e = deposit_global_pass(ndim, shape, P, rows, gate, sm_count, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace ptg
