// ptg_deposit_runs.cu -- method RUNS: run-length accumulation in registers.
//
// For cell-ordered input (PIC codes that sort their particles by cell write their dumps
// that way; 1e8..1e9 particles on 1e6 cells means ~100..1000 consecutive particles per
// cell) almost every particle lands on the same stencil as its predecessor.  Each thread
// therefore walks K CONSECUTIVE particles, keeps the stencil sums of the current run in
// registers and only commits when the base cell changes; the runs still open at the end
// of the K particles are merged across the warp with a segmented shuffle reduction, and
// one lane per run issues the (S+1)^ndim RED.F64.  Scatter traffic drops from
// (S+1)^ndim reductions per particle to (S+1)^ndim per run, so the kernel is bound by
// the streaming read of the particle arrays -- the HBM roofline -- instead of by the
// L2 atomic units.  Input without runs is still deposited correctly, only slowly (one
// uncoalesced RED per stencil cell per particle); AUTO selects this method only when
// the probe kernel saw runs.
//
// Particle arrays are read with 256-bit loads when contiguous and 32-byte aligned (a lane's four
// consecutive doubles are one sector; a warp reads 1 KB contiguous per array and batch).
//
// Replaces the loops of /root/reference/postpic/particles/_particlestogrid.pyx
// :70-129 (1-D), :181-276 (2-D), :341-463 (3-D).
#include "ptg_internal.cuh"

namespace ptg {

#ifndef PTG_RUNS_MERGE_MAX_HEADS
#define PTG_RUNS_MERGE_MAX_HEADS 20      // warp merge only when the 32 lanes hold at most this many runs
#endif
#ifndef PTG_RUNS_THREADS
#define PTG_RUNS_THREADS 256
#endif
#ifndef PTG_RUNS_BATCHES
#define PTG_RUNS_BATCHES 2
#endif
constexpr int kRT = PTG_RUNS_THREADS;      // threads per CTA
constexpr int kRK = 4;        // consecutive particles per load batch
constexpr int kRB = PTG_RUNS_BATCHES;        // batches per thread between two warp-level merges (8 consecutive particles;
                              // 16 measured slower: a lane's 128-byte span is then fetched in four visits)

// 256-bit load (sm_100: LDG.E.NA.EFL2.256): the four consecutive doubles of one lane are exactly one
// 32-byte sector, fetched once, not allocated in L1 and marked evict-first in L2.
__device__ __forceinline__ void ld_stream_v4(const double *p, double (&v)[4]) {
    unsigned long long a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0, %1, %2, %3}, [%4];"
                 : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    v[0] = __longlong_as_double((long long)a);
    v[1] = __longlong_as_double((long long)b);
    v[2] = __longlong_as_double((long long)c);
    v[3] = __longlong_as_double((long long)d);
}

// K consecutive values starting at element i (all K valid)
template <bool VEC>
__device__ __forceinline__ void load_k(const double *base, int64_t stride, int64_t i, double (&v)[kRK]) {
    if (VEC) {
        static_assert(kRK == 4, "one 256-bit load per batch");
        ld_stream_v4(base + i, v);
    } else {
#pragma unroll
        for (int j = 0; j < kRK; ++j) v[j] = __ldg(base + (i + j) * stride);   // L1 keeps the shared sectors
    }
}

// resident CTAs per SM asked of ptxas: 4 (64 registers) for register windows of up to 9 cells
template <int ND, int S, int WD> struct RunsMinBlocks {
    static constexpr int MW = StencilWidth<S>::value + WD;
    static constexpr int MV = ND == 1 ? MW : ND == 2 ? MW * MW : MW * MW * MW;
    static constexpr int value = MV <= 9 ? 4 : MV <= 16 ? 3 : 2;
};

template <int ND, int S, bool W, bool VEC, int WD, bool F>
__global__ void __launch_bounds__(kRT, RunsMinBlocks<ND, S, WD>::value) k_deposit_runs(const DepositParams P, const TileConfig *gate) {
    if (gate != nullptr && gate->mode != (WD ? kModeRunsWide : kModeRuns)) return;
    constexpr int NW = StencilWidth<S>::value;
    // register window of the open run: the stencil of its first particle, widened by WD cells per
    // axis so that neighbours whose base cell is one further along any axis still fit
    constexpr int MW = NW + WD;
    constexpr int MA = MW, MC = ND >= 2 ? MW : 1, MD = ND >= 3 ? MW : 1;
    constexpr int MV = MA * MC * MD;
    const int lane = threadIdx.x & 31;
    const int64_t g1 = ND >= 2 ? P.dims[1] : 1, g2 = ND >= 3 ? P.dims[2] : 1;

    auto commit = [&](int key, const double (&acc)[MV]) {
        // key = padded linear index of the window's first cell.  Exact zeros are skipped: they
        // carry nothing, and the outermost cells of a widened window may lie beyond the grid
        // (they can only ever hold zero there).
        double *g = P.grid + key;
#pragma unroll
        for (int a = 0; a < MA; ++a)
#pragma unroll
            for (int c = 0; c < MC; ++c)
#pragma unroll
                for (int d = 0; d < MD; ++d) {
                    const double v = acc[(a * MC + c) * MD + d];
                    if (!WD || v != 0.0) red_add(g + (a * g1 + c) * g2 + d, v);
                }
    };

    // contiguous span per CTA so that concurrently running CTAs work on different grid regions
    const int64_t per_cta = ((P.n + gridDim.x - 1) / gridDim.x + kRT * kRK * kRB - 1) / (kRT * kRK * kRB) * (kRT * kRK * kRB);
    const int64_t begin = (int64_t)blockIdx.x * per_cta;
    const int64_t end = begin + per_cta < P.n ? begin + per_cta : P.n;

    constexpr int PER = kRK * kRB;         // consecutive particles per thread between two warp merges
    for (int64_t c0 = begin; c0 < end; c0 += (int64_t)kRT * PER) {
        const int64_t t0 = c0 + (int64_t)threadIdx.x * PER;
        int cur = -1;                 // first cell of the open run's window (linear), -1 = none
        int ab[3] = {0, 0, 0};        // the same per axis
        double acc[MV];
#pragma unroll
        for (int k = 0; k < MV; ++k) acc[k] = 0.0;

        double x[kRK], y[kRK], z[kRK], w[kRK];
        auto load_batch = [&](int64_t i0) {
            if (i0 + kRK <= end) {
                load_k<VEC>(P.p[0], P.stride[0], i0, x);
                if (ND >= 2) load_k<VEC>(P.p[1], P.stride[1], i0, y);
                if (ND >= 3) load_k<VEC>(P.p[2], P.stride[2], i0, z);
                if (W && (!F || P.p[3])) load_k<VEC>(P.p[3], P.stride[3], i0, w);
            } else {
#pragma unroll
                for (int j = 0; j < kRK; ++j) {
                    const bool in = i0 + j < end;
                    x[j] = in ? P.p[0][(i0 + j) * P.stride[0]] : 0.0;
                    if (ND >= 2) y[j] = in ? P.p[1][(i0 + j) * P.stride[1]] : 0.0;
                    if (ND >= 3) z[j] = in ? P.p[2][(i0 + j) * P.stride[2]] : 0.0;
                    if (W && (!F || P.p[3])) w[j] = in ? P.p[3][(i0 + j) * P.stride[3]] : 0.0;
                }
            }
        };
#pragma unroll 1
        for (int bt = 0; bt < kRB; ++bt) {
            const int64_t i0 = t0 + bt * kRK;
            if (i0 >= end) break;
            load_batch(i0);
            const int valid = i0 + kRK <= end ? kRK : (int)(end - i0);      // particles of this batch that exist
            if constexpr (F) {
                if (!P.p[3]) {
#pragma unroll
                    for (int j = 0; j < kRK; ++j) w[j] = 1.0;
                }
#pragma unroll
                for (int j = 0; j < kRK; ++j)
                    if (j < valid) w[j] = __dmul_rn(w[j], fused_factor(P, i0 + j));
            }
#pragma unroll
            for (int j = 0; j < kRK; ++j) {
                int b[3];
                double wx[NW], wy[NW], wz[NW];
                const bool ok = particle_stencil<ND, S>(P, x[j], y[j], z[j], b, wx, wy, wz) && j < valid &&
                                particle_selected(P, i0 + j);
                if (!ok) continue;
                const double wj = W ? w[j] : 1.0;
                if (WD && W && !(fabs(wj) <= 1.79769313486231570e308)) {
                    // NaN / inf weight: 0 * w would poison the widened cells -> deposit this particle alone
#pragma unroll
                    for (int a = 0; a < NW; ++a)
#pragma unroll
                        for (int c = 0; c < (ND >= 2 ? NW : 1); ++c)
#pragma unroll
                            for (int d = 0; d < (ND >= 3 ? NW : 1); ++d) {
                                red_add(P.grid + (((int64_t)(b[0] + a)) * g1 + (b[1] + c)) * g2 + (b[2] + d),
                                        stencil_value<ND, S, true>(wx, wy, wz, a, c, d, wj));
                            }
                    continue;
                }
                bool fits = cur >= 0;
#pragma unroll
                for (int a = 0; a < ND; ++a) fits = fits && (unsigned)(b[a] - ab[a]) <= (unsigned)WD;
                if (!fits) {                 // rare on cell-ordered input: close the run, open a new one
                    if (cur >= 0) commit(cur, acc);
                    cur = (b[0] * P.dims[1] + b[1]) * P.dims[2] + b[2];      // dims beyond ndim are 1
#pragma unroll
                    for (int a = 0; a < 3; ++a) ab[a] = b[a];
#pragma unroll
                    for (int k = 0; k < MV; ++k) acc[k] = 0.0;
                }
                // stencil weights placed at this particle's offset (0 or 1) inside the window
                double sx[MW], sy[MW], sz[MW];
                if (WD) {
                    const bool ox = b[0] != ab[0], oy = b[1] != ab[1], oz = b[2] != ab[2];
#pragma unroll
                    for (int i = 0; i < MW; ++i) {
                        sx[i] = ox ? (i >= 1 ? wx[i - 1] : 0.0) : (i < NW ? wx[i] : 0.0);
                        sy[i] = oy ? (i >= 1 ? wy[i - 1] : 0.0) : (i < NW ? wy[i] : 0.0);
                        sz[i] = oz ? (i >= 1 ? wz[i - 1] : 0.0) : (i < NW ? wz[i] : 0.0);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < MW; ++i) { sx[i] = wx[i]; sy[i] = wy[i]; sz[i] = wz[i]; }
                }
#pragma unroll
                for (int a = 0; a < MA; ++a)
#pragma unroll
                    for (int c = 0; c < MC; ++c)
#pragma unroll
                        for (int d = 0; d < MD; ++d) {
                            // (a widened window needs the products also for shape 0: the shifted 1.0 / 0.0)
                            const double v = WD ? stencil_value<ND, 1, W>(sx, sy, sz, a, c, d, wj)
                                                : stencil_value<ND, S, W>(sx, sy, sz, a, c, d, wj);
                            const int k = (a * MC + c) * MD + d;
                            acc[k] = __dadd_rn(acc[k], v);
                        }
            }
        }
        // merge the runs still open at the end of each lane's particles across the warp
        const int key = cur >= 0 ? cur : -1 - lane;
        const int prev = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = lane == 0 || prev != key;
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        bool mine = cur >= 0;
        if (__popc(heads) <= PTG_RUNS_MERGE_MAX_HEADS) {
            const unsigned above = lane == 31 ? 0u : heads & ~((2u << lane) - 1u);
            const int next_head = above ? __ffs(above) - 1 : 32;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
                const bool take = lane + dd < next_head;
#pragma unroll
                for (int k = 0; k < MV; ++k) {
                    const double o = __shfl_down_sync(0xffffffffu, acc[k], dd);
                    // predicated add (one @P DADD) instead of the add + two 32-bit selects nvcc emits
                    asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p add.rn.f64 %0, %0, %1;\n\t}"
                        : "+d"(acc[k]) : "d"(o), "r"((int)take));
                }
            }
            mine = mine && head;
        }
        if (mine) commit(cur, acc);
    }
}

// widened windows are compiled where they fit the register file: (S+2)^ndim <= 27 accumulators
template <int ND, int S> struct WideOk {
    static constexpr int MW = StencilWidth<S>::value + 1;
    static constexpr bool value = (ND == 1 ? MW : ND == 2 ? MW * MW : MW * MW * MW) <= 27;
};

template <int ND, int S, bool W>
static cudaError_t launch_runs(const DepositParams &P, const TileConfig *gate, bool wide, int sm_count, cudaStream_t st) {
    if (P.n <= 0) return cudaSuccess;
    bool vec = true;
    for (int k = 0; k < 4; ++k) {
        const bool used = k < ND || (k == 3 && W && P.p[3] != nullptr);
        if (!used) continue;
        if (P.stride[k] != 1 || (reinterpret_cast<uintptr_t>(P.p[k]) & 31u)) vec = false;
    }
    int64_t want = (P.n + (int64_t)kRT * kRK * kRB - 1) / ((int64_t)kRT * kRK * kRB);
    int64_t cap = (int64_t)sm_count * 4;
    int blocks = (int)(want < cap ? want : cap);
    if (blocks < 1) blocks = 1;
    if (wide) {
        if constexpr (WideOk<ND, S>::value) {
            if (W && P.fuse_kind) {
                if (vec) k_deposit_runs<ND, S, W, true, 1, W><<<blocks, kRT, 0, st>>>(P, gate);
                else k_deposit_runs<ND, S, W, false, 1, W><<<blocks, kRT, 0, st>>>(P, gate);
            } else {
                if (vec) k_deposit_runs<ND, S, W, true, 1, false><<<blocks, kRT, 0, st>>>(P, gate);
                else k_deposit_runs<ND, S, W, false, 1, false><<<blocks, kRT, 0, st>>>(P, gate);
            }
        }
        return cudaGetLastError();      // (no widened instance for this stencil: nothing launched)
    }
    if (W && P.fuse_kind) {
        if (vec) k_deposit_runs<ND, S, W, true, 0, W><<<blocks, kRT, 0, st>>>(P, gate);
        else k_deposit_runs<ND, S, W, false, 0, W><<<blocks, kRT, 0, st>>>(P, gate);
    } else {
        if (vec) k_deposit_runs<ND, S, W, true, 0, false><<<blocks, kRT, 0, st>>>(P, gate);
        else k_deposit_runs<ND, S, W, false, 0, false><<<blocks, kRT, 0, st>>>(P, gate);
    }
    return cudaGetLastError();
}

template <int ND, int S>
static cudaError_t runs_w(const DepositParams &P, const TileConfig *gate, bool wide, int sm, cudaStream_t st) {
    return (P.p[3] || P.fuse_kind) ? launch_runs<ND, S, true>(P, gate, wide, sm, st) : launch_runs<ND, S, false>(P, gate, wide, sm, st);
}

template <int ND>
static cudaError_t runs_s(int shape, const DepositParams &P, const TileConfig *gate, bool wide, int sm, cudaStream_t st) {
    switch (shape) {
        case 0: return runs_w<ND, 0>(P, gate, wide, sm, st);
        case 1: return runs_w<ND, 1>(P, gate, wide, sm, st);
        case 2: return runs_w<ND, 2>(P, gate, wide, sm, st);
        default: return runs_w<ND, 3>(P, gate, wide, sm, st);
    }
}

bool runs_wide_available(int ndim, int shape) {
    const int mw = (shape == 0 ? 1 : shape + 1) + 1;
    int mv = 1;
    for (int a = 0; a < ndim; ++a) mv *= mw;
    return mv <= 27;
}

cudaError_t deposit_runs(int ndim, int shape, const DepositParams &P, const TileConfig *gate, bool wide, int sm_count,
                         cudaStream_t st) {
    switch (ndim) {
        case 1: return runs_s<1>(shape, P, gate, wide, sm_count, st);
        case 2: return runs_s<2>(shape, P, gate, wide, sm_count, st);
        default: return runs_s<3>(shape, P, gate, wide, sm_count, st);
    }
}

}  // namespace ptg
