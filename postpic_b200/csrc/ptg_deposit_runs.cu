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
#include <cstdlib>

#include "ptg_internal.cuh"
#include "ptg_runs_common.cuh"

namespace ptg {

// PTG_RUNS_IMPL=b|v forces the bulk-copy / the vector-load kernel wherever it exists (A/B runs, tests)

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

    auto commit = [&](int key, const double (&acc)[MV]) { runs_commit<MA, MC, MD, WD != 0>(P.grid, g1, g2, key, acc); };

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
            if (F && P.any_akind) {             // axis expressions: evaluated particle by particle
#pragma unroll
                for (int j = 0; j < kRK; ++j) {
                    const bool in = i0 + j < end;
                    x[j] = in ? particle_coord<F>(P, 0, i0 + j) : 0.0;
                    if (ND >= 2) y[j] = in ? particle_coord<F>(P, 1, i0 + j) : 0.0;
                    if (ND >= 3) z[j] = in ? particle_coord<F>(P, 2, i0 + j) : 0.0;
                    if (W && P.p[3]) w[j] = in ? P.p[3][(i0 + j) * P.stride[3]] : 0.0;
                }
            } else if (i0 + kRK <= end) {
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
                    for (int j = 0; j < kRK; ++j) w[j] = P.f_weight;
                }
                if (P.fuse_kind != 0) {
#pragma unroll
                    for (int j = 0; j < kRK; ++j)
                        if (j < valid) w[j] = __dmul_rn(w[j], fused_factor(P, i0 + j));
                }
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
#pragma unroll
                    for (int a = 0; a < 3; ++a) ab[a] = b[a];
                    if (WD && (S & 1)) {
                        // Odd shapes: the base cell is the NEAREST node, f or f + 1 for a particle of floor cell f.
                        // Anchor the widened window at the floor cell, not at this particle's base: every particle of
                        // the cell then fits (offset 0 or 1 on each axis) and the run ends at the cell boundary --
                        // anchored at the first particle's base, a later particle on the lower side would close it
                        // (measured on cell-ordered cfg2: CIC 107 -> see profiles/r02/perf_runs_wide_anchor.log).
                        const double c3[3] = {x[j], ND >= 2 ? y[j] : 0.0, ND >= 3 ? z[j] : 0.0};
#pragma unroll
                        for (int a = 0; a < ND; ++a) {
                            const double u = scaled(c3[a], P.lo[a], P.inv[a]);
                            if (u > 0.0 && __double2int_rz(__dadd_rn(u, 0.5)) > __double2int_rz(u)) ab[a] -= 1;
                        }
                    }
                    cur = (ab[0] * P.dims[1] + ab[1]) * P.dims[2] + ab[2];      // dims beyond ndim are 1
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
        if (runs_warp_merge<MV>(cur, acc, lane)) commit(cur, acc);
    }
}

// ---------------------------------------------------------------------------------
// RUNS with the particle stream decoupled from the arithmetic: every warp owns a ring of shared-memory
// tiles that the copy engine fills with 1-D bulk copies (cp.async.bulk, completion counted in bytes on an
// mbarrier), kBulkStages tiles ahead.  A tile is 256 consecutive particles (2 KB per array); lane l takes
// particles [8 l, 8 l + 8) of it into registers with four 128-bit shared loads per array (granules in a
// lane-rotated order, so the eight lanes of a 128-bit phase fall on eight different bank groups), the
// slot is handed back to the copy engine at once, and the run accumulation / warp merge / commit are
// those of k_deposit_runs.  No load instruction waits on HBM any more: the only thing a warp can stall
// on is a tile that has not landed.
// ---------------------------------------------------------------------------------
template <int ND, int S, bool W>
__global__ void __launch_bounds__(BulkCfg<ND, W>::THREADS, 1) k_deposit_runs_bulk(const DepositParams P, const TileConfig *gate) {
    if (gate != nullptr && gate->mode != kModeRuns) return;
    using C = BulkCfg<ND, W>;
    constexpr int NW = StencilWidth<S>::value;
    constexpr int MA = NW, MC = ND >= 2 ? NW : 1, MD = ND >= 3 ? NW : 1, MV = MA * MC * MD;
    constexpr int NA = C::NA, STAGES = C::STAGES;
    extern __shared__ __align__(128) unsigned char bulk_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *const ring = reinterpret_cast<double *>(bulk_smem) + (size_t)warp * STAGES * NA * kBulkTile;
    // per warp and stage: `bars` = "tile has landed" (1 arrival + the copy's bytes), `empty` = "every lane
    // has its copy of the tile" (32 arrivals)
    uint64_t *const bars = reinterpret_cast<uint64_t *>(bulk_smem + (size_t)C::WARPS * STAGES * C::TILE_BYTES) + warp * 2 * STAGES;
    uint64_t *const empty = bars + STAGES;
    unsigned *const sink = reinterpret_cast<unsigned *>(bulk_smem + (size_t)C::WARPS * STAGES * (C::TILE_BYTES + 16)) + threadIdx.x;
    const int64_t g1 = ND >= 2 ? P.dims[1] : 1, g2 = ND >= 3 ? P.dims[2] : 1;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(bars + s, 1); mbar_init(empty + s, 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    // every warp owns a contiguous range of whole tiles
    const int64_t n_tiles = P.n / kBulkTile;
    const int64_t gw = (int64_t)blockIdx.x * C::WARPS + warp, GW = (int64_t)gridDim.x * C::WARPS;
#if PTG_BULK_INTERLEAVE
    // the warps of a CTA take adjacent tiles in turn: the CTA streams one contiguous range
    const int64_t per_cta = ((n_tiles + gridDim.x - 1) / gridDim.x + C::WARPS - 1) / C::WARPS * C::WARPS;
    const int64_t c_begin = (int64_t)blockIdx.x * per_cta < n_tiles ? (int64_t)blockIdx.x * per_cta : n_tiles;
    const int64_t c_end = c_begin + per_cta < n_tiles ? c_begin + per_cta : n_tiles;
    constexpr int64_t t_step = C::WARPS;
    const int64_t t_begin = c_begin + warp < c_end ? c_begin + warp : c_end;
    const int64_t t_end = c_end;
#else
    const int64_t per_warp = (n_tiles + GW - 1) / GW;
    constexpr int64_t t_step = 1;
    const int64_t t_begin = gw * per_warp < n_tiles ? gw * per_warp : n_tiles;
    const int64_t t_end = t_begin + per_warp < n_tiles ? t_begin + per_warp : n_tiles;
#endif

    auto issue = [&](int64_t t, int s) {                     // lane 0 only
        mbar_expect_tx(bars + s, (unsigned)C::TILE_BYTES);
#pragma unroll
        for (int a = 0; a < NA; ++a) {
            const double *src = (a < ND ? P.p[a] : P.p[3]) + t * kBulkTile;
            bulk_load(ring + ((size_t)s * NA + a) * kBulkTile, src, kBulkTile * 8, bars + s);
        }
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s)
            if (t_begin + s * t_step < t_end) issue(t_begin + s * t_step, s);
    }

    for (int64_t t = t_begin; t < t_end; t += t_step) {
        const int64_t k = (t - t_begin) / t_step;
        const int s = (int)(k % STAGES);
        mbar_wait(bars + s, (unsigned)((k / STAGES) & 1));
        // this lane's particles -> registers (granule order rotated by lane: conflict-free 128-bit phases)
        double v[NA][kBulkPer];
        const double *mine = ring + (size_t)s * NA * kBulkTile + lane * kBulkPer;
#pragma unroll
        for (int a = 0; a < NA; ++a)
#pragma unroll
            for (int q = 0; q < kBulkPer / 2; ++q) {
                const int gq = (q + lane * kBulkPer / 16) & (kBulkPer / 2 - 1);
                const double2 two = *reinterpret_cast<const double2 *>(mine + a * kBulkTile + 2 * gq);
                v[a][2 * q] = two.x;
                v[a][2 * q + 1] = two.y;
            }
        // Hand the slot back.  The copy engine may overwrite it only after every lane's shared loads have
        // RETURNED, not merely issued: behind a queue of reductions that wait on L2 a load can sit in the
        // memory pipeline for microseconds, longer than the next bulk copy takes to land (seen as particles
        // whose x, y and weight came from different tiles; profiles/r01/bulk_runs_diag_before_fix.log).
        // (1) a word computed from one register of each load is stored to shared memory in front of the
        // arrive (a store cannot be optimised away, and it cannot issue before the loads have delivered);
        // (2) the arrive is a release on the slot's `empty` barrier, and lane 0 refills the slot only after
        // all 32 arrivals.
        {
            unsigned chk = 0;
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int q = 0; q < kBulkPer / 2; ++q) chk ^= (unsigned)__double2hiint(v[a][2 * q]);
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(smem_u32(sink)), "r"(chk) : "memory");
        }
        mbar_arrive(empty + s);
        if (lane == 0 && t + STAGES * t_step < t_end) {
            mbar_wait(empty + s, (unsigned)((k / STAGES) & 1));
            issue(t + STAGES * t_step, s);
        }

        int cur = -1;
        double acc[MV];
#pragma unroll
        for (int i = 0; i < MV; ++i) acc[i] = 0.0;
#pragma unroll
        for (int j = 0; j < kBulkPer; ++j) {
            int b[3];
            double wx[NW], wy[NW], wz[NW];
            const double y = ND >= 2 ? v[ND >= 2 ? 1 : 0][j] : 0.0, z = ND >= 3 ? v[ND >= 3 ? 2 : 0][j] : 0.0;
            if (!particle_stencil<ND, S>(P, v[0][j], y, z, b, wx, wy, wz)) continue;
            const double wj = W ? v[NA - 1][j] : 1.0;
            const int key = (b[0] * P.dims[1] + b[1]) * P.dims[2] + b[2];            // dims beyond ndim are 1
            if (key != cur) {                // rare on cell-ordered input: close the run, open a new one
                if (cur >= 0) runs_commit<MA, MC, MD, false>(P.grid, g1, g2, cur, acc);
                cur = key;
#pragma unroll
                for (int i = 0; i < MV; ++i) acc[i] = 0.0;
            }
#pragma unroll
            for (int a = 0; a < MA; ++a)
#pragma unroll
                for (int c = 0; c < MC; ++c)
#pragma unroll
                    for (int d = 0; d < MD; ++d) {
                        const int i = (a * MC + c) * MD + d;
                        acc[i] = __dadd_rn(acc[i], stencil_value<ND, S, W>(wx, wy, wz, a, c, d, wj));
                    }
        }
        if (runs_warp_merge<MV>(cur, acc, lane)) runs_commit<MA, MC, MD, false>(P.grid, g1, g2, cur, acc);
    }

    // the particles behind the last whole tile (< 256): one per lane, straight to the grid
    if (gw == 0) {
        for (int64_t i = n_tiles * kBulkTile + lane; i < P.n; i += 32) {
            int b[3];
            double wx[NW], wy[NW], wz[NW];
            const double x = P.p[0][i], y = ND >= 2 ? P.p[1][i] : 0.0, z = ND >= 3 ? P.p[2][i] : 0.0;
            if (!particle_stencil<ND, S>(P, x, y, z, b, wx, wy, wz)) continue;
            const double wj = W ? P.p[3][i] : 1.0;
            double *g = P.grid + ((int64_t)b[0] * g1 + b[1]) * g2 + b[2];
#pragma unroll
            for (int a = 0; a < MA; ++a)
#pragma unroll
                for (int c = 0; c < MC; ++c)
#pragma unroll
                    for (int d = 0; d < MD; ++d)
                        red_add(g + (a * g1 + c) * g2 + d, stencil_value<ND, S, W>(wx, wy, wz, a, c, d, wj));
        }
    }
}

template <int ND, int S, bool W>
static cudaError_t launch_runs_bulk(const DepositParams &P, const TileConfig *gate, int sm_count, cudaStream_t st) {
    using C = BulkCfg<ND, W>;
    static bool opted_in[64] = {};                           // per device: > 48 KB of dynamic shared memory
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !opted_in[dev]) {
        e = cudaFuncSetAttribute(k_deposit_runs_bulk<ND, S, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) opted_in[dev] = true;
    }
    const int64_t want = (P.n / kBulkTile + C::WARPS - 1) / C::WARPS;
    const int blocks = (int)(want < sm_count ? (want < 1 ? 1 : want) : sm_count);     // one CTA per SM
    k_deposit_runs_bulk<ND, S, W><<<blocks, C::THREADS, C::SMEM, st>>>(P, gate);
    return cudaGetLastError();
}

// the bulk-copy kernel streams plain contiguous arrays only
template <int ND, int S, bool W>
static bool bulk_ok(const DepositParams &P) {
    constexpr int NW = StencilWidth<S>::value;
    if ((ND == 1 ? NW : ND == 2 ? NW * NW : NW * NW * NW) > 16) return false;      // accumulators + 8 particles in registers
    static const int forced = [] { const char *e = getenv("PTG_RUNS_IMPL"); return !e ? 0 : e[0] == 'b' ? 1 : 2; }();
    // Measured on cell-ordered cfg2-like input (profiles/r01/ab_runs_bulk_*.log): 2-D TSC weighted 155.6 -> 169.7 Gp/s
    // (cfg5: 166.0 -> 186.7), 2-D NGP weighted 163.2 -> 247.6 and counting 228.4 -> 367.6 (0.87 of the HBM
    // roofline).  Default since round 2: the slot hand-back hazard of its first version (a refill overtaking shared
    // loads still queued behind reductions) is closed by the release-arrive on the per-slot `empty` barrier, and
    // the whole GPU suite incl. the BASELINE-geometry parity tests passes with it forced on every instance
    // (profiles/r02/suite_runs_bulk.log).  PTG_RUNS_IMPL=v keeps the vector-load kernel (A/B runs).
    constexpr bool by_default = true;
    if (forced == 2 || (forced == 0 && !by_default)) return false;
    if (P.mask != nullptr || fused_mode(P) || P.n < (int64_t(1) << 13)) return false;
    for (int k = 0; k < 4; ++k) {
        const bool used = k < ND || (k == 3 && W);
        if (used && (P.stride[k] != 1 || (reinterpret_cast<uintptr_t>(P.p[k]) & 15u))) return false;
    }
    return true;
}

// widened windows are compiled where they fit the register file: (S+2)^ndim <= 27 accumulators
template <int ND, int S> struct WideOk {
    static constexpr int MW = StencilWidth<S>::value + 1;
    static constexpr bool value = (ND == 1 ? MW : ND == 2 ? MW * MW : MW * MW * MW) <= 27;
};

template <int ND, int S, bool W>
static cudaError_t launch_runs(const DepositParams &P, const TileConfig *gate, bool wide, int sm_count, cudaStream_t st) {
    if (P.n <= 0) return cudaSuccess;
    bool vec = P.any_akind == 0;
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
        // odd shapes in 2-D, CIC in 3-D: the widened window fed by the copy engine (ptg_deposit_runs_bulkwide.cu)
        if (runs_bulk_wide_applies(ND, S, P)) return deposit_runs_bulk_wide(ND, S, P, gate, sm_count, st);
        if constexpr (WideOk<ND, S>::value) {
            if (W && fused_mode(P)) {
                if (vec) k_deposit_runs<ND, S, W, true, 1, W><<<blocks, kRT, 0, st>>>(P, gate);
                else k_deposit_runs<ND, S, W, false, 1, W><<<blocks, kRT, 0, st>>>(P, gate);
            } else {
                if (vec) k_deposit_runs<ND, S, W, true, 1, false><<<blocks, kRT, 0, st>>>(P, gate);
                else k_deposit_runs<ND, S, W, false, 1, false><<<blocks, kRT, 0, st>>>(P, gate);
            }
        }
        return cudaGetLastError();      // (no widened instance for this stencil: nothing launched)
    }
    if (bulk_ok<ND, S, W>(P)) return launch_runs_bulk<ND, S, W>(P, gate, sm_count, st);
    if (W && fused_mode(P)) {
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
    return (P.p[3] || fused_mode(P)) ? launch_runs<ND, S, true>(P, gate, wide, sm, st) : launch_runs<ND, S, false>(P, gate, wide, sm, st);
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

// One translation unit per number of axes: this file is compiled four times (postpic_b200/build.py) -- as itself it
// holds the dispatcher only, ptg_deposit_runs_nd{1,2,3}.cu include it with PTG_ND set and instantiate the kernels of
// that dimensionality (the build is bound by its longest translation unit).
#define PTG_CAT2(a, b) a##b
#define PTG_CAT(a, b) PTG_CAT2(a, b)
#ifdef PTG_ND
cudaError_t PTG_CAT(deposit_runs_nd, PTG_ND)(int shape, const DepositParams &P, const TileConfig *gate, bool wide, int sm_count,
                                             cudaStream_t st) {
    return runs_s<PTG_ND>(shape, P, gate, wide, sm_count, st);
}
#else
cudaError_t deposit_runs_nd1(int, const DepositParams &, const TileConfig *, bool, int, cudaStream_t);
cudaError_t deposit_runs_nd2(int, const DepositParams &, const TileConfig *, bool, int, cudaStream_t);
cudaError_t deposit_runs_nd3(int, const DepositParams &, const TileConfig *, bool, int, cudaStream_t);

bool runs_wide_available(int ndim, int shape) {
    const int mw = (shape == 0 ? 1 : shape + 1) + 1;
    int mv = 1;
    for (int a = 0; a < ndim; ++a) mv *= mw;
    return mv <= 27;
}

cudaError_t deposit_runs(int ndim, int shape, const DepositParams &P, const TileConfig *gate, bool wide, int sm_count,
                         cudaStream_t st) {
    switch (ndim) {
        case 1: return deposit_runs_nd1(shape, P, gate, wide, sm_count, st);
        case 2: return deposit_runs_nd2(shape, P, gate, wide, sm_count, st);
        default: return deposit_runs_nd3(shape, P, gate, wide, sm_count, st);
    }
}
#endif

}  // namespace ptg
