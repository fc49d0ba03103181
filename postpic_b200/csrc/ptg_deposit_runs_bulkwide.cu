// ptg_deposit_runs_bulkwide.cu -- method RUNS_WIDE fed by the copy engine.
//
// Cell-ordered input with an ODD particle shape (CIC, spline3: the base cell is the nearest node, f or f + 1
// for a particle of floor cell f) and everything in 3-D needs the register window of the open run widened by
// one cell per axis (ptg_deposit_runs.cu, k_deposit_runs<.., WD = 1, ..>).  Those windows are large -- 9 (2-D
// CIC), 25 (2-D spline3) accumulators per thread --, the vector-load kernel holds the latter at 2 CTAs
// of 8 warps per SM, and at that occupancy it waits on its own particle loads half of the time
// (profiles/r02/ncu_full_runs_wide_spline3_cellsorted.txt: long-scoreboard 0.54 per issue, FP64 pipe 48 %).
// Here the particle stream is decoupled from the arithmetic exactly as in k_deposit_runs_bulk: every warp owns
// a ring of shared-memory tiles (256 consecutive particles per array) that cp.async.bulk fills, completion
// counted in bytes on an mbarrier.  Two ways to consume a tile:
//   EARLY (windows of <= 16 cells): all 8 particles of the lane go to registers, the slot is handed back at
//         once, the copy of the tile after next overlaps the arithmetic of this one;
//   LATE  (25 cells): the registers belong to the accumulators -- the lane takes its particles two at a
//         time out of shared memory and hands the slot back after the last pair; the arithmetic of one tile
//         (~150 FP64 instructions per particle) is longer than a copy takes to land, so one tile of lead is enough.
// Run accumulation, floor-cell anchoring, warp merge and commit are those of k_deposit_runs.
//
// Replaces the loops of /root/reference/postpic/particles/_particlestogrid.pyx
// :70-129 (1-D), :181-276 (2-D), :341-463 (3-D).
#include <cstdlib>

#include "ptg_internal.cuh"
#include "ptg_runs_common.cuh"

namespace ptg {

// One particle into the open run of this lane (widened window, WD = 1).
template <int ND, int S, bool W, int MV>
__device__ __forceinline__ void wide_take(const DepositParams &P, int64_t g1, int64_t g2, double x, double y, double z,
                                          double wj, int &cur, int (&ab)[3], double (&acc)[MV]) {
    constexpr int NW = StencilWidth<S>::value, MW = NW + 1;
    constexpr int MA = MW, MC = ND >= 2 ? MW : 1, MD = ND >= 3 ? MW : 1;
    static_assert(MV == MA * MC * MD, "window size");
    int b[3];
    double wx[NW], wy[NW], wz[NW];
    if (!particle_stencil<ND, S>(P, x, y, z, b, wx, wy, wz)) return;
    if (W && !(fabs(wj) <= 1.79769313486231570e308)) {
        // NaN / inf weight: 0 * w would poison the widened cells -> deposit this particle alone
#pragma unroll
        for (int a = 0; a < NW; ++a)
#pragma unroll
            for (int c = 0; c < (ND >= 2 ? NW : 1); ++c)
#pragma unroll
                for (int d = 0; d < (ND >= 3 ? NW : 1); ++d)
                    red_add(P.grid + (((int64_t)(b[0] + a)) * g1 + (b[1] + c)) * g2 + (b[2] + d),
                            stencil_value<ND, S, true>(wx, wy, wz, a, c, d, wj));
        return;
    }
    bool fits = cur >= 0;
#pragma unroll
    for (int a = 0; a < ND; ++a) fits = fits && (unsigned)(b[a] - ab[a]) <= 1u;
    if (!fits) {                 // rare on cell-ordered input: close the run, open a new one
        if (cur >= 0) runs_commit<MA, MC, MD, true>(P.grid, g1, g2, cur, acc);
#pragma unroll
        for (int a = 0; a < 3; ++a) ab[a] = b[a];
        if (S & 1) {
            // odd shapes: anchor the window at the FLOOR cell, every particle of the cell then fits
            // (see k_deposit_runs)
            const double c3[3] = {x, ND >= 2 ? y : 0.0, ND >= 3 ? z : 0.0};
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
    const bool ox = b[0] != ab[0], oy = b[1] != ab[1], oz = b[2] != ab[2];
#pragma unroll
    for (int i = 0; i < MW; ++i) {
        sx[i] = ox ? (i >= 1 ? wx[i - 1] : 0.0) : (i < NW ? wx[i] : 0.0);
        sy[i] = oy ? (i >= 1 ? wy[i - 1] : 0.0) : (i < NW ? wy[i] : 0.0);
        sz[i] = oz ? (i >= 1 ? wz[i - 1] : 0.0) : (i < NW ? wz[i] : 0.0);
    }
#pragma unroll
    for (int a = 0; a < MA; ++a)
#pragma unroll
        for (int c = 0; c < MC; ++c)
#pragma unroll
            for (int d = 0; d < MD; ++d) {
                // (the products are needed also for shape 0: the shifted 1.0 / 0.0)
                const int k = (a * MC + c) * MD + d;
                acc[k] = __dadd_rn(acc[k], stencil_value<ND, 1, W>(sx, sy, sz, a, c, d, wj));
            }
}

template <int ND, int S, bool W, bool LATE>
__global__ void __launch_bounds__(BulkCfg<ND, W>::THREADS, 1) k_deposit_runs_bulk_wide(const DepositParams P, const TileConfig *gate) {
    if (gate != nullptr && gate->mode != kModeRunsWide) return;
    using C = BulkCfg<ND, W>;
    constexpr int NW = StencilWidth<S>::value, MW = NW + 1;
    constexpr int MA = MW, MC = ND >= 2 ? MW : 1, MD = ND >= 3 ? MW : 1, MV = MA * MC * MD;
    constexpr int NA = C::NA, STAGES = C::STAGES;
    static_assert(kBulkPer == 8, "granule rotation below assumes 8 particles per lane and tile");
    extern __shared__ __align__(128) unsigned char bulk_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *const ring = reinterpret_cast<double *>(bulk_smem) + (size_t)warp * STAGES * NA * kBulkTile;
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
    const int64_t per_warp = (n_tiles + GW - 1) / GW;
    const int64_t t_begin = gw * per_warp < n_tiles ? gw * per_warp : n_tiles;
    const int64_t t_end = t_begin + per_warp < n_tiles ? t_begin + per_warp : n_tiles;

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
            if (t_begin + s < t_end) issue(t_begin + s, s);
    }

    for (int64_t t = t_begin; t < t_end; ++t) {
        const int64_t k = t - t_begin;
        const int s = (int)(k % STAGES);
        const unsigned parity = (unsigned)((k / STAGES) & 1);
        mbar_wait(bars + s, parity);
        const double *mine = ring + (size_t)s * NA * kBulkTile + lane * kBulkPer;
        // hand the slot back: see k_deposit_runs_bulk -- a word computed from every shared load is stored in
        // front of the release-arrive, so the loads have RETURNED before the copy engine may refill the slot
        auto hand_back = [&](unsigned chk) {
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(smem_u32(sink)), "r"(chk) : "memory");
            mbar_arrive(empty + s);
            if (lane == 0 && t + STAGES < t_end) {
                mbar_wait(empty + s, parity);
                issue(t + STAGES, s);
            }
        };

        int cur = -1;
        int ab[3] = {0, 0, 0};
        double acc[MV];
#pragma unroll
        for (int i = 0; i < MV; ++i) acc[i] = 0.0;

        if constexpr (!LATE) {
            double v[NA][kBulkPer];
            unsigned chk = 0;
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int q = 0; q < kBulkPer / 2; ++q) {
                    const int gq = (q + lane * kBulkPer / 16) & (kBulkPer / 2 - 1);
                    const double2 two = *reinterpret_cast<const double2 *>(mine + a * kBulkTile + 2 * gq);
                    v[a][2 * q] = two.x;
                    v[a][2 * q + 1] = two.y;
                    chk ^= (unsigned)__double2hiint(two.x);
                }
            hand_back(chk);
#pragma unroll
            for (int j = 0; j < kBulkPer; ++j) {
                const double y = ND >= 2 ? v[ND >= 2 ? 1 : 0][j] : 0.0, z = ND >= 3 ? v[ND >= 3 ? 2 : 0][j] : 0.0;
                wide_take<ND, S, W, MV>(P, g1, g2, v[0][j], y, z, W ? v[NA - 1][j] : 1.0, cur, ab, acc);
            }
        } else {
            unsigned chk = 0;
#pragma unroll 1
            for (int q = 0; q < kBulkPer / 2; ++q) {
                // granule order rotated by lane: the eight lanes of a 128-bit phase fall on eight different bank groups
                const int gq = (q + lane * kBulkPer / 16) & (kBulkPer / 2 - 1);
                double2 two[NA];
#pragma unroll
                for (int a = 0; a < NA; ++a) {
                    two[a] = *reinterpret_cast<const double2 *>(mine + a * kBulkTile + 2 * gq);
                    chk ^= (unsigned)__double2hiint(two[a].x);
                }
                wide_take<ND, S, W, MV>(P, g1, g2, two[0].x, ND >= 2 ? two[ND >= 2 ? 1 : 0].x : 0.0,
                                        ND >= 3 ? two[ND >= 3 ? 2 : 0].x : 0.0, W ? two[NA - 1].x : 1.0, cur, ab, acc);
                wide_take<ND, S, W, MV>(P, g1, g2, two[0].y, ND >= 2 ? two[ND >= 2 ? 1 : 0].y : 0.0,
                                        ND >= 3 ? two[ND >= 3 ? 2 : 0].y : 0.0, W ? two[NA - 1].y : 1.0, cur, ab, acc);
            }
            hand_back(chk);
        }
        if (runs_warp_merge<MV>(cur, acc, lane)) runs_commit<MA, MC, MD, true>(P.grid, g1, g2, cur, acc);
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
            for (int a = 0; a < NW; ++a)
#pragma unroll
                for (int c = 0; c < (ND >= 2 ? NW : 1); ++c)
#pragma unroll
                    for (int d = 0; d < (ND >= 3 ? NW : 1); ++d)
                        red_add(g + (a * g1 + c) * g2 + d, stencil_value<ND, S, W>(wx, wy, wz, a, c, d, wj));
        }
    }
}

template <int ND, int S, bool W>
static cudaError_t launch(const DepositParams &P, const TileConfig *gate, int sm_count, cudaStream_t st) {
    using C = BulkCfg<ND, W>;
    constexpr int MW = StencilWidth<S>::value + 1;
    constexpr bool LATE = (ND == 1 ? MW : ND == 2 ? MW * MW : MW * MW * MW) > 16;
    static bool opted_in[64] = {};                           // per device: > 48 KB of dynamic shared memory
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !opted_in[dev]) {
        e = cudaFuncSetAttribute(k_deposit_runs_bulk_wide<ND, S, W, LATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) opted_in[dev] = true;
    }
    const int64_t want = (P.n / kBulkTile + C::WARPS - 1) / C::WARPS;
    const int blocks = (int)(want < sm_count ? (want < 1 ? 1 : want) : sm_count);     // one CTA per SM
    k_deposit_runs_bulk_wide<ND, S, W, LATE><<<blocks, C::THREADS, C::SMEM, st>>>(P, gate);
    return cudaGetLastError();
}

template <int ND, int S>
static cudaError_t launch_w(const DepositParams &P, const TileConfig *gate, int sm, cudaStream_t st) {
    return P.p[3] ? launch<ND, S, true>(P, gate, sm, st) : launch<ND, S, false>(P, gate, sm, st);
}

// Instances: the odd shapes in 2-D.  Everything else keeps k_deposit_runs: 1-D and the even shapes (which only come
// here when the caller names RUNS_WIDE), and 3-D CIC -- measured on cell-ordered cfg3 (7.5 particles per cell: every lane
// commits its own 27-cell window, the kernel is bound by those reductions, not by its loads) 46.2 vs 47.8 Gp/s with the
// vector-load kernel (profiles/r02/perf_bulkwide.log).
bool runs_bulk_wide_applies(int ndim, int shape, const DepositParams &P) {
    static const int forced = [] { const char *e = getenv("PTG_RUNS_IMPL"); return !e ? 0 : e[0] == 'b' ? 1 : 2; }();
    if (forced == 2) return false;                       // PTG_RUNS_IMPL=v keeps the vector-load kernel (A/B runs)
    if (!(ndim == 2 && (shape == 1 || shape == 3))) return false;
    if (P.mask != nullptr || fused_mode(P) || P.n < (int64_t(1) << 13)) return false;
    for (int k = 0; k < 4; ++k) {
        const bool used = k < ndim || (k == 3 && P.p[3] != nullptr);
        if (used && (P.stride[k] != 1 || (reinterpret_cast<uintptr_t>(P.p[k]) & 15u))) return false;
    }
    return true;
}

cudaError_t deposit_runs_bulk_wide(int ndim, int shape, const DepositParams &P, const TileConfig *gate, int sm_count,
                                   cudaStream_t st) {
    if (ndim == 2 && shape == 1) return launch_w<2, 1>(P, gate, sm_count, st);
    if (ndim == 2 && shape == 3) return launch_w<2, 3>(P, gate, sm_count, st);
    return cudaErrorInvalidValue;
}

}  // namespace ptg
