// ptg_deposit_tiled.cu -- method TILED: shared-memory privatised sub-grid tiles, and the probe of method AUTO.
//
// One persistent CTA per SM walks a contiguous span of the particle arrays and keeps a
// window ("tile") of the ghost-padded grid in its 200 KB of shared memory:
//   * particles whose whole stencil falls inside the window are accumulated there with
//     shared-memory atomics;
//   * particles outside the window go straight to the grid in HBM/L2 with RED.F64,
//     so the result is correct for any input -- the window only decides the speed;
//   * the CTA watches its hit rate and, when most particles miss, merges the tile
//     into the global grid and re-centres the window on the particles it is
//     currently reading; at the end of its span the tile is merged once more.
// The window *shape* is chosen per launch by a 16-CTA probe kernel that looks at 16
// blocks of 512 consecutive particles: extents ~ 4 x the mean absolute deviation of
// the stencil base cell inside a block, scaled to the tile budget.  For inputs without
// locality (i.i.d. positions over a grid much larger than the tile) the probe marks
// the tile as useless and one of the other kernels launched next to it does the work.
//
// Kernels in this file: k_probe; k_deposit_tile_fx (the default windowed tile, fixed point);
// k_deposit_tiled (the fp64 CAS tile with warp aggregation of runs, PTG_EXACT_ACCUM);
// k_deposit_small (whole grid in shared memory, up to 32 replicas); k_deposit_oneshot (small
// calls on small grids in one launch).
//
// Accumulation inside the tile (round 2): sm_100a has no native fp64 shared-memory atomic -- atomicAdd(double)
// is a CAS loop (ATOMS.CAST.SPIN.64) that sustains 48 Gparticles/s for a 2-D TSC stencil -- but it has a
// native 32-bit integer one (ATOMS.ADD).  The tile therefore holds 64-bit FIXED-POINT sums as two 32-bit
// words per cell: a contribution v (computed in fp64 exactly as the reference does) is rounded to
// round(v / q) quanta, the low word is added with a returning atomic whose old value yields the carry, the
// high word (plus carry) with a second one: 143 Gparticles/s for the same stencil
// (profiles/r02/microbench5_b200.log).  q = 2^(e-45), 2^e = the power of two above the largest |weight| the
// probe sampled, so each contribution is rounded by at most 2^-46 of that weight (per bin |error| <=
// n_bin * q / 2, against the 1e-12 * max|bin| bar: cfg2 2.4e-11 worst case vs 6e-11; integer weights below
// 2^45 are exact, and integer addition is associative, so the tile sums do not depend on the order of the
// atomics).  The 64-bit range is guarded, not assumed: every lane tracks the sum of |w| it has put into the
// tile and the CTA merges the tile into the fp64 grid before 2^63 quanta could be reached in any cell;
// particles whose |w| exceeds 2^e (rarer than one in the probe's sample size) or is not finite take the exact
// RED.F64 path.  Flag PTG_EXACT_ACCUM keeps the fp64 CAS tile.
//
// Replaces the loops of /root/reference/postpic/particles/_particlestogrid.pyx
// :70-129 (1-D), :181-276 (2-D), :341-463 (3-D).
#include "ptg_internal.cuh"

namespace ptg {

constexpr int kTT = 512;                  // probe CTA width
// CTA width of the tiled kernel by stencil size: small stencils need few registers, so more warps
// can keep shared-memory atomics in flight (one CTA per SM owns the whole 200 KB tile)
#ifndef PTG_TILE_THREADS_9
#define PTG_TILE_THREADS_9 768
#endif
template <int NV> struct TileThreads { static constexpr int value = NV <= 4 ? 1024 : NV <= 9 ? PTG_TILE_THREADS_9 : 512; };
#ifndef PTG_TILE_EVAL
#define PTG_TILE_EVAL 4
#endif
constexpr int kEvalEvery = PTG_TILE_EVAL;  // sub-chunks between hit-rate evaluations
constexpr int kTV = 4;                    // consecutive particles per lane and sub-chunk (one 256-bit load per array)

// the four consecutive doubles of one lane = one 32-byte sector: LDG.E.NA.EFL2.256, not allocated in L1, evict-first in L2
__device__ __forceinline__ void tile_ld4(const double *p, double (&v)[4]) {
    unsigned long long a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0, %1, %2, %3}, [%4];"
                 : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    v[0] = __longlong_as_double((long long)a);
    v[1] = __longlong_as_double((long long)b);
    v[2] = __longlong_as_double((long long)c);
    v[3] = __longlong_as_double((long long)d);
}

// ---------------------------------------------------------------------------------
// block-wide sum of a long long (all threads get the result)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ long long block_sum(long long v, long long *scratch /* >= 33 */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        long long t = lane < (int)(blockDim.x >> 5) ? scratch[lane] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// block-wide sums of N long longs at once (all threads get the results): one pass through shared memory instead of N
template <int N>
__device__ __forceinline__ void block_sum_n(long long (&v)[N], long long *scratch /* >= 32 * N + N */) {
#pragma unroll
    for (int k = 0; k < N; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) scratch[k * 32 + warp] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            long long t = lane < nwarp ? scratch[k * 32 + lane] : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) scratch[32 * N + k] = t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = scratch[32 * N + k];
}

// block-wide maximum of a non-negative double (all threads get the result)
__device__ __forceinline__ double block_max(double v, long long *scratch /* >= 33 */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = __double_as_longlong(v);
    __syncthreads();
    if (warp == 0) {
        double t = lane < (int)(blockDim.x >> 5) ? __longlong_as_double(scratch[lane]) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
        if (lane == 0) scratch[32] = __double_as_longlong(t);
    }
    __syncthreads();
    return __longlong_as_double(scratch[32]);
}

// ---- 64-bit fixed-point accumulation on two native 32-bit shared-memory atomics --------------------
// lo[c] / hi[c] = low / high word of cell c's two's-complement sum of quanta.  The returning add on the low
// word tells whether it wrapped; the carry rides on the (non-returning) add of the high word.  Additions
// commute, so any interleaving of the atomics of different threads gives the same 64-bit sum.
constexpr int kFxHi = kTileCells;         // hi word of cell c lives kFxHi words after its lo word
__device__ __forceinline__ void fx_add(unsigned *lo, int c, double v, double inv_q) {
    // round(v * inv_q) as a 64-bit two's-complement integer without a conversion instruction: inv_q is a power of two, so
    // the fused multiply-add rounds exactly once (to nearest even, like __double2ll_rn) at unit weight when 1.5 * 2^52 is
    // added; the low word of the sum's bit pattern IS the low word of the integer, the high word is off by 0x43380000.
    // (|v * inv_q| <= 4 * 2^45 < 2^51.)
    const double d = __fma_rn(v, inv_q, 6755399441055744.0);
    const unsigned l = (unsigned)__double2loint(d), h = (unsigned)__double2hiint(d) - 0x43380000u;
    const unsigned old = atomicAdd(lo + c, l);
    atomicAdd(lo + kFxHi + c, h + ((old + l) < old ? 1u : 0u));
}
__device__ __forceinline__ long long fx_take(unsigned *lo, int c) {         // read and clear
    const unsigned l = lo[c], h = lo[kFxHi + c];
    lo[c] = 0u;
    lo[kFxHi + c] = 0u;
    return (long long)(((unsigned long long)h << 32) | l);
}
// |stencil weight| <= 2 on an axis the reference tests on the truncated int (extrapolated weights of particles
// just below the range), <= 1 elsewhere; at most two such axes -> |contribution| <= 4 |w|.
// A cell must stay below 2^63 quanta = 2^(18+e) in units of the weights: the lanes of a CTA of TT threads may
// put sum|w| < 2^(16+e) into the tile between two merges; a lane asks for a merge once its own share reaches
// fx_lane_limit(TT, slack) * 2^e, `slack` = particles a lane may add between two checks (each |w| <= 2^e).
__host__ __device__ constexpr double fx_lane_limit(int tt, int slack) { return 65536.0 / tt - slack; }

// stencil base cell of one particle on every axis (the weights are discarded): used by the probe
template <int ND, int S>
__device__ __forceinline__ bool base_cells(const DepositParams &P, double x, double y, double z, int (&b)[3]) {
    double wx[StencilWidth<S>::value], wy[StencilWidth<S>::value], wz[StencilWidth<S>::value];
    return particle_stencil<ND, S>(P, x, y, z, b, wx, wy, wz);
}

// ---------------------------------------------------------------------------------
// probe: pick the window shape
// ---------------------------------------------------------------------------------
template <int ND, int S>
__global__ void __launch_bounds__(kTT) k_probe(const DepositParams P, TileConfig *cfg, int force, int wide_ok, int exact) {
    __shared__ long long scratch[34];
    __shared__ long long big_scratch[32 * 7 + 7];
    __shared__ bool s_last;
    constexpr int NW = StencilWidth<S>::value;
    // one probe CTA per block of kTT consecutive particles; the last CTA to finish decides
    {
        TileConfig::Part part;
        part.mad[0] = part.mad[1] = part.mad[2] = 0.f;
        part.cnt = part.eq_pairs = part.all_pairs = part.near_pairs = 0;
        part.pad = 0;
        part.wmax = 0.0;
        if (P.n >= 2 * kTT) {
            const int64_t start = (P.n - kTT) / (kProbeBlocks - 1) * blockIdx.x;
            const int64_t i = start + threadIdx.x;
            const double x = particle_coord_rt(P, 0, i);
            const double y = ND >= 2 ? particle_coord_rt(P, 1, i) : 0.0;
            const double z = ND >= 3 ? particle_coord_rt(P, 2, i) : 0.0;
            int b[3];
            const bool ok = base_cells<ND, S>(P, x, y, z, b);
            // largest finite |weight| of the sample: sets the quantum of the fixed-point tile
            double aw = 1.0;
            if (P.p[3] != nullptr || fused_mode(P)) {
                aw = fabs(P.p[3] != nullptr ? ld_stream(P.p[3] + i * P.stride[3]) : P.f_weight);
                if (P.fuse_kind != 0) aw = fabs(__dmul_rn(aw, fused_factor(P, i)));
                if (!(aw <= 1.79769313486231570e308)) aw = 0.0;
            }
            part.wmax = block_max(aw, scratch);
            // every count and sum of the sample in ONE block reduction, the deviations in a second one
            // (27 barriers less than one reduction per quantity: the probe is on the critical path of every AUTO call)
            const int key = ok ? (b[0] * P.dims[1] + b[1]) * P.dims[2] + b[2] : -1;
            const int prev = __shfl_up_sync(0xffffffffu, key, 1);
            const bool pair = (threadIdx.x & 31) != 0 && key >= 0 && prev >= 0;   // pairs inside a warp only
            // neighbours that a register window one cell wider per axis would still hold
            // (either particle of the pair may be the one further along)
            bool near = pair;
            for (int a = 0; a < ND; ++a) {
                const int pb = __shfl_up_sync(0xffffffffu, b[a], 1);
                near = near && (b[a] - pb) >= -1 && (b[a] - pb) <= 1;
            }
            long long r1[7] = {ok ? 1 : 0, pair && key == prev ? 1 : 0, pair ? 1 : 0, near ? 1 : 0,
                               ok ? b[0] : 0, ok ? b[1] : 0, ok ? b[2] : 0};
            block_sum_n<7>(r1, big_scratch);
            const long long cnt = r1[0];
            if (cnt >= 32) {
                part.cnt = (int)cnt;
                part.eq_pairs = (int)r1[1];
                part.all_pairs = (int)r1[2];
                part.near_pairs = (int)r1[3];
                long long r2[3];
                for (int a = 0; a < 3; ++a) r2[a] = ok ? llabs((long long)b[a] - r1[4 + a] / cnt) : 0;
                block_sum_n<3>(r2, big_scratch);
                for (int a = 0; a < ND; ++a) part.mad[a] = (float)((double)r2[a] / (double)cnt);
            }
        }
        if (threadIdx.x == 0) {
            cfg->part[blockIdx.x] = part;
            __threadfence();
            const unsigned t = atomicAdd(&cfg->ticket, 1u);
            s_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
    }
    // the last CTA's first warp gathers the 16 partial records in parallel (one lane per probe CTA, all fields of a
    // record in flight at once) and reduces them with shuffles: a serial loop over volatile loads cost 15 us
    if (threadIdx.x >= 32) return;
    double mad_sum[3] = {0, 0, 0};
    int blocks_used = 0;
    long long eq_pairs = 0, all_pairs = 0, near_pairs = 0;
    double wmax = 0.0;
    int runs_blocks = 0, near_blocks = 0;
    {
        const int l = threadIdx.x;
        float m0 = 0.f, m1 = 0.f, m2 = 0.f;
        int cnt = 0, eq = 0, al = 0, ne = 0;
        double wm = 0.0;
        if (l < (int)gridDim.x) {
            const TileConfig::Part *q = &cfg->part[l];
            cnt = __ldcg(&q->cnt);
            m0 = __ldcg(&q->mad[0]); m1 = __ldcg(&q->mad[1]); m2 = __ldcg(&q->mad[2]);
            eq = __ldcg(&q->eq_pairs); al = __ldcg(&q->all_pairs); ne = __ldcg(&q->near_pairs);
            wm = __ldcg(&q->wmax);
        }
        const bool used = cnt >= 32;
        // blocks that are ordered in themselves: a dump that is ordered only in parts must not be sent to RUNS on the
        // strength of the average
        runs_blocks = __popc(__ballot_sync(0xffffffffu, used && al > 0 && eq * 2 >= al));
        near_blocks = __popc(__ballot_sync(0xffffffffu, used && al > 0 && ne * 4 >= 3 * al));
        double v[3] = {used ? (double)m0 : 0.0, used ? (double)m1 : 0.0, used ? (double)m2 : 0.0};
        int bu = used ? 1 : 0, e1 = used ? eq : 0, a1 = used ? al : 0, n1 = used ? ne : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int a = 0; a < 3; ++a) v[a] += __shfl_xor_sync(0xffffffffu, v[a], o);
            bu += __shfl_xor_sync(0xffffffffu, bu, o);
            e1 += __shfl_xor_sync(0xffffffffu, e1, o);
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            n1 += __shfl_xor_sync(0xffffffffu, n1, o);
            wm = fmax(wm, __shfl_xor_sync(0xffffffffu, wm, o));
        }
        for (int a = 0; a < 3; ++a) mad_sum[a] = v[a];
        blocks_used = bu; eq_pairs = e1; all_pairs = a1; near_pairs = n1; wmax = wm;
    }
    if (threadIdx.x != 0) return;
    struct { int mode; int ext[3]; float est_hit; float run_frac; float near_frac; } c;
    c.mode = kModeRows;
    c.est_hit = 0.f;
    c.run_frac = all_pairs > 0 ? (float)((double)eq_pairs / (double)all_pairs) : 0.f;
    c.near_frac = all_pairs > 0 ? (float)((double)near_pairs / (double)all_pairs) : 0.f;
    c.ext[0] = c.ext[1] = c.ext[2] = 1;
    if (blocks_used > 0 || force) {
        double need[3] = {1, 1, 1}, prod = 1;
        for (int a = 0; a < ND; ++a) {
            const double mad = blocks_used ? mad_sum[a] / blocks_used : P.dims[a];
            need[a] = 4.0 * mad + NW + 3;
            // a spread that covers (nearly) the whole axis is the whole axis: the 512-particle sample
            // of a uniform axis gives 4*MAD within a few per cent of its length
            if (need[a] > 0.85 * P.dims[a]) need[a] = P.dims[a];
            prod *= need[a];
        }
        double ext[3] = {need[0], need[1], need[2]};
        if (prod > kTileCells) {
            // shrink all axes by a common factor, never below the stencil width
            const double ratio = (double)kTileCells / prod;
            const double f = ND == 1 ? ratio : ND == 2 ? sqrt(ratio) : cbrt(ratio);
            for (int a = 0; a < ND; ++a) { ext[a] = floor(need[a] * f); if (ext[a] < NW + 1) ext[a] = NW + 1; }
        }
        for (int a = 0; a < ND; ++a) c.ext[a] = (int)ext[a] < P.dims[a] ? (int)ext[a] : P.dims[a];
        // the last axis keeps an even extent when it is a true window (16-byte alignment of rows)
        if (c.ext[ND - 1] < P.dims[ND - 1] && (c.ext[ND - 1] & 1)) c.ext[ND - 1] -= 1;
        long long cells = 1;
        for (int a = 0; a < ND; ++a) cells *= c.ext[a];
        while (cells > kTileCells) {                       // rounding safety
            int big = 0;
            for (int a = 1; a < ND; ++a) if (c.ext[a] > c.ext[big]) big = a;
            cells = cells / c.ext[big] * (c.ext[big] - 2);
            c.ext[big] -= 2;
        }
        // spend what is left of the budget on the slowest axis (bands of whole rows)
        const long long rest = cells / c.ext[0];
        long long rows = kTileCells / rest;
        if (rows > P.dims[0]) rows = P.dims[0];
        if (rows > c.ext[0]) c.ext[0] = (int)rows;
        double hit = 1;
        for (int a = 0; a < ND; ++a) hit *= fmin(1.0, (double)c.ext[a] / need[a]);
        c.est_hit = (float)hit;
        // With the fp64 CAS tile (PTG_EXACT_ACCUM) a band of whole rows (x-ordered input, uniform along the
        // last axis) is no win: the CAS loop runs at about the rate of the L2 reductions of method ROWS
        // (measured 36 vs 38 Gp/s, 2-D TSC), so the tile is chosen only when it is a true window on the last
        // axis (bunches) or holds the whole grid (1-D and small histograms).  The fixed-point tile is ~3x
        // faster than either and takes every input that mostly hits the window.
        bool whole = true;
        for (int a = 0; a < ND; ++a) whole = whole && c.ext[a] >= P.dims[a];
        const bool windowed_last = c.ext[ND - 1] < P.dims[ND - 1];
        const bool fx = !exact && wmax > 0.0;
        if (force || (hit >= 0.45 && (whole || windowed_last || ND == 1 || fx))) c.mode = kModeTiled;
    }
    // long runs of equal base cells: register run-length accumulation beats the tile
    // (break-even is an average run of ~2.5 particles)
    // (and in at least three quarters of the sampled blocks individually, see above)
    if (!force && c.run_frac >= 0.6f && runs_blocks * 4 >= blocks_used * 3) c.mode = kModeRuns;
    else if (!force && c.near_frac >= 0.75f && near_blocks * 4 >= blocks_used * 3 && wide_ok) c.mode = kModeRunsWide;
    // fixed-point scale: 2^e = the power of two above the largest sampled |w|
    {
        double inv_q = 0.0, q = 0.0, wcap = 0.0;
        if (wmax > 0.0) {
            const int e = ilogb(wmax) + 1;
            if (e > -900 && e < 40) { q = scalbn(1.0, e - 45); inv_q = scalbn(1.0, 45 - e); wcap = scalbn(1.0, e); }
        }
        cfg->fx_inv_q = inv_q;
        cfg->fx_q = q;
        cfg->fx_wcap = wcap;
    }
    cfg->mode = c.mode;
    for (int a = 0; a < 3; ++a) cfg->ext[a] = c.ext[a];
    cfg->est_hit = c.est_hit;
    cfg->run_frac = c.run_frac;
    cfg->near_frac = c.near_frac;
    cfg->ticket = 0;                      // ready for the next call that gets this ring entry
}

// ---------------------------------------------------------------------------------
// the tiled deposit kernel
// ---------------------------------------------------------------------------------
struct Window {
    int o[3];      // origin (padded cell index of the first cell held)
    int e[3];      // extent
};

template <int ND, int TT, bool FX>
__device__ __forceinline__ void flush_tile(double *tile, const Window &wd, const DepositParams &P, int cells,
                                           int replicas, double fx_q) {
    // merge the tile (all its replicas) into the global grid, coalesced along the last axis, and clear it
    for (int c = threadIdx.x; c < cells; c += TT) {
        double v;
        long long kq = 0;
        if constexpr (FX) {
            // replicas are summed as integers (exact), one conversion and one multiply by the quantum (a power of two)
            unsigned *lo = reinterpret_cast<unsigned *>(tile);
            long long k = fx_take(lo, c);
            for (int r = 1; r < replicas; ++r) k += fx_take(lo, r * cells + c);
            v = __dmul_rn(__ll2double_rn(k), fx_q);
            kq = k;
        } else {
            v = tile[c];
            tile[c] = 0.0;
            for (int r = 1; r < replicas; ++r) { v = __dadd_rn(v, tile[r * cells + c]); tile[r * cells + c] = 0.0; }
        }
        if (v != 0.0) {
            int64_t lin;
            if (ND == 1) lin = wd.o[0] + c;
            else if (ND == 2) {
                const int ix = c / wd.e[1], iy = c - ix * wd.e[1];
                lin = (int64_t)(wd.o[0] + ix) * P.dims[1] + wd.o[1] + iy;
            } else {
                const int plane = wd.e[1] * wd.e[2];
                const int ix = c / plane, r = c - ix * plane;
                const int iy = r / wd.e[2], iz = r - iy * wd.e[2];
                lin = ((int64_t)(wd.o[0] + ix) * P.dims[1] + wd.o[1] + iy) * P.dims[2] + wd.o[2] + iz;
            }
            if (FX && P.fixed_lo != nullptr) fixed_accumulate(P, lin, kq);   // method FIXED: stay integer
            else red_add(P.grid + lin, v);
        }
    }
}

// (the fp64 tile: method 'tiled_exact' / PTG_EXACT_ACCUM.  The default tile kernel is k_deposit_tile_fx below.)
template <int ND, int S, bool W, int TT, bool F>
__global__ void __launch_bounds__(TT, 1) k_deposit_tiled(const DepositParams P, const TileConfig *cfgp) {
    extern __shared__ double tile[];
    __shared__ long long scratch[34];
    __shared__ Window s_win;
    __shared__ unsigned s_acc, s_miss;
    constexpr int NW = StencilWidth<S>::value;

    if (cfgp->mode != kModeTiled) return;
    const TileConfig cfg = *cfgp;
    int cells = 1;
    bool whole = true;                       // the window holds the whole grid (small 1-D / 2-D histograms)
    for (int a = 0; a < ND; ++a) { cells *= cfg.ext[a]; whole = whole && cfg.ext[a] >= P.dims[a]; }
    // a grid that fits several times is replicated: warp w accumulates into replica w % R, which
    // divides the same-address contention of the shared-memory atomics by R
    int replicas = whole ? kTileCells / cells : 1;
    if (replicas > TT / 32) replicas = TT / 32;
    if (replicas < 1) replicas = 1;
    double *mytile = tile + ((threadIdx.x >> 5) % replicas) * cells;

    for (int c = threadIdx.x; c < cells * replicas; c += TT) tile[c] = 0.0;
    if (threadIdx.x == 0) {
        for (int a = 0; a < 3; ++a) { s_win.e[a] = a < ND ? cfg.ext[a] : 1; s_win.o[a] = 0; }
        // a window that covers the whole axis never needs to move; the others start "nowhere"
        for (int a = 0; a < ND; ++a) if (cfg.ext[a] < P.dims[a]) s_win.o[a] = -(1 << 28);
        s_acc = 0; s_miss = 0;
    }
    __syncthreads();
    Window wd = s_win;

    // contiguous span of this CTA, a multiple of the sub-chunk (kTV consecutive particles per lane)
    int64_t per = (P.n + gridDim.x - 1) / gridDim.x;
    per = (per + TT * kTV - 1) / (TT * kTV) * (TT * kTV);
    const int64_t begin = (int64_t)blockIdx.x * per;
    const int64_t end = begin + per < P.n ? begin + per : P.n;
    const int lane = threadIdx.x & 31;
    const double *px = P.p[0], *py = P.p[1], *pz = P.p[2];
    const int64_t sx = P.stride[0], sy = P.stride[1], sz = P.stride[2];
    // 256-bit loads (one sector per lane and array) when every streamed array is contiguous and 32-byte aligned
    bool vec = sx == 1 && (reinterpret_cast<uintptr_t>(px) & 31u) == 0;
    if (ND >= 2) vec = vec && sy == 1 && (reinterpret_cast<uintptr_t>(py) & 31u) == 0;
    if (ND >= 3) vec = vec && sz == 1 && (reinterpret_cast<uintptr_t>(pz) & 31u) == 0;
    if (W && !F) vec = vec && P.stride[3] == 1 && (reinterpret_cast<uintptr_t>(P.p[3]) & 31u) == 0;
    if (F && P.any_akind) vec = false;      // axis expressions are evaluated particle by particle

    int since_eval = kEvalEvery - 1;       // first evaluation right after the first sub-chunk
    unsigned acc_w = 0, miss_w = 0;        // per-warp counters (lane 0), published at evaluation time
    for (int64_t i0 = begin; i0 < end; i0 += (int64_t)TT * kTV) {
        const int64_t ib = i0 + (int64_t)threadIdx.x * kTV;
        double xv[kTV], yv[kTV], zv[kTV], wv[kTV];
        if (vec && ib + kTV <= end) {
            tile_ld4(px + ib, xv);
            if (ND >= 2) tile_ld4(py + ib, yv);
            if (ND >= 3) tile_ld4(pz + ib, zv);
            if (W && !F) tile_ld4(P.p[3] + ib, wv);
        } else {
#pragma unroll
            for (int j = 0; j < kTV; ++j) {
                const bool in = ib + j < end;
                xv[j] = in ? particle_coord<F>(P, 0, ib + j) : 0.0;
                if (ND >= 2) yv[j] = in ? particle_coord<F>(P, 1, ib + j) : 0.0;
                if (ND >= 3) zv[j] = in ? particle_coord<F>(P, 2, ib + j) : 0.0;
                if (W && !F) wv[j] = in ? ld_stream(P.p[3] + (ib + j) * P.stride[3]) : 0.0;
            }
        }
        if constexpr (W && F) {
#pragma unroll
            for (int j = 0; j < kTV; ++j) wv[j] = ib + j < end ? particle_weight<F>(P, ib + j) : 0.0;
        }
        bool ok = false;                   // (of the lane's last particle: seeds the window after a move)
        int b[3] = {0, 0, 0};
#pragma unroll
      for (int j = 0; j < kTV; ++j) {
        const int64_t i = ib + j;
        const double x = xv[j], y = ND >= 2 ? yv[j] : 0.0, z = ND >= 3 ? zv[j] : 0.0, w = W ? wv[j] : 1.0;
        // ---- stencil in registers --------------------------------------------------
        double wx[NW], wy[NW], wz[NW];
        ok = particle_stencil<ND, S>(P, x, y, z, b, wx, wy, wz) && i < end && particle_selected(P, i);
        // value of stencil cell (a, c, d), reference operation order: ((wx*wy)*wz)*w
        auto value = [&](int a, int c, int d) -> double { return stencil_value<ND, S, W>(wx, wy, wz, a, c, d, w); };
        constexpr int NV = ND == 1 ? NW : ND == 2 ? NW * NW : NW * NW * NW;
        // stencils of up to 16 cells are materialised so that runs can be summed across the warp;
        // the big 3-D stencils (27, 64 cells) are evaluated on the fly and never aggregated
        constexpr bool AGG = NV <= 16;
        constexpr int NA = NW, NC = ND >= 2 ? NW : 1, NDD = ND >= 3 ? NW : 1;
        double v[AGG ? NV : 1];
        bool commit = ok;
        if constexpr (AGG) {
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int d = 0; d < NDD; ++d) v[(a * NC + c) * NDD + d] = value(a, c, d);
            // ---- warp aggregation of runs of equal base cells -------------------------
            const int key = ok ? (ND == 1 ? b[0] : ND == 2 ? b[0] * P.dims[1] + b[1]
                                                           : (b[0] * P.dims[1] + b[1]) * P.dims[2] + b[2])
                               : -1 - lane;
            const int prev = __shfl_up_sync(0xffffffffu, key, 1);
            const bool head = lane == 0 || prev != key;
            const unsigned heads = __ballot_sync(0xffffffffu, head);
            if (__popc(heads) <= 16) {         // warp-uniform: runs are long enough to pay for the shuffles
                const unsigned above = lane == 31 ? 0u : heads & ~((2u << lane) - 1u);
                const int next_head = above ? __ffs(above) - 1 : 32;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    const bool take = lane + dd < next_head;
#pragma unroll
                    for (int k = 0; k < NV; ++k) {
                        const double o = __shfl_down_sync(0xffffffffu, v[k], dd);
                        if (take) v[k] = __dadd_rn(v[k], o);
                    }
                }
                commit = ok && head;
            }
        }
        auto cellval = [&](int a, int c, int d) -> double {
            if constexpr (AGG) return v[(a * NC + c) * NDD + d];
            else return value(a, c, d);
        };
        // ---- deposit: tile if the stencil fits the window, else straight to the grid ----
        bool hit = commit;
#pragma unroll
        for (int a = 0; a < ND; ++a) hit = hit && b[a] >= wd.o[a] && b[a] + NW <= wd.o[a] + wd.e[a];
        if (hit) {
            const int p1 = ND >= 2 ? wd.e[1] : 1, p2 = ND >= 3 ? wd.e[2] : 1;
            const int off = ((b[0] - wd.o[0]) * p1 + (ND >= 2 ? b[1] - wd.o[1] : 0)) * p2 + (ND >= 3 ? b[2] - wd.o[2] : 0);
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int d = 0; d < NDD; ++d) atomicAdd(mytile + off + (a * p1 + c) * p2 + d, cellval(a, c, d));
        } else if (commit) {
            const int64_t g1 = ND >= 2 ? P.dims[1] : 1, g2 = ND >= 3 ? P.dims[2] : 1;
            double *g = P.grid + ((int64_t)b[0] * g1 + b[1]) * g2 + b[2];
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int d = 0; d < NDD; ++d) red_add(g + (a * g1 + c) * g2 + d, cellval(a, c, d));
        }
        // ---- hit-rate bookkeeping and window moves -------------------------------------
        const unsigned m_ok = __ballot_sync(0xffffffffu, ok);
        unsigned inside = ok;
#pragma unroll
        for (int a = 0; a < ND; ++a) inside = inside && b[a] >= wd.o[a] && b[a] + NW <= wd.o[a] + wd.e[a];
        const unsigned m_in = __ballot_sync(0xffffffffu, inside);
        acc_w += __popc(m_ok);
        miss_w += __popc(m_ok & ~m_in);
      }
        if (++since_eval == kEvalEvery && i0 + (int64_t)TT * kTV < end) {
            since_eval = 0;
            if (threadIdx.x == 0) { s_acc = 0; s_miss = 0; }
            __syncthreads();
            if (lane == 0 && acc_w) { atomicAdd(&s_acc, acc_w); atomicAdd(&s_miss, miss_w); }
            acc_w = 0; miss_w = 0;
            __syncthreads();
            const unsigned acc = s_acc, miss = s_miss;
            const bool move = acc >= 64 && miss * 2 > acc;
            __syncthreads();
            if (move) {
                flush_tile<ND, TT, false>(tile, wd, P, cells, replicas, 0.0);
                // new centre: mean base cell of the sub-chunk in registers
                const long long cnt = block_sum(ok ? 1 : 0, scratch);
                long long mean[3] = {0, 0, 0};
                for (int a = 0; a < ND; ++a) mean[a] = block_sum(ok ? b[a] : 0, scratch) / (cnt > 0 ? cnt : 1);
                if (threadIdx.x == 0) {
                    for (int a = 0; a < ND; ++a) {
                        if (wd.e[a] >= P.dims[a]) { s_win.o[a] = 0; continue; }
                        long long o = mean[a] + NW / 2 - wd.e[a] / 2;
                        if (a == ND - 1) o &= ~1ll;
                        if (o < 0) o = 0;
                        if (o > P.dims[a] - wd.e[a]) o = P.dims[a] - wd.e[a];
                        s_win.o[a] = (int)o;
                    }
                }
                __syncthreads();
                wd = s_win;
            }
        }
    }
    __syncthreads();
    flush_tile<ND, TT, false>(tile, wd, P, cells, replicas, 0.0);
}

// ---------------------------------------------------------------------------------
// the fixed-point tile kernel (the default of method TILED for grids larger than the tile)
// ---------------------------------------------------------------------------------
// The same window scheme as k_deposit_tiled, cut down to what the native 32-bit atomics need: no warp
// aggregation of runs (a same-address collision costs one extra pass of a native atomic, not a CAS retry,
// and input with long runs is method RUNS' job), the stencil products are formed on the fly right in front
// of their two atomics, hit / miss statistics live in two registers per lane and are voted on with
// __syncthreads_count.  This keeps a 2-D TSC instance at 64 registers -- 32 warps per SM, which is what
// the shared-memory atomic pipe needs to stay busy (the general kernel with the fixed-point adds bolted on:
// 80 registers + spills at 24 warps, 72 Gp/s on x-ordered cfg2; profiles/r02/perf_fx_variants.log).
template <int NW>
__device__ __forceinline__ double pick_w(const double (&wl)[NW], int j) {
    double v = wl[0];
#pragma unroll
    for (int k = 1; k < NW; ++k) v = j == k ? wl[k] : v;
    return v;
}

// Deposit `count` (<= 8) queued particles (x, y, z, w per entry) straight into the grid with all 32 lanes: 4 lanes per
// particle, lane j of a group takes cell j of every stencil row along the last axis (the pattern of k_deposit_quads);
// the products keep the reference's operation order.
template <int ND, int S, bool W>
__device__ __forceinline__ void drain_rows(const DepositParams &P, const double *q, int count, int lane, double inv_q) {
    auto put = [&](double *cell, double v) {
        if (P.fixed_lo != nullptr) fixed_add(P, cell - P.grid, v, inv_q);
        else red_add(cell, v);
    };
    constexpr int NW = StencilWidth<S>::value;
    const int g = lane >> 2, j = lane & 3;
    if (g >= count || j >= NW) return;
    const double *e = q + g * 4;
    int b[3];
    double wx[NW], wy[NW], wz[NW];
    if (!particle_stencil<ND, S>(P, e[0], e[1], e[2], b, wx, wy, wz)) return;      // (accepted when it was queued)
    const double w = e[3];
    const int64_t g1 = ND >= 2 ? P.dims[1] : 1, g2 = ND >= 3 ? P.dims[2] : 1;
    double *cell = P.grid + ((int64_t)b[0] * g1 + b[1]) * g2 + b[2] + j;
    if (ND == 1) {
        put(cell, W ? __dmul_rn(pick_w<NW>(wx, j), w) : pick_w<NW>(wx, j));
    } else if (ND == 2) {
        const double wl = pick_w<NW>(wy, j);
#pragma unroll
        for (int a = 0; a < NW; ++a) {
            const double c = __dmul_rn(wx[a], wl);
            put(cell + a * g1, W ? __dmul_rn(c, w) : c);
        }
    } else {
        const double wl = pick_w<NW>(wz, j);
#pragma unroll
        for (int a = 0; a < NW; ++a)
#pragma unroll
            for (int c2 = 0; c2 < NW; ++c2) {
                const double c = __dmul_rn(__dmul_rn(wx[a], wy[c2]), wl);
                put(cell + (a * g1 + c2) * g2, W ? __dmul_rn(c, w) : c);
            }
    }
}

#ifndef PTG_FX_THREADS_9
#define PTG_FX_THREADS_9 1024
#endif
template <int ND, int S> struct FxThreads {
    static constexpr int NW = StencilWidth<S>::value;
    static constexpr int NV = ND == 1 ? NW : ND == 2 ? NW * NW : NW * NW * NW;
    static constexpr int value = NV <= 4 ? 1024 : NV <= 9 ? PTG_FX_THREADS_9 : NV <= 16 ? 768 : 512;
};

template <int ND, int S, bool W, bool F>
__global__ void __launch_bounds__((FxThreads<ND, S>::value), 1) k_deposit_tile_fx(const DepositParams P, const TileConfig *cfgp) {
    constexpr int TT = FxThreads<ND, S>::value;
    // sub-chunks between two evaluations: small stencils leave room in the 64-bit range for twice the interval
    constexpr int EV = FxThreads<ND, S>::NV <= 9 ? 2 * kEvalEvery : kEvalEvery;
    extern __shared__ double tile[];
    __shared__ long long scratch[34];
    __shared__ Window s_win;
    constexpr int NW = StencilWidth<S>::value;
    constexpr int NA = NW, NC = ND >= 2 ? NW : 1, NDD = ND >= 3 ? NW : 1;

    if (cfgp->mode != kModeTiled) return;
    const double fx_inv_q = cfgp->fx_inv_q, fx_q = cfgp->fx_q, fx_wcap = cfgp->fx_wcap;
    const double fx_limit = fx_lane_limit(TT, EV * kTV) * fx_wcap;
    double fx_sum = 0.0;
    unsigned *const lo = reinterpret_cast<unsigned *>(tile);
    // Particles that miss the window (the thin background around a bunch) are not deposited by the lane that
    // found them -- 1 lane in 10 would drag its whole warp through (S+1)^ndim uncoalesced REDs -- but parked in a
    // per-warp queue behind the tile and deposited 8 at a time by all 32 lanes, 4 lanes per particle, each lane
    // one cell of a stencil row: the sector-coalesced pattern of method ROWS (3 instead of 9 L2 sector
    // reductions per 2-D TSC particle).
    constexpr bool QUEUE = S >= 1;
    constexpr int kQCap = 16;              // entries per warp (at most 7 left over + 8 new)
    double *const wq = tile + kTileCells + (threadIdx.x >> 5) * (kQCap * 4);
    int qcount = 0;                        // warp-uniform
    const int lane = threadIdx.x & 31;
    int cells = 1;
    for (int a = 0; a < ND; ++a) cells *= cfgp->ext[a];
    for (int c = threadIdx.x; c < cells; c += TT) { lo[c] = 0u; lo[kFxHi + c] = 0u; }
    if (threadIdx.x == 0) {
        for (int a = 0; a < 3; ++a) { s_win.e[a] = a < ND ? cfgp->ext[a] : 1; s_win.o[a] = 0; }
        // a window that covers the whole axis never needs to move; the others start "nowhere"
        for (int a = 0; a < ND; ++a) if (cfgp->ext[a] < P.dims[a]) s_win.o[a] = -(1 << 28);
    }
    __syncthreads();
    Window wd = s_win;

    int64_t per = (P.n + gridDim.x - 1) / gridDim.x;
    per = (per + TT * kTV - 1) / (TT * kTV) * (TT * kTV);
    const int64_t begin = (int64_t)blockIdx.x * per;
    const int64_t end = begin + per < P.n ? begin + per : P.n;
    const double *px = P.p[0], *py = P.p[1], *pz = P.p[2];
    const int64_t sx = P.stride[0], sy = P.stride[1], sz = P.stride[2];
    bool vec = sx == 1 && (reinterpret_cast<uintptr_t>(px) & 31u) == 0;
    if (ND >= 2) vec = vec && sy == 1 && (reinterpret_cast<uintptr_t>(py) & 31u) == 0;
    if (ND >= 3) vec = vec && sz == 1 && (reinterpret_cast<uintptr_t>(pz) & 31u) == 0;
    if (W && !F) vec = vec && P.stride[3] == 1 && (reinterpret_cast<uintptr_t>(P.p[3]) & 31u) == 0;
    if (F && P.any_akind) vec = false;      // axis expressions are evaluated particle by particle
    const int64_t g1 = ND >= 2 ? P.dims[1] : 1, g2 = ND >= 3 ? P.dims[2] : 1;

    int since_eval = EV - 1;       // first evaluation right after the first sub-chunk
    unsigned n_ok = 0, n_miss = 0;         // this lane's particles since the last evaluation
    // (A lane-strided order -- the lanes of a warp thousands of particles apart, so that cell-ordered input does not put
    // adjacent lanes on one address -- was measured and dropped: cell-ordered input stays at 16 Gp/s either way, because
    // all warps of the CTA still update the same few cells and same-address shared atomics do not pipeline; such input
    // is method RUNS' job.  profiles/r02/perf_fx_lanestride.log)
    for (int64_t i0 = begin; i0 < end; i0 += (int64_t)TT * kTV) {
        const int64_t ib = i0 + (int64_t)threadIdx.x * kTV;
        const bool more = i0 + (int64_t)TT * kTV < end;
        double xv[kTV], yv[kTV], zv[kTV], wv[kTV];
        if (vec && ib + kTV <= end) {
            tile_ld4(px + ib, xv);
            if (ND >= 2) tile_ld4(py + ib, yv);
            if (ND >= 3) tile_ld4(pz + ib, zv);
            if (W && !F) tile_ld4(P.p[3] + ib, wv);
        } else {
#pragma unroll
            for (int j = 0; j < kTV; ++j) {
                const bool in = ib + j < end;
                xv[j] = in ? particle_coord<F>(P, 0, ib + j) : 0.0;
                if (ND >= 2) yv[j] = in ? particle_coord<F>(P, 1, ib + j) : 0.0;
                if (ND >= 3) zv[j] = in ? particle_coord<F>(P, 2, ib + j) : 0.0;
                if (W && !F) wv[j] = in ? ld_stream(P.p[3] + (ib + j) * P.stride[3]) : 0.0;
            }
        }
        if constexpr (W && F) {
#pragma unroll
            for (int j = 0; j < kTV; ++j) wv[j] = ib + j < end ? particle_weight<F>(P, ib + j) : 0.0;
        }
        bool ok = false;                   // (of the lane's last particle: seeds the window after a move)
        int b[3] = {0, 0, 0};
#pragma unroll
        for (int j = 0; j < kTV; ++j) {
            const int64_t i = ib + j;
            const double w = W ? wv[j] : 1.0;
            double wx[NW], wy[NW], wz[NW];
            ok = particle_stencil<ND, S>(P, xv[j], ND >= 2 ? yv[j] : 0.0, ND >= 3 ? zv[j] : 0.0, b, wx, wy, wz) &&
                 i < end && particle_selected(P, i);
            bool direct = false;           // this particle goes to the grid exactly (misses the window / too heavy)
            if (ok) {
            bool inside = true;
#pragma unroll
            for (int a = 0; a < ND; ++a) inside = inside && b[a] >= wd.o[a] && b[a] + NW <= wd.o[a] + wd.e[a];
            ++n_ok;
            // the tile takes weights up to the cap (NaN / inf fail the comparison); the rest goes to the grid exactly
            if (inside && fx_inv_q != 0.0 && (!W || fabs(w) <= fx_wcap)) {
                const int p1 = ND >= 2 ? wd.e[1] : 1, p2 = ND >= 3 ? wd.e[2] : 1;
                const int off = ((b[0] - wd.o[0]) * p1 + (ND >= 2 ? b[1] - wd.o[1] : 0)) * p2 + (ND >= 3 ? b[2] - wd.o[2] : 0);
#pragma unroll
                for (int a = 0; a < NA; ++a)
#pragma unroll
                    for (int c = 0; c < NC; ++c)
#pragma unroll
                        for (int d = 0; d < NDD; ++d)
                            fx_add(lo, off + (a * p1 + c) * p2 + d, stencil_value<ND, S, W>(wx, wy, wz, a, c, d, w), fx_inv_q);
                fx_sum = __dadd_rn(fx_sum, W ? fabs(w) : 1.0);
            } else {
                if (!inside) ++n_miss;
                direct = true;
            }
            }
            // ---- particles for the exact path: queue them (few) or deposit them right here (many) ----
            const unsigned dm = __ballot_sync(0xffffffffu, direct);
            if (dm == 0u) continue;
            if (QUEUE && __popc(dm) <= 8) {
                if (direct) {
                    double *e = wq + (qcount + __popc(dm & ((1u << lane) - 1u))) * 4;
                    e[0] = xv[j];
                    e[1] = ND >= 2 ? yv[j] : 0.0;
                    e[2] = ND >= 3 ? zv[j] : 0.0;
                    e[3] = w;
                }
                qcount += __popc(dm);
                __syncwarp();
                if (qcount >= 8) {
                    qcount -= 8;
                    drain_rows<ND, S, W>(P, wq + qcount * 4, 8, lane, fx_inv_q);
                    __syncwarp();
                }
            } else if (direct) {
                const int64_t lin0 = ((int64_t)b[0] * g1 + b[1]) * g2 + b[2];
#pragma unroll
                for (int a = 0; a < NA; ++a)
#pragma unroll
                    for (int c = 0; c < NC; ++c)
#pragma unroll
                        for (int d = 0; d < NDD; ++d) {
                            const double v = stencil_value<ND, S, W>(wx, wy, wz, a, c, d, w);
                            if (P.fixed_lo != nullptr) fixed_add(P, lin0 + (a * g1 + c) * g2 + d, v, fx_inv_q);
                            else red_add(P.grid + lin0 + (a * g1 + c) * g2 + d, v);
                        }
            }
        }
        if (++since_eval == EV && more) {
            since_eval = 0;
            // lanes whose particles mostly missed the window vote for a move; any lane near its share of the
            // tile's 64-bit range asks for a merge
            const int votes = __syncthreads_count(n_miss * 2 > n_ok);
            const int full = __syncthreads_or(fx_sum > 0.0 && fx_sum >= fx_limit);
            n_ok = 0; n_miss = 0;
            const bool move = votes * 2 > TT;
            if (move || full) {
                fx_sum = 0.0;
                flush_tile<ND, TT, true>(tile, wd, P, cells, 1, fx_q);
                __syncthreads();
            }
            if (move) {
                // new centre: mean base cell of the particles in registers
                const long long cnt = block_sum(ok ? 1 : 0, scratch);
                long long mean[3] = {0, 0, 0};
                for (int a = 0; a < ND; ++a) mean[a] = block_sum(ok ? b[a] : 0, scratch) / (cnt > 0 ? cnt : 1);
                if (threadIdx.x == 0) {
                    for (int a = 0; a < ND; ++a) {
                        if (wd.e[a] >= P.dims[a]) { s_win.o[a] = 0; continue; }
                        long long o = mean[a] + NW / 2 - wd.e[a] / 2;
                        if (a == ND - 1) o &= ~1ll;
                        if (o < 0) o = 0;
                        if (o > P.dims[a] - wd.e[a]) o = P.dims[a] - wd.e[a];
                        s_win.o[a] = (int)o;
                    }
                }
                __syncthreads();
                wd = s_win;
            }
        }
    }
    if (QUEUE && qcount > 0) drain_rows<ND, S, W>(P, wq, qcount, lane, fx_inv_q);
    __syncthreads();
    flush_tile<ND, TT, true>(tile, wd, P, cells, 1, fx_q);
}

// ---------------------------------------------------------------------------------
// whole grid in shared memory (1-D spectra, small 2-D / 3-D histograms)
// ---------------------------------------------------------------------------------
// When the padded grid itself fits the tile there is no window to manage: a lean kernel keeps R
// replicas of the grid in shared memory (warp w adds to replica w % R, dividing same-address
// contention), streams the particles with kSU independent loads in flight per thread, and merges the
// replicas into the global grid at the end.  This is the path of postpic's 1-D energy spectra.
constexpr int kST = 1024;     // threads per CTA, one CTA per SM
constexpr int kSU = 4;        // particles in flight per thread
constexpr int kSCheck = 8;    // fixed point: main-loop iterations between two range checks

template <int ND, int S, bool W, bool F, bool FX>
__global__ void __launch_bounds__(kST, 1) k_deposit_small(const DepositParams P, const TileConfig *gate) {
    if (gate->mode != kModeTiled) return;
    extern __shared__ double tile[];
    constexpr int NW = StencilWidth<S>::value;
    constexpr int NA = NW, NC = ND >= 2 ? NW : 1, NDD = ND >= 3 ? NW : 1;
    // Plain counting (shape 0, no weights -- numpy.histogram's job): every contribution is exactly 1,
    // so the tile holds 32-bit counters and uses one native shared-memory atomic per particle
    // (ATOMS.ADD.U32); the counts are exact and are converted to
    // double at the merge.  A CTA sees fewer than 2^32 particles for any n < 148 * 2^32.
    constexpr bool COUNT = (S == 0 && !W);
    constexpr bool FXA = FX && !COUNT;          // 64-bit fixed point for everything else
    const int cells = P.dims[0] * P.dims[1] * P.dims[2];
    int replicas = kTileCells / cells;
    if (replicas > kST / 32) replicas = kST / 32;
    if (replicas < 1) replicas = 1;
    unsigned *ctile = reinterpret_cast<unsigned *>(tile);
    if (COUNT) { for (int c = threadIdx.x; c < cells * replicas; c += kST) ctile[c] = 0u; }
    else if (FXA) { for (int c = threadIdx.x; c < cells * replicas; c += kST) { ctile[c] = 0u; ctile[kFxHi + c] = 0u; } }
    else       { for (int c = threadIdx.x; c < cells * replicas; c += kST) tile[c] = 0.0; }
    __syncthreads();
    double *mine = tile + ((threadIdx.x >> 5) % replicas) * cells;
    unsigned *cmine = ctile + ((threadIdx.x >> 5) % replicas) * cells;
    const int p1 = ND >= 2 ? P.dims[1] : 1, p2 = ND >= 3 ? P.dims[2] : 1;
    double fx_inv_q = 0.0, fx_q = 0.0, fx_wcap = 0.0, fx_sum = 0.0;
    if (FXA) { fx_inv_q = gate->fx_inv_q; fx_q = gate->fx_q; fx_wcap = gate->fx_wcap; }
    const double fx_limit = fx_lane_limit(kST, kSCheck * kSU) * fx_wcap;

    auto one = [&](double x, double y, double z, double w) {
        int b[3];
        double wx[NW], wy[NW], wz[NW];
        const bool ok = particle_stencil<ND, S>(P, x, y, z, b, wx, wy, wz);
        if (!ok) return;
        if constexpr (COUNT) {
            atomicAdd(cmine + (b[0] * p1 + b[1]) * p2 + b[2], 1u);
            return;
        }
        const int off = (b[0] * p1 + b[1]) * p2 + b[2];
        if constexpr (FXA) {
            if (!(fx_inv_q != 0.0 && (!W || fabs(w) <= fx_wcap))) {
                // outside the fixed-point range (or NaN / inf): exact reductions straight into the grid
#pragma unroll
                for (int a = 0; a < NA; ++a)
#pragma unroll
                    for (int c = 0; c < NC; ++c)
#pragma unroll
                        for (int d = 0; d < NDD; ++d) {
                            const double v = stencil_value<ND, S, W>(wx, wy, wz, a, c, d, w);
                            if (P.fixed_lo != nullptr) fixed_add(P, off + (a * p1 + c) * p2 + d, v, fx_inv_q);
                            else red_add(P.grid + off + (a * p1 + c) * p2 + d, v);
                        }
                return;
            }
            fx_sum = __dadd_rn(fx_sum, W ? fabs(w) : 1.0);
        }
#pragma unroll
        for (int a = 0; a < NA; ++a)
#pragma unroll
            for (int c = 0; c < NC; ++c)
#pragma unroll
                for (int d = 0; d < NDD; ++d) {
                    const double v = stencil_value<ND, S, W>(wx, wy, wz, a, c, d, w);
                    if constexpr (FXA) fx_add(cmine, off + (a * p1 + c) * p2 + d, v, fx_inv_q);
                    else atomicAdd(mine + off + (a * p1 + c) * p2 + d, v);
                }
    };
    // merge all replicas into the grid (and clear them: the fixed-point tile is merged whenever its range runs out)
    auto merge = [&]() {
        for (int c = threadIdx.x; c < cells; c += kST) {
            double v;
            if (COUNT) {
                unsigned long long k = ctile[c];
                for (int r = 1; r < replicas; ++r) k += ctile[r * cells + c];
                v = (double)k;
            } else if (FXA) {
                long long k = fx_take(ctile, c);
                for (int r = 1; r < replicas; ++r) k += fx_take(ctile, r * cells + c);
                if (P.fixed_lo != nullptr) {             // method FIXED: stay integer
                    if (k != 0) fixed_accumulate(P, c, k);
                    continue;
                }
                v = __dmul_rn(__ll2double_rn(k), fx_q);
            } else {
                v = tile[c];
                for (int r = 1; r < replicas; ++r) v = __dadd_rn(v, tile[r * cells + c]);
            }
            if (v != 0.0) red_add(P.grid + c, v);
        }
    };

    const int64_t nthreads = (int64_t)gridDim.x * kST;
    const int64_t base = (int64_t)blockIdx.x * kST + threadIdx.x;
    const double *px = P.p[0], *py = P.p[1], *pz = P.p[2];
    const int64_t sx = P.stride[0], sy = P.stride[1], sz = P.stride[2];
    // the same number of full iterations for every thread of the grid (the range checks below are CTA-wide)
    const int64_t full = P.n / (kSU * nthreads);
    for (int64_t it = 0; it < full; ++it) {
        const int64_t i = base + it * kSU * nthreads;
        double x[kSU], y[kSU], z[kSU], w[kSU];
#pragma unroll
        for (int u = 0; u < kSU; ++u) {
            const int64_t k = i + u * nthreads;
            x[u] = particle_coord<F>(P, 0, k);
            y[u] = ND >= 2 ? particle_coord<F>(P, 1, k) : 0.0;
            z[u] = ND >= 3 ? particle_coord<F>(P, 2, k) : 0.0;
            w[u] = W ? particle_weight<F>(P, k) : 1.0;
        }
#pragma unroll
        for (int u = 0; u < kSU; ++u)
            if (particle_selected(P, i + u * nthreads)) one(x[u], y[u], z[u], w[u]);
        if (FXA && (it % kSCheck) == kSCheck - 1) {
            if (__syncthreads_or(fx_sum > 0.0 && fx_sum >= fx_limit)) {
                merge();
                fx_sum = 0.0;
                __syncthreads();
            }
        }
    }
    // at most kSU particles per lane are left
    for (int64_t i = base + full * kSU * nthreads; i < P.n; i += nthreads)
        if (particle_selected(P, i)) one(particle_coord<F>(P, 0, i), ND >= 2 ? particle_coord<F>(P, 1, i) : 0.0,
                                         ND >= 3 ? particle_coord<F>(P, 2, i) : 0.0, W ? particle_weight<F>(P, i) : 1.0);
    __syncthreads();
    merge();
}

// ---------------------------------------------------------------------------------
// small calls in ONE launch (postpic's 1-D spectra of 1e5..1e6 particles: BASELINE cfg1 as named)
// ---------------------------------------------------------------------------------
// A call of a million particles is 2 us of HBM time; memset + probe + deposit kernel + their launch gaps made it
// 30 us.  This kernel needs nothing before or after it: every CTA finds the largest |weight| of ITS particles (the
// quantum of its own fixed-point tile), deposits them into shared memory, writes its tile to a scratch slab as doubles,
// and the last CTA to finish (ticket) sums the slabs in CTA order and STORES the grid (or adds to it when the caller
// did not ask for a cleared grid) -- no memset, no probe, no global atomics, and a result that is bit-reproducible.
constexpr int kOneShotCells = 4096;
constexpr int kOT = 1024;

template <int ND, int S, bool W, bool F>
__global__ void __launch_bounds__(kOT, 1) k_deposit_oneshot(const DepositParams P, double *scratch, unsigned *ticket,
                                                            int zero_first) {
    extern __shared__ double tile[];
    __shared__ long long red_scratch[34];
    __shared__ bool s_last;
    constexpr int NW = StencilWidth<S>::value;
    constexpr int NA = NW, NC = ND >= 2 ? NW : 1, NDD = ND >= 3 ? NW : 1;
    const int cells = P.dims[0] * P.dims[1] * P.dims[2];
    int replicas = kOneShotCells / cells;
    replicas = replicas > 8 ? 8 : (replicas < 1 ? 1 : replicas);
    unsigned *lo = reinterpret_cast<unsigned *>(tile);             // [replicas * cells] low words, then high words
    unsigned *hi = lo + kOneShotCells;
    double *extra = tile + kOneShotCells;                          // NaN / inf contributions, fp64
    for (int c = threadIdx.x; c < kOneShotCells; c += kOT) { lo[c] = 0u; hi[c] = 0u; extra[c] = 0.0; }
    const int64_t step = (int64_t)gridDim.x * kOT;
    const int64_t first = (int64_t)blockIdx.x * kOT + threadIdx.x;
    // pass 1: this CTA's largest finite |weight|  (the particles are re-read from L2 / L1 in pass 2)
    double wmax = W ? 0.0 : 1.0;
    if (W) {
        for (int64_t i = first; i < P.n; i += step) {
            if (!particle_selected(P, i)) continue;
            const double a = fabs(particle_weight<F>(P, i));
            if (a <= 1.79769313486231570e308) wmax = fmax(wmax, a);
        }
        wmax = block_max(wmax, red_scratch);
    } else {
        __syncthreads();
    }
    int e = wmax > 0.0 ? ilogb(wmax) + 1 : 0;
    e = e < -960 ? -960 : (e > 960 ? 960 : e);
    // a CTA sees at most 2^14 particles (the launcher guarantees it): |contribution| <= 4 |w| < 2^(e+2) leaves room for a
    // quantum of 2^(e-47) in the 64-bit cells -- 4x finer than the windowed tile's
    const double q = scalbn(1.0, e - 47), inv_q = scalbn(1.0, 47 - e);
    unsigned *mylo = lo + ((threadIdx.x >> 5) % replicas) * cells;
    const int p1 = ND >= 2 ? P.dims[1] : 1, p2 = ND >= 3 ? P.dims[2] : 1;
    // pass 2: deposit
    for (int64_t i = first; i < P.n; i += step) {
        if (!particle_selected(P, i)) continue;
        const double x = particle_coord<F>(P, 0, i);
        const double y = ND >= 2 ? particle_coord<F>(P, 1, i) : 0.0;
        const double z = ND >= 3 ? particle_coord<F>(P, 2, i) : 0.0;
        const double w = W ? particle_weight<F>(P, i) : 1.0;
        int b[3];
        double wx[NW], wy[NW], wz[NW];
        if (!particle_stencil<ND, S>(P, x, y, z, b, wx, wy, wz)) continue;
        const int off = (b[0] * p1 + b[1]) * p2 + b[2];
        const bool finite = !W || fabs(w) <= 1.79769313486231570e308;
#pragma unroll
        for (int a = 0; a < NA; ++a)
#pragma unroll
            for (int c = 0; c < NC; ++c)
#pragma unroll
                for (int d = 0; d < NDD; ++d) {
                    const double v = stencil_value<ND, S, W>(wx, wy, wz, a, c, d, w);
                    const int cell = off + (a * p1 + c) * p2 + d;
                    if (finite) {
                        const long long k = __double2ll_rn(__dmul_rn(v, inv_q));
                        const unsigned l = (unsigned)k, h = (unsigned)((unsigned long long)k >> 32);
                        const unsigned old = atomicAdd(mylo + cell, l);
                        atomicAdd(mylo + kOneShotCells + cell, h + ((old + l) < old ? 1u : 0u));
                    } else {
                        atomicAdd(extra + cell, v);
                    }
                }
    }
    __syncthreads();
    // this CTA's tile -> its scratch slab, as doubles
    double *slab = scratch + (size_t)blockIdx.x * cells;
    for (int c = threadIdx.x; c < cells; c += kOT) {
        long long k = 0;
        for (int r = 0; r < replicas; ++r)
            k += (long long)(((unsigned long long)hi[r * cells + c] << 32) | lo[r * cells + c]);
        slab[c] = __dadd_rn(__dmul_rn(__ll2double_rn(k), q), extra[c]);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int c = threadIdx.x; c < cells; c += kOT) {
        // slabs are summed in a fixed order (eight interleaved partial sums, then those): reproducible, and eight loads
        // in flight per thread instead of a chain of 148 dependent L2 round trips
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        unsigned blk = 0;
        for (; blk + 8 <= gridDim.x; blk += 8) {
            double t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] = __ldcg(scratch + (size_t)(blk + k) * cells + c);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = __dadd_rn(acc[k], t[k]);
        }
        for (; blk < gridDim.x; ++blk) acc[blk & 7] = __dadd_rn(acc[blk & 7], __ldcg(scratch + (size_t)blk * cells + c));
        const double v = __dadd_rn(__dadd_rn(__dadd_rn(acc[0], acc[1]), __dadd_rn(acc[2], acc[3])),
                                   __dadd_rn(__dadd_rn(acc[4], acc[5]), __dadd_rn(acc[6], acc[7])));
        if (zero_first) P.grid[c] = v;
        else if (v != 0.0) red_add(P.grid + c, v);      // other streams (the host pipeline's chunks) may be adding too
    }
    if (threadIdx.x == 0) *ticket = 0u;               // the ring entry is ready for its next user
}

template <int ND, int S, bool W>
static cudaError_t launch_oneshot(const DepositParams &P, TileConfig *cfg, bool zero_first, int sm_count, cudaStream_t st) {
    const int cells = P.dims[0] * P.dims[1] * P.dims[2];
    int64_t want = (P.n + 2047) / 2048;
    const int blocks = (int)(want < sm_count ? (want < 1 ? 1 : want) : sm_count);
    void *scratch = nullptr;
    cudaError_t e = scratch_alloc(&scratch, (size_t)blocks * cells * sizeof(double), st);
    if (e != cudaSuccess) return e;
    const size_t smem = (size_t)kOneShotCells * 2 * sizeof(double);
    const bool fuse = W && fused_mode(P);
    auto kern = fuse ? k_deposit_oneshot<ND, S, W, W> : k_deposit_oneshot<ND, S, W, false>;
    static bool opted_in[2][64] = {};                        // per device and instance: > 48 KB of dynamic shared memory
    int dev = 0;
    e = cudaGetDevice(&dev);
    if (e == cudaSuccess && (dev < 0 || dev >= 64 || !opted_in[fuse ? 1 : 0][dev])) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess && dev >= 0 && dev < 64) opted_in[fuse ? 1 : 0][dev] = true;
    }
    if (e == cudaSuccess) {
        kern<<<blocks, kOT, smem, st>>>(P, static_cast<double *>(scratch), &cfg->ticket, zero_first ? 1 : 0);
        e = cudaGetLastError();
    }
    cudaError_t e2 = cudaFreeAsync(scratch, st);
    return e != cudaSuccess ? e : e2;
}

template <int ND>
static cudaError_t oneshot_s(int shape, const DepositParams &P, TileConfig *cfg, bool zf, int sm, cudaStream_t st) {
    const bool w = P.p[3] || fused_mode(P);
    switch (shape) {
        case 0: return w ? launch_oneshot<ND, 0, true>(P, cfg, zf, sm, st) : launch_oneshot<ND, 0, false>(P, cfg, zf, sm, st);
        case 1: return w ? launch_oneshot<ND, 1, true>(P, cfg, zf, sm, st) : launch_oneshot<ND, 1, false>(P, cfg, zf, sm, st);
        case 2: return w ? launch_oneshot<ND, 2, true>(P, cfg, zf, sm, st) : launch_oneshot<ND, 2, false>(P, cfg, zf, sm, st);
        default: return w ? launch_oneshot<ND, 3, true>(P, cfg, zf, sm, st) : launch_oneshot<ND, 3, false>(P, cfg, zf, sm, st);
    }
}

// (calls of a few thousand particles keep the path without a sample, which adds every contribution in fp64: a single
// particle is then deposited bit for bit like the reference does)
#define PTG_CAT2(a, b) a##b
#define PTG_CAT(a, b) PTG_CAT2(a, b)
// One translation unit per number of axes: this file is compiled four times (postpic_b200/build.py) -- as itself it
// holds the dispatchers only, ptg_deposit_tiled_nd{1,2,3}.cu include it with PTG_ND set and instantiate the kernels
// of that dimensionality (the build is bound by its longest translation unit).
#ifdef PTG_ND
cudaError_t PTG_CAT(deposit_oneshot_nd, PTG_ND)(int shape, const DepositParams &P, TileConfig *cfg, bool zero_first, int sm_count,
                                                cudaStream_t st) {
    return oneshot_s<PTG_ND>(shape, P, cfg, zero_first, sm_count, st);
}
#else
cudaError_t deposit_oneshot_nd1(int, const DepositParams &, TileConfig *, bool, int, cudaStream_t);
cudaError_t deposit_oneshot_nd2(int, const DepositParams &, TileConfig *, bool, int, cudaStream_t);
cudaError_t deposit_oneshot_nd3(int, const DepositParams &, TileConfig *, bool, int, cudaStream_t);

bool oneshot_applies(const DepositParams &P, int sm_count) {
    return P.n >= 4096 && P.n < (int64_t(1) << 21) && P.n <= (int64_t)sm_count * 16384 &&
           (int64_t)P.dims[0] * P.dims[1] * P.dims[2] <= kOneShotCells;
}

cudaError_t deposit_oneshot(int ndim, int shape, const DepositParams &P, TileConfig *cfg, bool zero_first, int sm_count,
                            cudaStream_t st) {
    switch (ndim) {
        case 1: return deposit_oneshot_nd1(shape, P, cfg, zero_first, sm_count, st);
        case 2: return deposit_oneshot_nd2(shape, P, cfg, zero_first, sm_count, st);
        default: return deposit_oneshot_nd3(shape, P, cfg, zero_first, sm_count, st);
    }
}
#endif

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
template <int ND, int S, bool W, bool FX>
static cudaError_t launch_tiled(const DepositParams &P, TileConfig *cfg, int force, int sm_count, cudaStream_t st,
                                const FixedState *fixed, cudaEvent_t after_probe) {
    constexpr int NW = StencilWidth<S>::value;
    constexpr int TT = TileThreads<(ND == 1 ? NW : ND == 2 ? NW * NW : NW * NW * NW)>::value;
    const bool fuse = W && fused_mode(P);
    auto kern = fuse ? k_deposit_tiled<ND, S, W, TT, W> : k_deposit_tiled<ND, S, W, TT, false>;
    const size_t smem = (size_t)kTileCells * sizeof(double);
    // per-device attribute, cheap to set: done on every launch so that any device works
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_probe<ND, S><<<kProbeBlocks, kTT, 0, st>>>(P, cfg, force, runs_wide_available(ND, S) ? 1 : 0, FX ? 0 : 1);
    if (fixed != nullptr) fixed_set_scale(cfg, fixed, st);      // method FIXED: the quantum comes from the exact max |w|
    if (after_probe != nullptr) {
        e = cudaEventRecord(after_probe, st);
        if (e != cudaSuccess) return e;
    }
    if ((int64_t)P.dims[0] * P.dims[1] * P.dims[2] <= kTileCells) {
        // the whole padded grid fits the tile: lean kernel without window management
        auto small = fuse ? k_deposit_small<ND, S, W, W, FX> : k_deposit_small<ND, S, W, false, FX>;
        e = cudaFuncSetAttribute(small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int64_t want = (P.n + (int64_t)kST * kSU - 1) / ((int64_t)kST * kSU);
        small<<<(int)(want < sm_count ? (want < 1 ? 1 : want) : sm_count), kST, smem, st>>>(P, cfg);
        return cudaGetLastError();
    }
    if constexpr (FX) {
        auto lean = fuse ? k_deposit_tile_fx<ND, S, W, W> : k_deposit_tile_fx<ND, S, W, false>;
        const size_t smem_fx = smem + (size_t)(FxThreads<ND, S>::value / 32) * 16 * 4 * sizeof(double);   // + the miss queues
        e = cudaFuncSetAttribute(lean, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fx);
        if (e != cudaSuccess) return e;
        lean<<<sm_count, FxThreads<ND, S>::value, smem_fx, st>>>(P, cfg);
        return cudaGetLastError();
    }
    kern<<<sm_count, TT, smem, st>>>(P, cfg);
    return cudaGetLastError();
}

template <int ND, int S>
static cudaError_t tiled_w(const DepositParams &P, TileConfig *cfg, int force, bool exact, int sm, cudaStream_t st,
                           const FixedState *fixed, cudaEvent_t ev) {
    const bool w = P.p[3] || fused_mode(P);
    if (exact) return w ? launch_tiled<ND, S, true, false>(P, cfg, force, sm, st, nullptr, ev) : launch_tiled<ND, S, false, false>(P, cfg, force, sm, st, nullptr, ev);
    return w ? launch_tiled<ND, S, true, true>(P, cfg, force, sm, st, fixed, ev) : launch_tiled<ND, S, false, true>(P, cfg, force, sm, st, fixed, ev);
}

template <int ND>
static cudaError_t tiled_s(int shape, const DepositParams &P, TileConfig *cfg, int force, bool exact, int sm, cudaStream_t st,
                           const FixedState *fixed, cudaEvent_t ev) {
    switch (shape) {
        case 0: return tiled_w<ND, 0>(P, cfg, force, exact, sm, st, fixed, ev);
        case 1: return tiled_w<ND, 1>(P, cfg, force, exact, sm, st, fixed, ev);
        case 2: return tiled_w<ND, 2>(P, cfg, force, exact, sm, st, fixed, ev);
        default: return tiled_w<ND, 3>(P, cfg, force, exact, sm, st, fixed, ev);
    }
}

#ifdef PTG_ND
cudaError_t PTG_CAT(deposit_tiled_nd, PTG_ND)(int shape, const DepositParams &P, TileConfig *cfg, bool force, bool exact,
                                              int sm_count, cudaStream_t st, const FixedState *fixed, cudaEvent_t after_probe) {
    return tiled_s<PTG_ND>(shape, P, cfg, force, exact, sm_count, st, fixed, after_probe);
}
#else
cudaError_t deposit_tiled_nd1(int, const DepositParams &, TileConfig *, bool, bool, int, cudaStream_t, const FixedState *, cudaEvent_t);
cudaError_t deposit_tiled_nd2(int, const DepositParams &, TileConfig *, bool, bool, int, cudaStream_t, const FixedState *, cudaEvent_t);
cudaError_t deposit_tiled_nd3(int, const DepositParams &, TileConfig *, bool, bool, int, cudaStream_t, const FixedState *, cudaEvent_t);

cudaError_t deposit_tiled(int ndim, int shape, const DepositParams &P, TileConfig *cfg, bool force, bool exact,
                          int sm_count, cudaStream_t st, const FixedState *fixed, cudaEvent_t after_probe) {
    if (P.n <= 0) return cudaSuccess;
    switch (ndim) {
        case 1: return deposit_tiled_nd1(shape, P, cfg, force, exact, sm_count, st, fixed, after_probe);
        case 2: return deposit_tiled_nd2(shape, P, cfg, force, exact, sm_count, st, fixed, after_probe);
        default: return deposit_tiled_nd3(shape, P, cfg, force, exact, sm_count, st, fixed, after_probe);
    }
}
#endif

}  // namespace ptg
