// ptg_deposit_tiled.cu -- method TILED: shared-memory privatised sub-grid tiles.
//
// One persistent CTA per SM walks a contiguous span of the particle arrays and keeps a
// window ("tile") of the ghost-padded grid in its 200 KB of shared memory:
//   * particles whose whole stencil falls inside the window are accumulated there
//     with shared-memory atomics; runs of adjacent lanes that hit the same base cell
//     (cell-ordered PIC output, tight bunches) are first summed across the warp with
//     shuffles so that one lane commits the run (warp-aggregated atomics);
//   * particles outside the window go straight to the grid in HBM/L2 with RED.F64,
//     so the result is correct for any input -- the window only decides the speed;
//   * the CTA watches its hit rate and, when most particles miss, merges the tile
//     into the global grid and re-centres the window on the particles it is
//     currently reading; at the end of its span the tile is merged once more.
// The window *shape* is chosen per launch by a one-CTA probe kernel that looks at 16
// blocks of 1024 consecutive particles: extents ~ 4 x the mean absolute deviation of
// the stencil base cell inside a block, scaled to the tile budget.  For inputs without
// locality (i.i.d. positions over a grid much larger than the tile) the probe marks
// the tile as useless and the launcher's second kernel (method ROWS) does the work.
//
// Replaces the loops of /root/reference/postpic/particles/_particlestogrid.pyx
// :70-129 (1-D), :181-276 (2-D), :341-463 (3-D).
#include "ptg_internal.cuh"

namespace ptg {

constexpr int kTT = 512;                  // probe CTA width
// CTA width of the tiled kernel by stencil size: small stencils need few registers, so more warps
// can keep shared-memory atomics in flight (one CTA per SM owns the whole 200 KB tile)
template <int NV> struct TileThreads { static constexpr int value = NV <= 4 ? 1024 : NV <= 9 ? 768 : 512; };
constexpr int kEvalEvery = 4;             // sub-chunks between hit-rate evaluations

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
__global__ void __launch_bounds__(kTT) k_probe(const DepositParams P, TileConfig *cfg, int force, int wide_ok) {
    __shared__ long long scratch[34];
    __shared__ bool s_last;
    constexpr int NW = StencilWidth<S>::value;
    // one probe CTA per block of kTT consecutive particles; the last CTA to finish decides
    {
        TileConfig::Part part;
        part.mad[0] = part.mad[1] = part.mad[2] = 0.f;
        part.cnt = part.eq_pairs = part.all_pairs = part.near_pairs = 0;
        part.pad = 0;
        if (P.n >= 2 * kTT) {
            const int64_t start = (P.n - kTT) / (kProbeBlocks - 1) * blockIdx.x;
            const int64_t i = start + threadIdx.x;
            const double x = ld_stream(P.p[0] + i * P.stride[0]);
            const double y = ND >= 2 ? ld_stream(P.p[1] + i * P.stride[1]) : 0.0;
            const double z = ND >= 3 ? ld_stream(P.p[2] + i * P.stride[2]) : 0.0;
            int b[3];
            const bool ok = base_cells<ND, S>(P, x, y, z, b);
            const long long cnt = block_sum(ok ? 1 : 0, scratch);
            if (cnt >= 32) {
                part.cnt = (int)cnt;
                // consecutive particles with the same base cell (pairs inside a warp only)
                const int key = ok ? (b[0] * P.dims[1] + b[1]) * P.dims[2] + b[2] : -1;
                const int prev = __shfl_up_sync(0xffffffffu, key, 1);
                const bool pair = (threadIdx.x & 31) != 0 && key >= 0 && prev >= 0;
                part.eq_pairs = (int)block_sum(pair && key == prev ? 1 : 0, scratch);
                part.all_pairs = (int)block_sum(pair ? 1 : 0, scratch);
                // neighbours that a register window one cell wider per axis would still hold
                // (either particle of the pair may be the one further along)
                bool near = pair;
                for (int a = 0; a < ND; ++a) {
                    const int pb = __shfl_up_sync(0xffffffffu, b[a], 1);
                    near = near && (b[a] - pb) >= -1 && (b[a] - pb) <= 1;
                }
                part.near_pairs = (int)block_sum(near ? 1 : 0, scratch);
                for (int a = 0; a < ND; ++a) {
                    const long long sum = block_sum(ok ? b[a] : 0, scratch);
                    const long long mean = sum / cnt;
                    const long long dev = block_sum(ok ? llabs((long long)b[a] - mean) : 0, scratch);
                    part.mad[a] = (float)((double)dev / (double)cnt);
                }
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
    double mad_sum[3] = {0, 0, 0};
    int blocks_used = 0;
    long long eq_pairs = 0, all_pairs = 0, near_pairs = 0;
    for (int blk = 0; blk < (int)gridDim.x; ++blk) {
        const volatile TileConfig::Part *q = &cfg->part[blk];
        if (q->cnt < 32) continue;
        ++blocks_used;
        for (int a = 0; a < ND; ++a) mad_sum[a] += q->mad[a];
        eq_pairs += q->eq_pairs;
        all_pairs += q->all_pairs;
        near_pairs += q->near_pairs;
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
            const double f = pow((double)kTileCells / prod, 1.0 / ND);
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
        // A band of whole rows (x-ordered input, uniform along the last axis) is no win: its
        // shared-memory CAS atomics run at about the rate of the L2 reductions of method ROWS
        // (measured 36 vs 38 Gp/s, 2-D TSC).  The tile pays off when it is a true window on the
        // last axis (bunches), or when the whole grid fits (1-D and small histograms).
        bool whole = true;
        for (int a = 0; a < ND; ++a) whole = whole && c.ext[a] >= P.dims[a];
        const bool windowed_last = c.ext[ND - 1] < P.dims[ND - 1];
        if (force || (hit >= 0.45 && (whole || windowed_last || ND == 1))) c.mode = kModeTiled;
    }
    // long runs of equal base cells: register run-length accumulation beats the tile
    // (break-even is an average run of ~2.5 particles)
    if (!force && c.run_frac >= 0.6f) c.mode = kModeRuns;
    else if (!force && c.near_frac >= 0.75f && wide_ok) c.mode = kModeRunsWide;
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

template <int ND, int TT>
__device__ __forceinline__ void flush_tile(double *tile, const Window &wd, const DepositParams &P, int cells,
                                           int replicas) {
    // merge the tile (all its replicas) into the global grid, coalesced along the last axis, and clear it
    for (int c = threadIdx.x; c < cells; c += TT) {
        double v = tile[c];
        tile[c] = 0.0;
        for (int r = 1; r < replicas; ++r) { v = __dadd_rn(v, tile[r * cells + c]); tile[r * cells + c] = 0.0; }
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
            red_add(P.grid + lin, v);
        }
    }
}

template <int ND, int S, bool W, int TT, bool F>
__global__ void __launch_bounds__(TT, 1) k_deposit_tiled(const DepositParams P, const TileConfig *cfgp) {
    extern __shared__ double tile[];
    __shared__ long long scratch[34];
    __shared__ Window s_win;
    __shared__ unsigned s_acc, s_miss;
    constexpr int NW = StencilWidth<S>::value;

    const TileConfig cfg = *cfgp;
    if (cfg.mode != kModeTiled) return;
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

    // contiguous span of this CTA, a multiple of the CTA width
    int64_t per = (P.n + gridDim.x - 1) / gridDim.x;
    per = (per + TT - 1) / TT * TT;
    const int64_t begin = (int64_t)blockIdx.x * per;
    const int64_t end = begin + per < P.n ? begin + per : P.n;
    const int lane = threadIdx.x & 31;
    const double *px = P.p[0], *py = P.p[1], *pz = P.p[2];
    const int64_t sx = P.stride[0], sy = P.stride[1], sz = P.stride[2];

    // software pipeline: the raw values of the next sub-chunk are in flight while this one is deposited
    double nx = 0, ny = 0, nz = 0, nwt = 1.0;
    {
        const int64_t i = begin + threadIdx.x;
        if (i < end) {
            nx = ld_stream(px + i * sx);
            if (ND >= 2) ny = ld_stream(py + i * sy);
            if (ND >= 3) nz = ld_stream(pz + i * sz);
            if (W) nwt = particle_weight<F>(P, i);
        }
    }
    int since_eval = kEvalEvery - 1;       // first evaluation right after the first sub-chunk
    unsigned acc_w = 0, miss_w = 0;        // per-warp counters (lane 0), published at evaluation time
    for (int64_t i0 = begin; i0 < end; i0 += TT) {
        const int64_t i = i0 + threadIdx.x;
        const double x = nx, y = ny, z = nz, w = nwt;
        {
            const int64_t j = i + TT;
            if (j < end) {
                nx = ld_stream(px + j * sx);
                if (ND >= 2) ny = ld_stream(py + j * sy);
                if (ND >= 3) nz = ld_stream(pz + j * sz);
                if (W) nwt = particle_weight<F>(P, j);
            }
        }
        // ---- stencil in registers --------------------------------------------------
        int b[3];
        double wx[NW], wy[NW], wz[NW];
        const bool ok = particle_stencil<ND, S>(P, x, y, z, b, wx, wy, wz) && i < end && particle_selected(P, i);
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
            double *t = mytile + ((b[0] - wd.o[0]) * p1 + (ND >= 2 ? b[1] - wd.o[1] : 0)) * p2 + (ND >= 3 ? b[2] - wd.o[2] : 0);
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int d = 0; d < NDD; ++d) atomicAdd(t + (a * p1 + c) * p2 + d, cellval(a, c, d));
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
        if (++since_eval == kEvalEvery && i0 + TT < end) {
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
                flush_tile<ND, TT>(tile, wd, P, cells, replicas);
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
    flush_tile<ND, TT>(tile, wd, P, cells, replicas);
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

template <int ND, int S, bool W, bool F>
__global__ void __launch_bounds__(kST, 1) k_deposit_small(const DepositParams P, const TileConfig *gate) {
    if (gate != nullptr && gate->mode != kModeTiled) return;
    extern __shared__ double tile[];
    constexpr int NW = StencilWidth<S>::value;
    constexpr int NA = NW, NC = ND >= 2 ? NW : 1, NDD = ND >= 3 ? NW : 1;
    // Plain counting (shape 0, no weights -- numpy.histogram's job): every contribution is exactly 1,
    // so the tile holds 32-bit counters and uses the one NATIVE shared-memory atomic of sm_100a
    // (ATOMS.ADD.U32, 3.7x the rate of the fp64 CAS loop); the counts are exact and are converted to
    // double at the merge.  A CTA sees fewer than 2^32 particles for any n < 148 * 2^32.
    constexpr bool COUNT = (S == 0 && !W);
    const int cells = P.dims[0] * P.dims[1] * P.dims[2];
    int replicas = kTileCells / cells;
    if (replicas > kST / 32) replicas = kST / 32;
    if (replicas < 1) replicas = 1;
    unsigned *ctile = reinterpret_cast<unsigned *>(tile);
    if (COUNT) { for (int c = threadIdx.x; c < cells * replicas; c += kST) ctile[c] = 0u; }
    else       { for (int c = threadIdx.x; c < cells * replicas; c += kST) tile[c] = 0.0; }
    __syncthreads();
    double *mine = tile + ((threadIdx.x >> 5) % replicas) * cells;
    unsigned *cmine = ctile + ((threadIdx.x >> 5) % replicas) * cells;
    const int p1 = ND >= 2 ? P.dims[1] : 1, p2 = ND >= 3 ? P.dims[2] : 1;

    auto one = [&](double x, double y, double z, double w) {
        int b[3];
        double wx[NW], wy[NW], wz[NW];
        const bool ok = particle_stencil<ND, S>(P, x, y, z, b, wx, wy, wz);
        if (!ok) return;
        if constexpr (COUNT) {
            atomicAdd(cmine + (b[0] * p1 + b[1]) * p2 + b[2], 1u);
            return;
        }
        double *t = mine + (b[0] * p1 + b[1]) * p2 + b[2];
#pragma unroll
        for (int a = 0; a < NA; ++a)
#pragma unroll
            for (int c = 0; c < NC; ++c)
#pragma unroll
                for (int d = 0; d < NDD; ++d) {
                    atomicAdd(t + (a * p1 + c) * p2 + d, stencil_value<ND, S, W>(wx, wy, wz, a, c, d, w));
                }
    };

    const int64_t nthreads = (int64_t)gridDim.x * kST;
    int64_t i = (int64_t)blockIdx.x * kST + threadIdx.x;
    const double *px = P.p[0], *py = P.p[1], *pz = P.p[2];
    const int64_t sx = P.stride[0], sy = P.stride[1], sz = P.stride[2];
    for (; i + (kSU - 1) * nthreads < P.n; i += kSU * nthreads) {
        double x[kSU], y[kSU], z[kSU], w[kSU];
#pragma unroll
        for (int u = 0; u < kSU; ++u) {
            const int64_t k = i + u * nthreads;
            x[u] = ld_stream(px + k * sx);
            y[u] = ND >= 2 ? ld_stream(py + k * sy) : 0.0;
            z[u] = ND >= 3 ? ld_stream(pz + k * sz) : 0.0;
            w[u] = W ? particle_weight<F>(P, k) : 1.0;
        }
#pragma unroll
        for (int u = 0; u < kSU; ++u)
            if (particle_selected(P, i + u * nthreads)) one(x[u], y[u], z[u], w[u]);
    }
    for (; i < P.n; i += nthreads)
        if (particle_selected(P, i)) one(ld_stream(px + i * sx), ND >= 2 ? ld_stream(py + i * sy) : 0.0, ND >= 3 ? ld_stream(pz + i * sz) : 0.0,
            W ? particle_weight<F>(P, i) : 1.0);
    __syncthreads();
    for (int c = threadIdx.x; c < cells; c += kST) {
        double v;
        if (COUNT) {
            unsigned long long k = ctile[c];
            for (int r = 1; r < replicas; ++r) k += ctile[r * cells + c];
            v = (double)k;
        } else {
            v = tile[c];
            for (int r = 1; r < replicas; ++r) v = __dadd_rn(v, tile[r * cells + c]);
        }
        if (v != 0.0) red_add(P.grid + c, v);
    }
}

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
template <int ND, int S, bool W>
static cudaError_t launch_tiled(const DepositParams &P, TileConfig *cfg, int force, int sm_count, cudaStream_t st) {
    constexpr int NW = StencilWidth<S>::value;
    constexpr int TT = TileThreads<(ND == 1 ? NW : ND == 2 ? NW * NW : NW * NW * NW)>::value;
    const bool fuse = W && P.fuse_kind != 0;
    auto kern = fuse ? k_deposit_tiled<ND, S, W, TT, W> : k_deposit_tiled<ND, S, W, TT, false>;
    const size_t smem = (size_t)kTileCells * sizeof(double);
    // per-device attribute, cheap to set: done on every launch so that any device works
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_probe<ND, S><<<kProbeBlocks, kTT, 0, st>>>(P, cfg, force, runs_wide_available(ND, S) ? 1 : 0);
    if ((int64_t)P.dims[0] * P.dims[1] * P.dims[2] <= kTileCells) {
        // the whole padded grid fits the tile: lean kernel without window management
        auto small = fuse ? k_deposit_small<ND, S, W, W> : k_deposit_small<ND, S, W, false>;
        e = cudaFuncSetAttribute(small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int64_t want = (P.n + (int64_t)kST * kSU - 1) / ((int64_t)kST * kSU);
        small<<<(int)(want < sm_count ? (want < 1 ? 1 : want) : sm_count), kST, smem, st>>>(P, cfg);
        return cudaGetLastError();
    }
    kern<<<sm_count, TT, smem, st>>>(P, cfg);
    return cudaGetLastError();
}

template <int ND, int S>
static cudaError_t tiled_w(const DepositParams &P, TileConfig *cfg, int force, int sm, cudaStream_t st) {
    return (P.p[3] || P.fuse_kind) ? launch_tiled<ND, S, true>(P, cfg, force, sm, st) : launch_tiled<ND, S, false>(P, cfg, force, sm, st);
}

template <int ND>
static cudaError_t tiled_s(int shape, const DepositParams &P, TileConfig *cfg, int force, int sm, cudaStream_t st) {
    switch (shape) {
        case 0: return tiled_w<ND, 0>(P, cfg, force, sm, st);
        case 1: return tiled_w<ND, 1>(P, cfg, force, sm, st);
        case 2: return tiled_w<ND, 2>(P, cfg, force, sm, st);
        default: return tiled_w<ND, 3>(P, cfg, force, sm, st);
    }
}

cudaError_t deposit_tiled(int ndim, int shape, const DepositParams &P, TileConfig *cfg, bool force, int sm_count,
                          cudaStream_t st) {
    if (P.n <= 0) return cudaSuccess;
    switch (ndim) {
        case 1: return tiled_s<1>(shape, P, cfg, force, sm_count, st);
        case 2: return tiled_s<2>(shape, P, cfg, force, sm_count, st);
        default: return tiled_s<3>(shape, P, cfg, force, sm_count, st);
    }
}

}  // namespace ptg
