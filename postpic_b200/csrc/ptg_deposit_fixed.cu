// ptg_deposit_fixed.cu -- method FIXED: deterministic, ORDER-INDEPENDENT deposition by integer accumulation.
//
// Floating-point atomics make the result depend on the order in which the hardware happens to apply them.  Method
// SORTED removes that by sorting the particles and reducing every cell in a fixed order -- at 4 % of the HBM
// roofline.  This method removes the order dependence at the root instead: every contribution v (computed in fp64
// exactly as the reference does) is rounded once to an integer number of quanta q = 2^(e-45) (2^e = the power of two
// above the EXACT largest |weight|, found by one reduction pass over the weights) and all sums are integer sums --
// 64-bit two-word fixed point in the shared-memory tile (ptg_deposit_tiled.cu), two 64-bit words per cell in global
// memory (see fixed_accumulate() in ptg_stencil.cuh), so no number of particles can overflow a cell.  Integer addition is associative and
// commutative: the result is bit-identical from run to run, for any launch geometry, whichever kernel path the probe
// chooses, and even for any permutation of the input.  One conversion per cell at the end (k_fixed_finish).
// Per-contribution rounding <= 2^-46 of the largest weight; integer-valued weights below 2^45 are exact.
// NaN / inf weights poison their cells through the fp64 grid directly (the outcome of that does not depend on the
// order either).
//
// Replaces the loops of /root/reference/postpic/particles/_particlestogrid.pyx :70-129, :181-276, :341-463.
#include <cstdlib>
#include <cstring>

#include "ptg_internal.cuh"

namespace ptg {

constexpr int kFT = 256;

__global__ void k_fixed_init(FixedState *st, unsigned long long bits) { st->wmax_bits = bits; st->pad = 0; }

// largest finite |weight| of the selected particles (the weight as the deposit kernels see it, fused factor included)
template <bool F>
__global__ void __launch_bounds__(kFT) k_fixed_absmax(const DepositParams P, FixedState *st) {
    unsigned long long m = 0ull;
    const int64_t step = (int64_t)gridDim.x * kFT;
    for (int64_t i = (int64_t)blockIdx.x * kFT + threadIdx.x; i < P.n; i += step) {
        if (!particle_selected(P, i)) continue;
        const double a = fabs(particle_weight<F>(P, i));
        if (a <= 1.79769313486231570e308) {
            const unsigned long long b = (unsigned long long)__double_as_longlong(a);
            m = b > m ? b : m;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
        m = t > m ? t : m;
    }
    if ((threadIdx.x & 31) == 0 && m != 0ull) atomicMax(&st->wmax_bits, m);
}

__global__ void k_fixed_scale(TileConfig *cfg, const FixedState *st) {
    const double wmax = __longlong_as_double((long long)st->wmax_bits);
    int e = wmax > 0.0 ? ilogb(wmax) + 1 : 0;
    e = e < -960 ? -960 : (e > 960 ? 960 : e);
    cfg->fx_q = scalbn(1.0, e - 45);
    cfg->fx_inv_q = scalbn(1.0, 45 - e);
    cfg->fx_wcap = scalbn(1.0, e);
}

void fixed_set_scale(TileConfig *cfg, const FixedState *state, cudaStream_t st) { k_fixed_scale<<<1, 1, 0, st>>>(cfg, state); }

// Input the probe did not send to the shared-memory tile: straight into the 128-bit accumulators, 4 lanes per particle,
// lane j one cell of every stencil row along the last axis (adjacent lanes -> adjacent words: one L2 request per row).
template <int ND, int S, bool W, bool F>
__global__ void __launch_bounds__(kFT) k_fixed_deposit(const DepositParams P, const TileConfig *cfg) {
    if (cfg->mode == kModeTiled) return;
    constexpr int NW = StencilWidth<S>::value;
    const double inv_q = cfg->fx_inv_q;
    const int j = threadIdx.x & 3, group = threadIdx.x >> 2;
    constexpr int kGroups = kFT / 4;
    const int64_t span = ((P.n + gridDim.x - 1) / gridDim.x + kGroups - 1) / kGroups * kGroups;
    const int64_t begin = (int64_t)blockIdx.x * span;
    const int64_t end = begin + span < P.n ? begin + span : P.n;
    auto deposit_one = [&](int64_t i) {
        if (!particle_selected(P, i)) return;
        const double x = particle_coord<F>(P, 0, i);
        const double y = ND >= 2 ? particle_coord<F>(P, 1, i) : 0.0;
        const double z = ND >= 3 ? particle_coord<F>(P, 2, i) : 0.0;
        const double w = W ? particle_weight<F>(P, i) : 1.0;
        int b[3];
        double wx[NW], wy[NW], wz[NW];
        if (!particle_stencil<ND, S>(P, x, y, z, b, wx, wy, wz) || j >= NW || b[0] < P.slab_lo || b[0] >= P.slab_hi) return;
        // word index of this lane's cell in the accumulators (plain or staggered layout) and the fp64 grid index
        const int col = ND == 1 ? b[0] : ND == 2 ? b[1] : b[2];
        const int64_t rowlin = ND == 1 ? 0 : ND == 2 ? b[0] : (int64_t)b[0] * P.dims[1] + b[1];
        const bool stag = P.fixed_slo != nullptr;
        unsigned long long *const alo = stag ? P.fixed_slo : P.fixed_lo;
        long long *const ahi = stag ? P.fixed_shi : P.fixed_hi;
        const int64_t pitch = stag ? P.fixed_pitch : P.dims[ND - 1];
        int64_t lin = rowlin * pitch + col + j;
        if (stag && (NW == 3 ? (col & 2) != 0 : (col & 3) == 3)) lin += P.fixed_copy + 2;
        const int64_t glin = rowlin * P.dims[ND - 1] + col + j;
        const int64_t rs = ND == 1 ? 0 : pitch, ss = ND == 3 ? (int64_t)P.dims[1] * pitch : 0;              // accumulators
        const int64_t grs = ND == 1 ? 0 : P.dims[ND - 1], gss = ND == 3 ? (int64_t)P.dims[1] * P.dims[2] : 0;   // fp64 grid
        auto put = [&](int64_t off, int64_t goff, double v) {
            if (fabs(v) <= 1.79769313486231570e308)
                fixed_accumulate_words(alo + lin + off, ahi + lin + off, P.fixed_split, __double2ll_rn(__dmul_rn(v, inv_q)));
            else red_add(P.grid + glin + goff, v);
        };
        const double (&wl)[NW] = ND == 1 ? wx : ND == 2 ? wy : wz;
        double wj = wl[0];
#pragma unroll
        for (int k = 1; k < NW; ++k) wj = j == k ? wl[k] : wj;
        if (ND == 1) {
            put(0, 0, S == 0 ? (W ? w : 1.0) : (W ? __dmul_rn(wj, w) : wj));
        } else if (ND == 2) {
#pragma unroll
            for (int a = 0; a < NW; ++a) {
                const double c = __dmul_rn(wx[a], wj);
                put(a * rs, a * grs, S == 0 ? (W ? w : 1.0) : (W ? __dmul_rn(c, w) : c));
            }
        } else {
#pragma unroll
            for (int a = 0; a < NW; ++a)
#pragma unroll
                for (int c2 = 0; c2 < NW; ++c2) {
                    const double c = __dmul_rn(__dmul_rn(wx[a], wy[c2]), wj);
                    put(a * ss + c2 * rs, a * gss + c2 * grs, S == 0 ? (W ? w : 1.0) : (W ? __dmul_rn(c, w) : c));
                }
        }
    };
    if (P.slab_lo > -(1 << 30) || P.slab_hi < (1 << 30)) {
        // Slab pass: most particles belong to another pass.  Every LANE looks at one particle's axis-0 cell (the same
        // accept test and base cell as particle_stencil), the warp votes, and only the survivors -- compacted eight at
        // a time onto the eight 4-lane groups -- get the full stencil: one cheap test per particle instead of a whole
        // stencil evaluation on four lanes for particles that are then discarded (the kernel is issue-bound,
        // profiles/r02/ncu_full_fixed_direct_slab_cfg3.txt).  Integer sums: the order change alters nothing.
        const int lane = threadIdx.x & 31;
        for (int64_t base = begin + (threadIdx.x >> 5) * 32; base < end; base += kFT) {
            const int64_t i = base + lane;
            bool keep = false;
            if (i < end && particle_selected(P, i)) {
                const double x = particle_coord<F>(P, 0, i);
                int b0 = 0;
                bool in;
                if constexpr (S == 0) {
                    in = ngp_axis(x, P.lo[0], P.inv[0], P.fbins[0], b0);
                } else {
                    double t[NW];
                    in = shape_axis<S, false>(x, P.lo[0], P.inv[0], P.fbins_p1[0], b0, t);
                }
                keep = in && b0 >= P.slab_lo && b0 < P.slab_hi;
            }
            unsigned m = __ballot_sync(0xffffffffu, keep);
            while (m != 0u) {                                        // (uniform across the warp)
                const unsigned src = __fns(m, 0u, (lane >> 2) + 1);  // group g takes the g-th survivor
                if (src != 0xffffffffu) deposit_one(base + src);
#pragma unroll
                for (int k = 0; k < 8; ++k) m &= m - 1u;             // drop the eight just served (0 & anything stays 0)
            }
        }
        return;
    }
    for (int64_t i = begin + group; i < end; i += kGroups) deposit_one(i);
}

// grid[c] += (128-bit sum of cell c) * q: one thread per cell, one conversion, no atomics
__global__ void __launch_bounds__(256) k_fixed_finish(const DepositParams P, int64_t cells, const TileConfig *cfg) {
    const double q = cfg->fx_q;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += step) {
        // the cell's sum as one 128-bit integer, whichever representation was accumulated: the plain pair plus, when
        // the direct kernel used them, both copies of the staggered pair
        auto word = [&](const unsigned long long *lo, const long long *hi, int64_t i) -> __int128 {
            const unsigned long long lo0 = lo[i];
            const long long hi0 = hi[i];
            return P.fixed_split ? ((__int128)hi0 << 31) + (__int128)lo0 : ((__int128)hi0 << 64) + (__int128)lo0;
        };
        __int128 t = word(P.fixed_lo, P.fixed_hi, c);
        if (P.fixed_slo != nullptr) {
            const unsigned row = (unsigned)c / (unsigned)P.fixed_last;
            const int64_t idx = (int64_t)row * P.fixed_pitch + ((unsigned)c - row * (unsigned)P.fixed_last);
            t += word(P.fixed_slo, P.fixed_shi, idx) + word(P.fixed_slo, P.fixed_shi, P.fixed_copy + idx + 2);
        }
        if (t == 0) continue;
        const bool neg = t < 0;
        if (neg) t = -t;                             // magnitude: no cancellation in the conversion below
        const unsigned long long lo = (unsigned long long)t, hi = (unsigned long long)(t >> 64);
        double v = __ull2double_rn(lo);
        if (hi != 0) v = __dadd_rn(__dmul_rn(__ull2double_rn(hi), 18446744073709551616.0), v);
        v = __dmul_rn(neg ? -v : v, q);
        P.grid[c] = __dadd_rn(P.grid[c], v);
    }
}

template <int ND, int S, bool W>
static cudaError_t launch_fixed_deposit(const DepositParams &P, const TileConfig *cfg, int sm_count, cudaStream_t st) {
    int64_t want = (P.n + kFT / 4 - 1) / (kFT / 4);
    int64_t cap = (int64_t)sm_count * 8;
    const int blocks = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
    if (W && fused_mode(P)) k_fixed_deposit<ND, S, W, W><<<blocks, kFT, 0, st>>>(P, cfg);
    else k_fixed_deposit<ND, S, W, false><<<blocks, kFT, 0, st>>>(P, cfg);
    return cudaGetLastError();
}

template <int ND, int S>
static cudaError_t fixed_w(const DepositParams &P, const TileConfig *cfg, int sm, cudaStream_t st) {
    return (P.p[3] || fused_mode(P)) ? launch_fixed_deposit<ND, S, true>(P, cfg, sm, st) : launch_fixed_deposit<ND, S, false>(P, cfg, sm, st);
}

template <int ND>
static cudaError_t fixed_s(int shape, const DepositParams &P, const TileConfig *cfg, int sm, cudaStream_t st) {
    switch (shape) {
        case 0: return fixed_w<ND, 0>(P, cfg, sm, st);
        case 1: return fixed_w<ND, 1>(P, cfg, sm, st);
        case 2: return fixed_w<ND, 2>(P, cfg, sm, st);
        default: return fixed_w<ND, 3>(P, cfg, sm, st);
    }
}

cudaError_t deposit_fixed(int ndim, int shape, const DepositParams &P0, TileConfig *cfg, int sm_count, cudaStream_t st) {
    if (P0.n <= 0) return cudaSuccess;
    const int64_t cells = (int64_t)P0.dims[0] * P0.dims[1] * P0.dims[2];
    // accumulator layout: staggered (two sector-aligned copies, as method ROWS) for stencil rows of 2 or 3 cells
    const int last = P0.dims[ndim - 1];
    const int pitch = (last + 2 + 3) / 4 * 4;
    const int64_t nrows = cells / last;
    const int64_t copy = (nrows * pitch + 4) / 4 * 4;
    const bool stagger = (shape == 1 || shape == 2) && ndim >= 2 && P0.n >= (int64_t(1) << 20) &&
                         (size_t)(2 * copy) * 16 <= ((size_t)96 << 20);
    const int64_t swords = stagger ? 2 * copy + 4 : 0;
    const int64_t pwords = (cells + 3) / 4 * 4;          // keeps the staggered copies 32-byte aligned
    void *buf = nullptr;
    const size_t bytes = (size_t)(pwords + swords) * 16 + sizeof(FixedState);
    cudaError_t e = scratch_alloc(&buf, bytes, st);
    if (e != cudaSuccess) return e;
    auto fail = [&](cudaError_t err) { cudaFreeAsync(buf, st); return err; };
    if ((e = cudaMemsetAsync(buf, 0, bytes, st)) != cudaSuccess) return fail(e);
    DepositParams P = P0;
    P.fixed_lo = static_cast<unsigned long long *>(buf);
    P.fixed_hi = reinterpret_cast<long long *>(P.fixed_lo + pwords);
    P.fixed_slo = stagger ? P.fixed_lo + 2 * pwords : nullptr;
    P.fixed_shi = stagger ? reinterpret_cast<long long *>(P.fixed_slo + swords) : nullptr;
    FixedState *state = reinterpret_cast<FixedState *>(P.fixed_lo + 2 * (pwords + swords));
    P.fixed_pitch = stagger ? pitch : 0;
    P.fixed_last = last;
    P.fixed_copy = stagger ? copy : 0;
    // two fire-and-forget reductions per contribution while no cell can see 2^33 of them, the carry form beyond
    const int nw = shape == 0 ? 1 : shape + 1;
    const int64_t per_particle = ndim == 1 ? nw : ndim == 2 ? nw * nw : nw * nw * nw;
    static const bool force_carry = getenv("PTG_FIXED_CARRY") != nullptr;        // test hook
    P.fixed_split = (!force_carry && P.n < (int64_t(1) << 33) / per_particle) ? 1 : 0;
    const bool weighted = P.p[3] != nullptr || fused_mode(P);
    if (weighted) {
        int64_t want = (P.n + kFT * 4 - 1) / (kFT * 4);
        int64_t cap = (int64_t)sm_count * 8;
        const int blocks = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
        if (fused_mode(P)) k_fixed_absmax<true><<<blocks, kFT, 0, st>>>(P, state);
        else k_fixed_absmax<false><<<blocks, kFT, 0, st>>>(P, state);
    } else {
        const double one = 1.0;
        unsigned long long bits;
        memcpy(&bits, &one, 8);
        k_fixed_init<<<1, 1, 0, st>>>(state, bits);
    }
    // probe (+ the exact quantum) + the shared-memory tile kernel when the input mostly hits a window ...
    if ((e = deposit_tiled(ndim, shape, P, cfg, false, false, sm_count, st, state)) != cudaSuccess) return fail(e);
    // ... and the direct kernel for everything else
    // Accumulators larger than L2 (cfg3: 258^3 x 16 B = 275 MB) make nearly every reduction miss; as in method ROWS
    // (ptg_deposit_global.cu) such grids are deposited in slab passes over axis 0, each pass re-reading the particles
    // but reducing into an L2-resident slice.  Integer sums: the passes change nothing in the result.  cfg3 i.i.d.:
    // 3.6 -> 7.3 Gp/s with 72 MiB slices (100 / 48 / 36 MiB: 7.1 / 5.6 / 4.6; profiles/r02/perf_fixed_slabs.log).  Deciding
    // on axis 0 per 4-lane group was measured and is slower (5.9 Gp/s: the warp diverges without shedding work,
    // perf_fixed_slabs_earlyout.log), and so is batching the loads (6.2 Gp/s, perf_fixed_batched_rejected.log); one pass is
    // issue-bound (ncu_full_fixed_direct_slab_cfg3.txt) -- hence the vote-and-compact form of the kernel's slab branch.
    int passes = 1;
    {
        static const int64_t slab_mb = [] { const char *s = getenv("PTG_FIXED_SLAB_MB"); int64_t v = s ? atoll(s) : 72; return v < 1 ? 1 : v; }();
        const int64_t acc_bytes = cells * 16;
        if (acc_bytes > ((int64_t)96 << 20) && P.n >= (int64_t(1) << 22)) passes = (int)((acc_bytes + (slab_mb << 20) - 1) / (slab_mb << 20));
        if (passes > P.dims[0]) passes = P.dims[0];
    }
    for (int s = 0; s < passes; ++s) {
        P.slab_lo = s == 0 ? -(1 << 30) : (int)((int64_t)P.dims[0] * s / passes);
        P.slab_hi = s == passes - 1 ? (1 << 30) : (int)((int64_t)P.dims[0] * (s + 1) / passes);
        switch (ndim) {
            case 1: e = fixed_s<1>(shape, P, cfg, sm_count, st); break;
            case 2: e = fixed_s<2>(shape, P, cfg, sm_count, st); break;
            default: e = fixed_s<3>(shape, P, cfg, sm_count, st); break;
        }
        if (e != cudaSuccess) return fail(e);
    }
    int64_t fb = (cells + 255) / 256;
    int64_t cap = (int64_t)sm_count * 8;
    k_fixed_finish<<<(int)(fb < cap ? fb : cap), 256, 0, st>>>(P, cells, cfg);
    e = cudaGetLastError();
    cudaError_t e2 = cudaFreeAsync(buf, st);
    return e != cudaSuccess ? e : e2;
}

}  // namespace ptg
