// ptg_stencil.cuh -- per-axis shape-function evaluation, in registers.
//
// Restates the arithmetic of /root/reference/postpic/particles/_particlestogrid.pyx
// (1-D: :71-129, 2-D: :182-276, 3-D: :342-463) for one axis.  Every floating-point
// operation is written with an explicit round-to-nearest intrinsic (__dadd_rn,
// __dmul_rn ...) so that nvcc can never contract a multiply and an add into an FMA:
// the reference build (gcc -O2, baseline x86-64) contains no FMA, and a fused
// `u + 0.5` right after the scaling multiply could move a particle that sits within
// one ulp of a cell boundary into the neighbouring cell.  With this, the contribution
// of every particle to every cell is bit-identical to the reference's; only the
// order in which contributions are summed differs.
#pragma once
#include <cstdint>

namespace ptg {

// Number of stencil cells along one axis.
template <int S> struct StencilWidth { static constexpr int value = S + 1; };
template <> struct StencilWidth<0> { static constexpr int value = 1; };

// x / 6 correctly rounded without the ~30-instruction IEEE division sequence:
// q = RN(x * RN(1/6)); r = x - 6 q (exact in an FMA); result = RN(q + r * RN(1/6)).
// (Markstein's correction step; checked bit-for-bit against `/ 6.0` by
// tests/test_gpu_parity.py::test_div6_bit_exact through ptg_selftest_div6.)
__device__ __forceinline__ double div6(double x) {
    const double y = 1.0 / 6.0;
    const double q = __dmul_rn(x, y);
    const double r = __fma_rn(-6.0, q, x);
    return __fma_rn(r, y, q);
}

// Scaled coordinate u = (d - lo) * inv  (pyx:71, :182-183, :342-344 ...).
__device__ __forceinline__ double scaled(double d, double lo, double inv) {
    return __dmul_rn(__dsub_rn(d, lo), inv);
}

// Shape 0 (NGP) on one axis: accept iff u > 0 and u < bins (both strict; pyx:72, :184,
// :345); cell = trunc(u).  NaN fails both comparisons.
__device__ __forceinline__ bool ngp_axis(double d, double lo, double inv, double fbins, int &cell) {
    const double u = scaled(d, lo, inv);
    const bool ok = (u > 0.0) && (u < fbins);
    cell = ok ? __double2int_rz(u) : 0;
    return ok;
}

// Shapes 1..3 on one axis.
//   ON_DOUBLE = false : the reference tests the truncated int, `r >= 0 and r <= bins`
//                       (x always; z for 3-D shapes 2,3; pyx:85,:101,:119,:200,:228,:258,:399,:438)
//   ON_DOUBLE = true  : the reference tests the raw double on the left, `u >= 0 and r <= bins`
//                       (y always; z for 3-D shape 1; pyx:200,:228,:258,:363)
// The int tests are evaluated in the double domain (trunc(v) >= 0 <=> v > -1,
// trunc(v) <= bins <=> v < bins + 1), which gives x86's outcome for NaN / inf / |v| >= 2^31
// on an int-tested axis (cvttsd2si -> INT_MIN -> rejected) and DROPS the particles for
// which the reference runs into undefined behaviour on a double-tested axis.
// base = first padded cell index; wt[0..S] = stencil weights.
template <int S, bool ON_DOUBLE>
__device__ __forceinline__ bool shape_axis(double d, double lo, double inv, double fbins_p1,
                                           int &base, double (&wt)[S + 1]) {
    static_assert(S >= 1 && S <= 3, "shape order");
    const double u = scaled(d, lo, inv);
    const double v = (S == 2) ? u : __dadd_rn(u, 0.5);
    bool ok = v < fbins_p1;
    if (ON_DOUBLE) ok = ok && (u >= 0.0);
    else           ok = ok && (v > -1.0);
    if (!ok) { base = 0; return false; }
    const int r = __double2int_rz(v);
    const double rd = __int2double_rn(r);
    if (S == 1) {
        wt[0] = __dadd_rn(__dsub_rn(0.5, u), rd);        // (0.5 - u) + r   pyx:88,:201
        wt[1] = __dsub_rn(__dadd_rn(0.5, u), rd);        // (0.5 + u) - r   pyx:87,:202
        base = r;
    } else if (S == 2) {
        const double t = __dsub_rn(u, rd);               // pyx:100,:226
        const double omt = __dsub_rn(1.0, t);
        const double tt = __dmul_rn(t, t);
        wt[0] = __dmul_rn(0.5, __dmul_rn(omt, omt));     // 0.5*(1-t)^2     pyx:103,:229
        wt[1] = __dsub_rn(__dadd_rn(0.5, t), tt);        // 0.5 + t - t^2   pyx:104,:230
        wt[2] = __dmul_rn(0.5, tt);                      // 0.5*t^2         pyx:105,:231
        base = r + 1;
    } else {
        const double t = __dadd_rn(__dsub_rn(u, rd), 0.5);   // pyx:118,:256
        const double sixth = 1.0 / 6.0, twothirds = 2.0 / 3.0;
        const double th = __dmul_rn(t, 0.5);                 // t/2 (exact)
        // 1/6 + t*(-0.5 + (0.5 - t/6)*t)                       pyx:121,:259
        wt[0] = __dadd_rn(sixth, __dmul_rn(t, __dadd_rn(-0.5, __dmul_rn(__dsub_rn(0.5, div6(t)), t))));
        // 2/3 + (-1 + t/2)*t*t                                 pyx:122,:260
        wt[1] = __dadd_rn(twothirds, __dmul_rn(__dmul_rn(__dadd_rn(-1.0, th), t), t));
        // 1/6 + t*(0.5 + (0.5 - t/2)*t)                        pyx:123,:261
        wt[2] = __dadd_rn(sixth, __dmul_rn(t, __dadd_rn(0.5, __dmul_rn(__dsub_rn(0.5, th), t))));
        // t*t*t/6                                              pyx:124,:262
        wt[3] = div6(__dmul_rn(__dmul_rn(t, t), t));
        base = r + 1;
    }
    return true;
}

// Everything a deposit kernel needs to know about the job (passed by value as the
// kernel argument: lives in the constant bank, costs no registers until used).
struct DepositParams {
    const double *p[4];     // x, y, z, w   (w may be null)
    int64_t stride[4];      // element strides
    int64_t n;
    double lo[3], inv[3];
    double fbins[3];        // (double) bins
    double fbins_p1[3];     // (double) bins + 1.0
    int bins[3];
    int dims[3];            // padded dims = bins + 2*shape
    double *grid;
    // slab pass (methods GLOBAL / ROWS on grids larger than L2): only particles whose stencil base
    // cell on axis 0 lies in [slab_lo, slab_hi) are deposited by this launch
    int slab_lo, slab_hi;
    // staggered staging grid of method ROWS (null = deposit straight into `grid`): two copies of
    // the grid with rows padded to `stag_pitch` doubles (a multiple of 4, 32-byte aligned base);
    // copy 1 stores column j at j + 2.  A stencil row of 2 or 3 cells is sent to the copy in
    // which it does not straddle a 32-byte sector, so every row costs exactly one L2 sector-RED.
    double *stag;
    int stag_pitch;
    int64_t stag_copy;      // distance between the two copies, in doubles
    // fused kinetic weight (0 = off): the particle's weight is weight[i] (or 1) times a scalar of
    // its momentum, evaluated in the kernel instead of in a separate pass
    // predicated deposition (null = every particle): particle i takes part iff mask[i * mstride] != 0 --
    // postpic's filter / compress (particles.py:219-232, :561-582) without materialising compressed copies
    const unsigned char *mask;
    int64_t mstride;
    int fuse_kind;          // scalar (PTG_S_* of ptg_b200.h) the particle weight is multiplied with, 0 = none
    // fused AXIS expressions (0 = coordinate a is the array p[a]): coordinate a of particle i is the scalar
    // akind[a] of postpic's default scalar context, evaluated in the kernel from the atomic arrays below --
    // createField('x_um', 'angle_xy', weights='gamma') is then ONE pass with no N-sized temporary
    int akind[3];
    int any_akind;
    const double *q[3];     // px, py, pz
    int64_t qstride[3];
    const double *pos[3];   // atomic x, y, z (for expression kinds; independent of the coordinate slots p[0..2])
    int64_t pos_stride[3];
    double f_mass, f_mc2, f_c2, f_qe;   // mass, (mass*c)^2, c^2, elementary charge
    double f_c, f_charge, f_time, f_amu, f_weight;   // c, species charge, dump time, atomic mass unit, constant macro weight
    // method FIXED (deterministic): every contribution is added as an integer number of quanta to a 128-bit
    // accumulator per cell (low / high 64-bit words); null = the ordinary fp64 grid is the target
    unsigned long long *fixed_lo;
    long long *fixed_hi;
    int fixed_split;        // representation of the accumulators, see fixed_accumulate()
    // A second pair of accumulators in the staggered layout of method ROWS, for the kernel that deposits unordered
    // input directly (k_fixed_deposit); null when not allocated.  Rows of the last axis (fixed_last cells) are padded to
    // fixed_pitch words, a second copy fixed_copy words further on stores column j at j + 2; a stencil row of 2 or 3
    // cells goes to the copy in which it does not straddle a 32-byte sector.  Tile merges and miss paths use the plain
    // pair above (cell c at word c); k_fixed_finish adds everything up.
    unsigned long long *fixed_slo;
    long long *fixed_shi;
    int fixed_pitch, fixed_last;
    int64_t fixed_copy;
};

// the kernels' fused (F = true) instances do the work when a weight or an axis expression is evaluated in the kernel
__host__ __device__ inline bool fused_mode(const DepositParams &P) { return P.fuse_kind != 0 || P.any_akind != 0; }

__device__ __forceinline__ bool particle_selected(const DepositParams &P, int64_t i) {
#ifdef PTG_NO_MASK
    return true;                       // A/B builds only
#endif
    return P.mask == nullptr || P.mask[i * P.mstride] != 0;
}

// Stencil of one particle on all axes: accept test, padded base cell b[a] and the per-axis weights
// (a single 1.0 for shape 0).  Axis kinds as in the reference: x on the truncated int, y on the raw
// double, z on the raw double only for shape 1.
template <int ND, int S>
__device__ __forceinline__ bool particle_stencil(const DepositParams &P, double x, double y, double z, int (&b)[3],
                                                 double (&wx)[StencilWidth<S>::value],
                                                 double (&wy)[StencilWidth<S>::value],
                                                 double (&wz)[StencilWidth<S>::value]) {
    b[0] = b[1] = b[2] = 0;
    bool ok;
    if constexpr (S == 0) {
        ok = ngp_axis(x, P.lo[0], P.inv[0], P.fbins[0], b[0]);
        if (ND >= 2) ok = ngp_axis(y, P.lo[1], P.inv[1], P.fbins[1], b[1]) && ok;
        if (ND >= 3) ok = ngp_axis(z, P.lo[2], P.inv[2], P.fbins[2], b[2]) && ok;
        wx[0] = wy[0] = wz[0] = 1.0;
    } else {
        ok = shape_axis<S, false>(x, P.lo[0], P.inv[0], P.fbins_p1[0], b[0], wx);
        if (ND >= 2) ok = shape_axis<S, true>(y, P.lo[1], P.inv[1], P.fbins_p1[1], b[1], wy) && ok;
        if (ND >= 3) ok = shape_axis<S, (S == 1)>(z, P.lo[2], P.inv[2], P.fbins_p1[2], b[2], wz) && ok;
    }
    return ok;
}

// Contribution of a particle to stencil cell (a, c, d) in the reference's operation order:
// 1-D  stencil*w (pyx:90,:107);  2-D (wx*wy)*w (pyx:211,:244,:276);  3-D ((wx*wy)*wz)*w (pyx:382,:421,:463);
// shape 0 adds the weight itself (or 1.0), the unweighted branch has no multiply by w.
template <int ND, int S, bool W, int N>
__device__ __forceinline__ double stencil_value(const double (&wx)[N], const double (&wy)[N], const double (&wz)[N],
                                                int a, int c, int d, double w) {
    if (S == 0) return W ? w : 1.0;
    if (ND == 1) return W ? __dmul_rn(wx[a], w) : wx[a];
    double v = __dmul_rn(wx[a], wy[c]);
    if (ND == 3) v = __dmul_rn(v, wz[d]);
    return W ? __dmul_rn(v, w) : v;
}

// Streaming 8-byte load: read-only path, do not allocate in L1 and mark the line evict-first
// in L2 (each particle is touched exactly once; the L2 belongs to the grid the reductions
// hit).  ptxas folds the policy into the load's descriptor, it costs no instruction.
__device__ __forceinline__ double ld_stream(const double *p) {
    double v;
#ifdef PTG_NO_EVICT_FIRST
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
#endif
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}

// The kinetic scalars of postpic's default scalar context (scalarproperties.py:183-205) in the order
// the nested expressions prescribe: _np2 = (px**2+py**2+pz**2)/(mass*c)**2, gamma_m1 = _np2/(sqrt(1+_np2)+1),
// Ekin = gamma_m1*mass*c**2, Ekin_MeV = Ekin/elementary_charge/1e6.  kind: PTG_W_* of ptg_b200.h.
__device__ __forceinline__ double kinetic_value(int kind, double x, double y, double z, double mass, double mc2,
                                                double c2, double qe) {
    const double p2 = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
    if (kind == 6) return sqrt(p2);                                        // p
    const double np2 = p2 / mc2;
    const double gm1 = np2 / __dadd_rn(sqrt(__dadd_rn(1.0, np2)), 1.0);
    switch (kind) {
        case 1: return gm1;                                                // gamma_m1
        case 2: return __dadd_rn(gm1, 1.0);                                // gamma
        case 3: return __dmul_rn(__dmul_rn(gm1, mass), c2);                // Ekin [J]
        case 4: return __dmul_rn(__dmul_rn(gm1, mass), c2) / qe / 1e6;     // Ekin_MeV
        default: return __dmul_rn(__dmul_rn(gm1, mass), c2) / qe / 1e3;    // Ekin_keV
    }
}

// Scalar `kind` (PTG_S_* of ptg_b200.h) of particle i: postpic's default scalar context
// (/root/reference/postpic/particles/scalarproperties.py:159-222) as enumerated kinds, every expression in the
// operation order its definition string prescribes (the library is built with -fmad=false: no contraction).
// The atomic arrays are read through the read-only path; a kind only touches the arrays it needs.
__device__ __forceinline__ double scalar_value(const DepositParams &P, int kind, int64_t i) {
    auto X = [&](int a) { return __ldg(P.pos[a] + i * P.pos_stride[a]); };
    auto Q = [&](int a) { return __ldg(P.q[a] + i * P.qstride[a]); };
    if (kind <= 6) return kinetic_value(kind, Q(0), Q(1), Q(2), P.f_mass, P.f_mc2, P.f_c2, P.f_qe);
    switch (kind) {
        case 7: return X(0);
        case 8: return X(1);
        case 9: return X(2);
        case 10: return Q(0);
        case 11: return Q(1);
        case 12: return Q(2);
        case 13: return P.p[3] ? __ldg(P.p[3] + i * P.stride[3]) : P.f_weight;      // weight
        case 14: return P.f_mass;
        case 15: return P.f_charge;
        case 16: return P.f_time;
        case 17: return P.f_mass / P.f_amu;                                        // mass_u
        case 18: return P.f_charge / P.f_qe;                                       // charge_e
        case 19: return X(0) * 1e6;                                                // x_um
        case 20: return X(1) * 1e6;
        case 21: return X(2) * 1e6;
        case 46: { const double a = X(0), b = X(1); return sqrt(a * a + b * b); }  // r_xy
        case 47: { const double a = X(1), b = X(2); return sqrt(a * a + b * b); }  // r_yz
        case 48: { const double a = X(2), b = X(0); return sqrt(a * a + b * b); }  // r_zx
        case 49: { const double a = X(0), b = X(1), c = X(2); return sqrt(a * a + b * b + c * c); }   // r_xyz
        default: break;
    }
    const double px = Q(0), py = Q(1), pz = Q(2);
    switch (kind) {
        case 37: return atan2(sqrt(py * py + pz * pz), px);                        // angle_xaxis
        case 38: return atan2(sqrt(pz * pz + px * px), py);
        case 39: return atan2(sqrt(px * px + py * py), pz);
        case 40: return atan2(py, px);                                             // angle_xy
        case 41: return atan2(px, py);
        case 42: return atan2(pz, py);
        case 43: return atan2(py, pz);
        case 44: return atan2(px, pz);
        case 45: return atan2(pz, px);
        default: break;
    }
    const double np2 = (px * px + py * py + pz * pz) / P.f_mc2;                    // _np2
    const double gm1 = np2 / (sqrt(1.0 + np2) + 1.0);                              // gamma_m1
    const double gamma = gm1 + 1.0;
    const double ekin = gm1 * P.f_mass * P.f_c2;
    const double mass_u = P.f_mass / P.f_amu, charge_e = P.f_charge / P.f_qe;
    switch (kind) {
        case 22: return np2;
        case 23: return gamma * P.f_mass;                                          // gamma_m
        case 24: return sqrt(gm1 * gm1 + 2.0 * gm1) / (gm1 + 1.0);                 // beta
        case 25: return sqrt(gm1 * gm1 + 2.0 * gm1) / (gm1 + 1.0) * P.f_c;         // v
        case 26: return px / (gamma * P.f_mass);                                   // vx
        case 27: return py / (gamma * P.f_mass);
        case 28: return pz / (gamma * P.f_mass);
        case 29: return px / (gamma * P.f_mass) / P.f_c;                           // betax
        case 30: return py / (gamma * P.f_mass) / P.f_c;
        case 31: return pz / (gamma * P.f_mass) / P.f_c;
        case 32: return P.f_mass * P.f_c2;                                         // Eruhe
        case 33: return ekin / P.f_qe / 1e6 / mass_u;                              // Ekin_MeV_amu
        case 34: return ekin / P.f_qe / 1e6 / mass_u * charge_e;                   // Ekin_MeV_qm
        case 35: return ekin / P.f_qe / 1e3 / mass_u;                              // Ekin_keV_amu
        case 36: return ekin / P.f_qe / 1e3 / mass_u * charge_e;                   // Ekin_keV_qm
        default: return 0.0;
    }
}

// Coordinate `a` of particle i: the array in slot a, or -- in the fused instances -- an axis expression.
template <bool F>
__device__ __forceinline__ double particle_coord(const DepositParams &P, int a, int64_t i) {
    if (F && P.akind[a] != 0) return scalar_value(P, P.akind[a], i);
    return ld_stream(P.p[a] + i * P.stride[a]);
}
// the same for kernels that are not templated on F (probe, sort keys): run-time check
__device__ __forceinline__ double particle_coord_rt(const DepositParams &P, int a, int64_t i) {
    if (P.akind[a] != 0) return scalar_value(P, P.akind[a], i);
    return ld_stream(P.p[a] + i * P.stride[a]);
}

// Weight of particle i for the weighted kernels: the weights array (1 / the species' constant when absent), times
// the fused scalar when one is requested.
__device__ __forceinline__ double fused_factor(const DepositParams &P, int64_t i) {
    return P.fuse_kind != 0 ? scalar_value(P, P.fuse_kind, i) : 1.0;
}
// F = false is the plain weighted kernel (the weights array exists); F = true instantiates the fused
// variants, whose weights array is optional.  A template flag, not a run-time one: the fused code
// costs the plain kernels no register.
template <bool F>
__device__ __forceinline__ double particle_weight(const DepositParams &P, int64_t i) {
    if (!F) return ld_stream(P.p[3] + i * P.stride[3]);
    const double w = P.p[3] ? ld_stream(P.p[3] + i * P.stride[3]) : P.f_weight;
    return P.fuse_kind != 0 ? __dmul_rn(w, scalar_value(P, P.fuse_kind, i)) : w;
}

// Fire-and-forget fp64 reduction into global memory: a single native
// RED.E.ADD.F64.RN.STRONG.GPU, resolved by the L2 atomic units.
__device__ __forceinline__ void red_add(double *addr, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}
// The same with the accumulator line marked evict-last in L2 when `keep` is set -- the counterpart of the evict-first
// particle loads.  Used by the slab passes of grids larger than L2 (ptg_deposit_global.cu): without the hint the
// particle stream pushes the pass's slab out of L2 (hit rate 47 %, 2.9 GB of accumulator lines written back per pass,
// profiles/r02/ncu_full_quads_cfg3.txt); with it cfg3 i.i.d. goes from 24.1 to 33.5 Gp/s.  On L2-resident grids the
// hint buys nothing (cfg2 52.4 vs 51.3 Gp/s, profiles/r02/perf_evictlast.log), hence the switch.
__device__ __forceinline__ unsigned long long l2_keep_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void red_add_hint(double *addr, double v, bool keep, unsigned long long pol) {
    if (keep) asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(addr), "d"(v), "l"(pol) : "memory");
    else asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

// Integer accumulation in global memory, two 64-bit words per cell, in one of two representations:
//   split (fixed_split = 1, the default): value = hi * 2^31 + lo.  A contribution k is added as its low 31 bits to
//     `lo` and k >> 31 (arithmetic) to `hi` with two fire-and-forget reductions (RED.E.ADD.64, no return trip).
//     `lo` takes 2^33 contributions per cell before it could wrap; the host selects this form when n * cells-per-
//     particle stays below that.
//   carry (fixed_split = 0): value = hi * 2^64 + lo, a true 128-bit sum: the returning add on the low word tells
//     whether it wrapped, the high word gets the carry / sign extension.  No limit, ~2x slower (return trips).
// Integer addition is associative and commutative: the sums do not depend on the order of the atomics.
__device__ __forceinline__ void fixed_accumulate_words(unsigned long long *lo, long long *hi, int split, long long k) {
    if (split) {
        atomicAdd(lo, (unsigned long long)(k & 0x7fffffffll));
        const long long h = k >> 31;
        if (h != 0) atomicAdd(reinterpret_cast<unsigned long long *>(hi), (unsigned long long)h);
    } else {
        const unsigned long long a = (unsigned long long)k;
        const unsigned long long old = atomicAdd(lo, a);
        const long long h = (k < 0 ? -1ll : 0ll) + ((old + a) < old ? 1ll : 0ll);
        if (h != 0) atomicAdd(reinterpret_cast<unsigned long long *>(hi), (unsigned long long)h);
    }
}
// `lin` = plain linear cell index of the padded grid
__device__ __forceinline__ void fixed_accumulate(const DepositParams &P, int64_t lin, long long k) {
    fixed_accumulate_words(P.fixed_lo + lin, P.fixed_hi + lin, P.fixed_split, k);
}
// one contribution for method FIXED: finite values as quanta, NaN / inf straight into the fp64 grid (they poison the
// cell whatever the order)
__device__ __forceinline__ void fixed_add(const DepositParams &P, int64_t lin, double v, double inv_q) {
    if (fabs(v) <= 1.79769313486231570e308) fixed_accumulate(P, lin, __double2ll_rn(__dmul_rn(v, inv_q)));
    else red_add(P.grid + lin, v);
}

}  // namespace ptg
