// ptg_util.cu -- the small full-array passes either side of the deposit:
//   min/max of a coordinate array (replaces np.min/np.max at
//   /root/reference/postpic/particles/_routines.py:117-126, used when a range end is None)
//   grid *= 1/dV (replaces `h /= dV`, /root/reference/postpic/particles/particles.py:1101-1102)
#include "ptg_internal.cuh"

namespace ptg {

// Order-preserving map double -> uint64 so that the native 64-bit integer
// atomicMin/atomicMax of the L2 can do the cross-block step.  NaN is mapped to the
// extreme key of the respective reduction, so a NaN anywhere yields NaN like numpy.
__device__ __forceinline__ unsigned long long key_of(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double value_of(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
constexpr unsigned long long kNanLow = 0ull, kNanHigh = ~0ull;

__global__ void k_minmax_init(unsigned long long *keys) {
    keys[0] = kNanHigh - 1;   // larger than the key of +inf
    keys[1] = kNanLow + 1;    // smaller than the key of -inf
}

__global__ void __launch_bounds__(256) k_minmax(int64_t n, const double *x, int64_t stride,
                                                unsigned long long *keys) {
    unsigned long long lo = kNanHigh - 1, hi = kNanLow + 1;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const double v = ld_stream(x + i * stride);
        unsigned long long kl, kh;
        if (v != v) { kl = kNanLow; kh = kNanHigh; } else { kl = kh = key_of(v); }
        lo = kl < lo ? kl : lo;
        hi = kh > hi ? kh : hi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ol = __shfl_xor_sync(0xffffffffu, lo, o);
        const unsigned long long oh = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = ol < lo ? ol : lo;
        hi = oh > hi ? oh : hi;
    }
    __shared__ unsigned long long slo[8], shi[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { slo[warp] = lo; shi[warp] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) { lo = slo[k] < lo ? slo[k] : lo; hi = shi[k] > hi ? shi[k] : hi; }
        atomicMin(keys, lo);
        atomicMax(keys + 1, hi);
    }
}

__global__ void k_minmax_decode(unsigned long long *keys) {
    const unsigned long long lo = keys[0], hi = keys[1];
    double *out = reinterpret_cast<double *>(keys);
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    out[0] = (lo == kNanLow) ? qnan : value_of(lo);
    out[1] = (hi == kNanHigh) ? qnan : value_of(hi);
}

cudaError_t minmax_f64(int64_t n, const double *x, int64_t stride, double *out2, int sm_count,
                       cudaStream_t st) {
    auto *keys = reinterpret_cast<unsigned long long *>(out2);
    k_minmax_init<<<1, 1, 0, st>>>(keys);
    if (n > 0) {
        int64_t want = (n + 256 * 8 - 1) / (256 * 8);
        int64_t cap = (int64_t)sm_count * 8;
        int blocks = (int)(want < cap ? want : cap);
        k_minmax<<<blocks, 256, 0, st>>>(n, x, stride, keys);
    }
    k_minmax_decode<<<1, 1, 0, st>>>(keys);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_scale(double *g, int64_t count, double f) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += step)
        g[i] = __dmul_rn(g[i], f);
}

// g[i] /= divisor with a true IEEE division: bit-identical to the reference's `h /= dV` (particles.py:1101-1102)
__global__ void __launch_bounds__(256) k_divide(double *g, int64_t count, double dv) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += step) g[i] = __ddiv_rn(g[i], dv);
}

cudaError_t divide_f64(double *grid, int64_t count, double divisor, int sm_count, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    int64_t want = (count + 255) / 256;
    int64_t cap = (int64_t)sm_count * 8;
    k_divide<<<(int)(want < cap ? want : cap), 256, 0, st>>>(grid, count, divisor);
    return cudaGetLastError();
}

cudaError_t scale_f64(double *grid, int64_t count, double factor, int sm_count, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    int64_t want = (count + 255) / 256;
    int64_t cap = (int64_t)sm_count * 8;
    k_scale<<<(int)(want < cap ? want : cap), 256, 0, st>>>(grid, count, factor);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// per-particle weight expressions in one pass
// ---------------------------------------------------------------------------------
// out[i] = weight[i] * f(px[i], py[i], pz[i]) for the kinetic scalars of postpic's default scalar
// context (/root/reference/postpic/particles/scalarproperties.py:183-205), evaluated in the order the
// nested expressions prescribe:
//   _np2     = (px**2 + py**2 + pz**2) / (mass * c)**2
//   gamma_m1 = _np2 / (sqrt(1 + _np2) + 1)
//   Ekin     = gamma_m1 * mass * c**2          Ekin_MeV = Ekin / elementary_charge / 1e6
// The reference gets there through ~10 numexpr passes and as many N-sized temporaries; here it is one
// read of px, py, pz (+ weight) and one write.
__global__ void __launch_bounds__(256) k_weights(int kind, int64_t n, const double *px, const double *py,
                                                 const double *pz, const double *weight, int64_t s0, int64_t s1,
                                                 int64_t s2, int64_t s3, double mass, double c, double qe,
                                                 double *out) {
    const double mc = __dmul_rn(mass, c);
    const double mc2 = __dmul_rn(mc, mc);                  // (mass * c)**2
    const double c2 = __dmul_rn(c, c);                      // c**2
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const double x = ld_stream(px + i * s0), y = ld_stream(py + i * s1), z = ld_stream(pz + i * s2);
        const double v = kinetic_value(kind, x, y, z, mass, mc2, c2, qe);
        out[i] = weight ? __dmul_rn(ld_stream(weight + i * s3), v) : v;
    }
}

cudaError_t weights_f64(int kind, int64_t n, const double *px, const double *py, const double *pz,
                        const double *weight, const int64_t *strides, double mass, double c, double qe, double *out,
                        int sm_count, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    int64_t want = (n + 256 * 4 - 1) / (256 * 4);
    int64_t cap = (int64_t)sm_count * 8;
    k_weights<<<(int)(want < cap ? want : cap), 256, 0, st>>>(kind, n, px, py, pz, weight, strides ? strides[0] : 1,
                                                             strides ? strides[1] : 1, strides ? strides[2] : 1,
                                                             strides ? strides[3] : 1, mass, c, qe, out);
    return cudaGetLastError();
}

// out[i] = scalar `kind` of particle i (the very device function the fused deposit kernels call)
__global__ void __launch_bounds__(256) k_scalar(const DepositParams P, int kind, double *out) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += step) out[i] = scalar_value(P, kind, i);
}

cudaError_t scalar_f64(const DepositParams &P, int kind, double *out, int sm_count, cudaStream_t st) {
    if (P.n <= 0) return cudaSuccess;
    int64_t want = (P.n + 255) / 256;
    int64_t cap = (int64_t)sm_count * 8;
    k_scalar<<<(int)(want < cap ? want : cap), 256, 0, st>>>(P, kind, out);
    return cudaGetLastError();
}

__global__ void k_div6(int64_t n, const double *in, double *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = div6(in[i]);
}

cudaError_t selftest_div6(int64_t n, const double *in, double *out, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    k_div6<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, in, out);
    return cudaGetLastError();
}

}  // namespace ptg
