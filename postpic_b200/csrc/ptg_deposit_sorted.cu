// ptg_deposit_sorted.cu -- method SORTED: deterministic deposition without atomics.
//
//   1. k_cell_keys      key[i] = padded linear index of particle i's stencil base cell
//                       (rejected particles get the sentinel key `cells`)
//   2. stable LSD radix sort of (key, particle index) pairs -- hand-written, 8-bit digits,
//      match_any-based stable ranking (ptg_radix_sort.cuh); only the bits a key can have are
//      sorted (3 passes for the 1028^2 grid)
//   3. k_cell_offsets   CSR offsets: particles of base cell c are idx[off[c] .. off[c+1])
//                       in ORIGINAL particle order (the sort is stable)
//   4. k_segment_reduce one launch per stencil "colour" (base cells with equal coordinates
//      modulo the stencil width): stencils of one colour never overlap, so one warp per
//      base cell sums its segment in registers (lane l takes particles l, l+32, ... in
//      order, then a fixed xor-butterfly) and adds the (S+1)^ndim totals to the grid with
//      plain loads and stores.  Colours run one after the other on the stream, so every
//      grid cell receives its (S+1)^ndim contributions in a fixed order.
// No floating-point atomic anywhere: the result is bit-reproducible from run to run
// and independent of the hardware's scheduling; it is the variant for clustered inputs
// where atomic contention is worst.  (It is NOT the reference's summation order -- that
// one is strictly sequential -- so parity is within rounding, not bit-exact.)
//
// Replaces the loops of /root/reference/postpic/particles/_particlestogrid.pyx
// :70-129 (1-D), :181-276 (2-D), :341-463 (3-D).
#include "ptg_internal.cuh"
#include "ptg_radix_sort.cuh"

namespace ptg {

template <int ND, int S>
__global__ void __launch_bounds__(256) k_cell_keys(const DepositParams P, uint32_t cells, uint32_t *keys,
                                                    uint32_t *idx) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += step) {
        const double x = particle_coord_rt(P, 0, i);
        const double y = ND >= 2 ? particle_coord_rt(P, 1, i) : 0.0;
        const double z = ND >= 3 ? particle_coord_rt(P, 2, i) : 0.0;
        int b[3];
        double wx[StencilWidth<S>::value], wy[StencilWidth<S>::value], wz[StencilWidth<S>::value];
        const bool ok = particle_stencil<ND, S>(P, x, y, z, b, wx, wy, wz) && particle_selected(P, i);
        uint32_t key = cells;
        if (ok) key = (uint32_t)((b[0] * P.dims[1] + b[1]) * P.dims[2] + b[2]);    // dims beyond ndim are 1
        keys[i] = key;
        idx[i] = (uint32_t)i;
    }
}

// off[c] = first position in the sorted key array whose key is >= c, for c in [0, cells]
__global__ void __launch_bounds__(256) k_cell_offsets(const uint32_t *keys, int64_t n, uint32_t cells,
                                                       uint32_t *off) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += step) {
        const uint32_t lo = i == 0 ? 0u : keys[i - 1] + 1u;         // first key value that starts at >= i
        const uint32_t hi = i == n ? cells : keys[i];
        if (i == 0 || i == n || keys[i - 1] != hi)
            for (uint32_t c = lo; c <= hi && c <= cells; ++c) off[c] = (uint32_t)i;
    }
}

constexpr uint32_t kLongSegment = 2048;     // segments longer than this get a whole CTA instead of a warp

// LONG = false: one warp per base cell, segments of up to kLongSegment particles.
// LONG = true : one CTA (8 warps) per base cell for the longer segments of clustered input: the
//               segment is cut into 8 equal contiguous parts, one per warp, and the 8 partial
//               totals are added in warp order -- still a fixed summation order.
template <int ND, int S, bool W, bool LONG, bool F>
__global__ void __launch_bounds__(256) k_segment_reduce(const DepositParams P, const uint32_t *idx,
                                                         const uint32_t *off, int cx, int cy, int cz) {
    constexpr int NW = StencilWidth<S>::value;
    constexpr int NA = NW, NC = ND >= 2 ? NW : 1, NDD = ND >= 3 ? NW : 1;
    constexpr int NV = NA * NC * NDD;
    __shared__ double s_part[LONG ? 8 * NV : 1];
    // base cells of this colour: b = colour + NW * k, with b + NW <= dims
    const int nx = (P.dims[0] - NW - cx) / NW + 1;
    const int ny = ND >= 2 ? (P.dims[1] - NW - cy) / NW + 1 : 1;
    const int nz = ND >= 3 ? (P.dims[2] - NW - cz) / NW + 1 : 1;
    if (nx <= 0 || ny <= 0 || nz <= 0) return;
    const int64_t ncell = (int64_t)nx * ny * nz;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nworkers = LONG ? (int64_t)gridDim.x : (int64_t)gridDim.x * 8;
    const int64_t worker = LONG ? (int64_t)blockIdx.x : (int64_t)blockIdx.x * 8 + warp;
    for (int64_t cc = worker; cc < ncell; cc += nworkers) {
        const int kz = (int)(cc % nz), ky = (int)((cc / nz) % ny), kx = (int)(cc / ((int64_t)nz * ny));
        const int bx = cx + NW * kx, by = cy + NW * ky, bz = cz + NW * kz;
        const uint32_t lin = (uint32_t)((bx * P.dims[1] + by) * P.dims[2] + bz);
        uint32_t s0 = off[lin], s1 = off[lin + 1];
        const uint32_t len = s1 - s0;
        if (len == 0 || (len > kLongSegment) != LONG) continue;          // warp- (CTA-) uniform
        if (LONG) {                                                       // this warp's part of the segment
            const uint32_t part = (len + 7) / 8;
            s0 = s0 + warp * part;
            s1 = s0 + part < s1 ? s0 + part : s1;
        }
        double acc[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] = 0.0;
        for (uint32_t p = s0 + lane; p < s1; p += 32) {
            const int64_t i = idx[p];
            const double x = particle_coord<F>(P, 0, i);
            const double y = ND >= 2 ? particle_coord<F>(P, 1, i) : 0.0;
            const double z = ND >= 3 ? particle_coord<F>(P, 2, i) : 0.0;
            const double w = W ? particle_weight<F>(P, i) : 1.0;
            double wx[NW], wy[NW], wz[NW];
            int b[3];
            particle_stencil<ND, S>(P, x, y, z, b, wx, wy, wz);      // accepted already: it has a cell key
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int d = 0; d < NDD; ++d) {
                        acc[(a * NC + c) * NDD + d] = __dadd_rn(acc[(a * NC + c) * NDD + d],
                                                                stencil_value<ND, S, W>(wx, wy, wz, a, c, d, w));
                    }
        }
        // fixed-order butterfly: every lane ends with the same total
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int k = 0; k < NV; ++k) acc[k] = __dadd_rn(acc[k], __shfl_xor_sync(0xffffffffu, acc[k], o));
        const int64_t g1 = ND >= 2 ? P.dims[1] : 1, g2 = ND >= 3 ? P.dims[2] : 1;
        double *g = P.grid + ((int64_t)bx * g1 + by) * g2 + bz;
        if (!LONG) {
            if (lane == 0) {
#pragma unroll
                for (int a = 0; a < NA; ++a)
#pragma unroll
                    for (int c = 0; c < NC; ++c)
#pragma unroll
                        for (int d = 0; d < NDD; ++d) {
                            double *q = g + (a * g1 + c) * g2 + d;
                            *q = __dadd_rn(*q, acc[(a * NC + c) * NDD + d]);
                        }
            }
        } else {
            __syncthreads();                 // previous cell's partials have been consumed
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < NV; ++k) s_part[warp * NV + k] = acc[k];
            }
            __syncthreads();
            if (threadIdx.x < NV) {
                const int k = threadIdx.x;
                double t = s_part[k];
                for (int wv = 1; wv < 8; ++wv) t = __dadd_rn(t, s_part[wv * NV + k]);
                const int d = k % NDD, c = (k / NDD) % NC, a = k / (NDD * NC);
                double *q = g + (a * g1 + c) * g2 + d;
                *q = __dadd_rn(*q, t);
            }
        }
    }
}

static int key_bits(uint32_t cells) {
    int b = 1;
    while (b < 32 && (cells >> b) != 0) ++b;
    return b;
}

size_t sorted_workspace_bytes(int64_t n, int64_t cells) {
    // 4 arrays of n u32 (double-buffered keys and indices), the CSR offsets, the sort's digit counts
    auto al = [](size_t b) { return (b + 255) / 256 * 256; };
    return 4 * al((size_t)n * 4) + al((size_t)(cells + 2) * 4) + al(radix_sort_scratch_bytes(n));
}

template <int ND, int S, bool W>
static cudaError_t run_sorted(const DepositParams &P, void *ws, size_t ws_bytes, int sm_count, cudaStream_t st) {
    const int64_t cells64 = (int64_t)P.dims[0] * P.dims[1] * P.dims[2];
    const uint32_t cells = (uint32_t)cells64;
    auto al = [](size_t b) { return (b + 255) / 256 * 256; };
    char *base = static_cast<char *>(ws);
    uint32_t *k0 = reinterpret_cast<uint32_t *>(base); base += al((size_t)P.n * 4);
    uint32_t *k1 = reinterpret_cast<uint32_t *>(base); base += al((size_t)P.n * 4);
    uint32_t *i0 = reinterpret_cast<uint32_t *>(base); base += al((size_t)P.n * 4);
    uint32_t *i1 = reinterpret_cast<uint32_t *>(base); base += al((size_t)P.n * 4);
    uint32_t *off = reinterpret_cast<uint32_t *>(base); base += al((size_t)(cells + 2) * 4);
    void *sort_scratch = base;
    if ((size_t)(base - static_cast<char *>(ws)) + radix_sort_scratch_bytes(P.n) > ws_bytes) return cudaErrorMemoryAllocation;

    const int64_t cap = (int64_t)sm_count * 8;
    int64_t want = (P.n + 256 * 4 - 1) / (256 * 4);
    int blocks = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
    k_cell_keys<ND, S><<<blocks, 256, 0, st>>>(P, cells, k0, i0);

    uint32_t *ks = nullptr, *is = nullptr;       // whichever buffer the last pass wrote
    cudaError_t e = radix_sort_pairs(k0, k1, i0, i1, P.n, key_bits(cells), sort_scratch, st, &ks, &is);
    if (e != cudaSuccess) return e;

    want = (P.n + 1 + 255) / 256;
    blocks = (int)(want < cap ? want : cap);
    k_cell_offsets<<<blocks, 256, 0, st>>>(ks, P.n, cells, off);

    constexpr int NW = StencilWidth<S>::value;
    const int ncx = NW, ncy = ND >= 2 ? NW : 1, ncz = ND >= 3 ? NW : 1;
    for (int cx = 0; cx < ncx; ++cx)
        for (int cy = 0; cy < ncy; ++cy)
            for (int cz = 0; cz < ncz; ++cz) {
                const int64_t ncell = ((int64_t)(P.dims[0] - NW - cx) / NW + 1) *
                                      (ND >= 2 ? (P.dims[1] - NW - cy) / NW + 1 : 1) *
                                      (ND >= 3 ? (P.dims[2] - NW - cz) / NW + 1 : 1);
                if (ncell <= 0) continue;
                int64_t wantb = (ncell + 7) / 8;                  // 8 warps per CTA
                int nb = (int)(wantb < cap * 4 ? wantb : cap * 4);
                if (W && fused_mode(P)) k_segment_reduce<ND, S, W, false, W><<<nb, 256, 0, st>>>(P, is, off, cx, cy, cz);
                else k_segment_reduce<ND, S, W, false, false><<<nb, 256, 0, st>>>(P, is, off, cx, cy, cz);
                // the long segments of the same colour (clustered input), one CTA each
                nb = (int)(ncell < cap ? ncell : cap);
                if (W && fused_mode(P)) k_segment_reduce<ND, S, W, true, W><<<nb, 256, 0, st>>>(P, is, off, cx, cy, cz);
                else k_segment_reduce<ND, S, W, true, false><<<nb, 256, 0, st>>>(P, is, off, cx, cy, cz);
            }
    return cudaGetLastError();
}

template <int ND, int S>
static cudaError_t sorted_w(const DepositParams &P, void *ws, size_t wsb, int sm, cudaStream_t st) {
    return (P.p[3] || fused_mode(P)) ? run_sorted<ND, S, true>(P, ws, wsb, sm, st) : run_sorted<ND, S, false>(P, ws, wsb, sm, st);
}

template <int ND>
static cudaError_t sorted_s(int shape, const DepositParams &P, void *ws, size_t wsb, int sm, cudaStream_t st) {
    switch (shape) {
        case 0: return sorted_w<ND, 0>(P, ws, wsb, sm, st);
        case 1: return sorted_w<ND, 1>(P, ws, wsb, sm, st);
        case 2: return sorted_w<ND, 2>(P, ws, wsb, sm, st);
        default: return sorted_w<ND, 3>(P, ws, wsb, sm, st);
    }
}

__global__ void k_iota(uint32_t *v, int64_t n) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) v[i] = (uint32_t)i;
}

// test hook: keys[0..n) sorted in place by their low `bits` bits (stable), perm[i] = original index
cudaError_t selftest_radix_sort(int64_t n, int bits, uint32_t *keys, uint32_t *perm, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    uint32_t *k1 = nullptr, *v0 = nullptr;
    void *scratch = nullptr;
    cudaError_t e = scratch_alloc(reinterpret_cast<void **>(&k1), (size_t)n * 4, st);
    if (e == cudaSuccess) e = scratch_alloc(reinterpret_cast<void **>(&v0), (size_t)n * 4, st);
    if (e == cudaSuccess) e = scratch_alloc(&scratch, radix_sort_scratch_bytes(n), st);
    if (e != cudaSuccess) return e;
    k_iota<<<1024, 256, 0, st>>>(v0, n);
    uint32_t *ks = nullptr, *vs = nullptr;
    e = radix_sort_pairs(keys, k1, v0, perm, n, bits, scratch, st, &ks, &vs);
    if (e == cudaSuccess && ks != keys) e = cudaMemcpyAsync(keys, ks, (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess && vs != perm) e = cudaMemcpyAsync(perm, vs, (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
    cudaFreeAsync(k1, st);
    cudaFreeAsync(v0, st);
    cudaFreeAsync(scratch, st);
    return e;
}

cudaError_t deposit_sorted(int ndim, int shape, const DepositParams &P, void *ws, size_t ws_bytes, int sm_count,
                           cudaStream_t st) {
    if (P.n <= 0) return cudaSuccess;
    switch (ndim) {
        case 1: return sorted_s<1>(shape, P, ws, ws_bytes, sm_count, st);
        case 2: return sorted_s<2>(shape, P, ws, ws_bytes, sm_count, st);
        default: return sorted_s<3>(shape, P, ws, ws_bytes, sm_count, st);
    }
}

}  // namespace ptg
