// atomics_mb4.cu -- is the native 128-bit shared-memory CAS (ATOMS.CAS.128, sm_90+) a faster way to
// add to two ADJACENT doubles than two fp64 atomicAdds (each a 64-bit CAS loop)?
//  P64   two atomicAdd(double) on cells (2k, 2k+1), k random in a 64 KB tile
//  P128  one CAS.128 loop on the aligned pair (2k, 2k+1)
//  S64   one atomicAdd(double) on a random cell (reference, = S0 of atomics_mb2)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t xs32(uint32_t &s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

__device__ __forceinline__ void add_pair_cas128(double *p, double a, double b) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(p);
    double2 cur = *reinterpret_cast<double2 *>(p);
    while (true) {
        const unsigned long long c0 = __double_as_longlong(cur.x), c1 = __double_as_longlong(cur.y);
        const unsigned long long n0 = __double_as_longlong(cur.x + a), n1 = __double_as_longlong(cur.y + b);
        unsigned long long r0, r1;
        asm volatile("{\n\t.reg .b128 c, n, r;\n\tmov.b128 c, {%2,%3};\n\tmov.b128 n, {%4,%5};\n\t"
                     "atom.shared.cas.b128 r, [%6], c, n;\n\tmov.b128 {%0,%1}, r;\n\t}"
                     : "=l"(r0), "=l"(r1) : "l"(c0), "l"(c1), "l"(n0), "l"(n1), "r"(addr) : "memory");
        if (r0 == c0 && r1 == c1) break;
        cur.x = __longlong_as_double(r0);
        cur.y = __longlong_as_double(r1);
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) k_pairs(double *out, int tile_cells, int iters) {
    extern __shared__ __align__(16) double tile[];
    for (int i = threadIdx.x; i < tile_cells; i += blockDim.x) tile[i] = 0.0;
    __syncthreads();
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 777u;
    xs32(s);
    for (int i = 0; i < iters; ++i) {
        const uint32_t k = __umulhi(xs32(s), (uint32_t)tile_cells / 2);
        if (MODE == 0) { atomicAdd(&tile[2 * k], 1.0); atomicAdd(&tile[2 * k + 1], 2.0); }
        else if (MODE == 1) add_pair_cas128(&tile[2 * k], 1.0, 2.0);
        else atomicAdd(&tile[2 * k], 1.0);
    }
    __syncthreads();
    double acc = 0;
    for (int i = threadIdx.x; i < tile_cells; i += blockDim.x) acc += tile[i];
    if (threadIdx.x == 0 && blockIdx.x == 0) out[MODE] = acc;
}
template <class F> float timeit(F f, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
    CK(cudaGetLastError());
    return best;
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double *out; CK(cudaMalloc(&out, 64)); CK(cudaMemset(out, 0, 64));
    for (int tile_kb : {8, 64}) {
        const int tile_cells = tile_kb * 1024 / 8, bl = sms * 3, th = 256, iters = 1024;
        CK(cudaFuncSetAttribute(k_pairs<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_kb * 1024));
        CK(cudaFuncSetAttribute(k_pairs<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_kb * 1024));
        CK(cudaFuncSetAttribute(k_pairs<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_kb * 1024));
        const double u = (double)bl * th * iters;
        float ms = timeit([&] { k_pairs<0><<<bl, th, tile_kb * 1024>>>(out, tile_cells, iters); });
        printf("tile %2d KB  P64  two atomicAdd(double) per pair : %7.3f ms  %7.1f Gpairs/s\n", tile_kb, ms, u / ms * 1e-6);
        ms = timeit([&] { k_pairs<1><<<bl, th, tile_kb * 1024>>>(out, tile_cells, iters); });
        printf("tile %2d KB  P128 one CAS.128 loop per pair      : %7.3f ms  %7.1f Gpairs/s\n", tile_kb, ms, u / ms * 1e-6);
        ms = timeit([&] { k_pairs<2><<<bl, th, tile_kb * 1024>>>(out, tile_cells, iters); });
        printf("tile %2d KB  S64  one atomicAdd(double)          : %7.3f ms  %7.1f Gupd/s\n", tile_kb, ms, u / ms * 1e-6);
    }
    double h[3]; CK(cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost));
    printf("checksums (block 0): %g %g %g\n", h[0], h[1], h[2]);
    return 0;
}
