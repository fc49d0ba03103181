// atomics_mb2.cu -- second-generation micro-benchmark: what bounds fp64 RED traffic on B200?
//
// Address generation is one xorshift32 + one __umulhi per update (a few integer ops), so the
// kernels are bound by the memory path, not by index arithmetic (atomics_mb.cu was).
//
//  L<k>  groups of k adjacent lanes write k adjacent doubles starting at a random k*8-byte aligned
//        address (k = 1,2,4,8,16,32): k=4 is exactly one 32 B sector per group, k=16 one 128 B line.
//        If the update rate is constant in sectors/s the cost is per sector; if it grows with k up
//        to 16 it is per 128 B line request.
//  U3    3 of 4 lanes, unaligned start (a TSC stencil row in the plain C-order grid)
//  A3    3 of 4 lanes, sector-aligned start (a TSC row in the staggered-copy layout)
//  B9    9 of 16 lanes: a 3x3 sub-block of one 128 B line (4x4-cell block layout)
//  S*    shared memory: atomicAdd(double) (CAS loop) vs plain RMW vs u32 ATOMS, random in 64 KB
//  T*    TMA bulk reduce (.add.f64) of 32 B / 128 B from shared to random aligned global rows,
//        4 operations in flight per thread
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t xs32(uint32_t &s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
__device__ __forceinline__ void red_f64(double *a, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(a), "d"(v) : "memory"); }

// groups of K lanes, `active` of them write; start index aligned to ALIGN doubles
template <int K, int ALIGN>
__global__ void __launch_bounds__(256) k_red_group(double *grid, uint32_t cells, int iters, int active) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t g = t / K, j = t % K;
    uint32_t s = g * 2654435761u + 12345u;
    xs32(s);
    const uint32_t span = (cells - 64) / ALIGN;
    for (int i = 0; i < iters; ++i) {
        const uint32_t c = __umulhi(xs32(s), span) * ALIGN;
        if ((int)j < active) red_f64(grid + c + j, 1.0);
    }
}

// 3x3 sub-block of a 4x4-cell block stored as one 128 B line: lanes 0..15 = the block, 9 active
__global__ void __launch_bounds__(256) k_red_block9(double *grid, uint32_t cells, int iters) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t g = t / 16, j = t % 16;
    uint32_t s = g * 2654435761u + 12345u;
    xs32(s);
    const uint32_t span = (cells - 64) / 16;
    for (int i = 0; i < iters; ++i) {
        const uint32_t r = xs32(s);
        const uint32_t c = __umulhi(r, span) * 16;
        const uint32_t ox = r & 1, oy = (r >> 1) & 1;            // where the 3x3 window sits in the 4x4 block
        const uint32_t bx = j / 4, by = j % 4;
        if (bx >= ox && bx < ox + 3 && by >= oy && by < oy + 3) red_f64(grid + c + j, 1.0);
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) k_smem(double *out, int tile_cells, int iters) {
    extern __shared__ double tile[];
    for (int i = threadIdx.x; i < tile_cells; i += blockDim.x) tile[i] = 0.0;
    __syncthreads();
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 777u;
    xs32(s);
    for (int i = 0; i < iters; ++i) {
        const uint32_t c = __umulhi(xs32(s), (uint32_t)tile_cells);
        if (MODE == 0) atomicAdd(&tile[c], 1.0);
        else if (MODE == 1) { volatile double *p = tile + c; *p = *p + 1.0; }
        else atomicAdd(reinterpret_cast<unsigned *>(tile) + 2 * c, 1u);
    }
    __syncthreads();
    double acc = 0;
    for (int i = threadIdx.x; i < tile_cells; i += blockDim.x) acc += tile[i];
    if (acc == -1.0) out[0] = acc;
}

template <int ROWB, int DEPTH>
__global__ void __launch_bounds__(128) k_tma_reduce(double *grid, uint32_t cells, int iters) {
    extern __shared__ __align__(128) double stage[];
    double *mine = stage + threadIdx.x * (ROWB / 8) * DEPTH;
    for (int k = 0; k < ROWB / 8 * DEPTH; ++k) mine[k] = 1.0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 99u;
    xs32(s);
    const uint32_t span = (cells - 64) / (ROWB / 8);
    for (int i = 0; i < iters; ++i) {
        const uint32_t c = __umulhi(xs32(s), span) * (ROWB / 8);
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(mine + (i % DEPTH) * (ROWB / 8));
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                     ::"l"(grid + c), "r"(sa), "n"(ROWB) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <class F> float timeit(F f, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

template <int K, int ALIGN>
void run_group(const char *tag, double *grid, uint32_t cells, int sms, int active, double sectors_per_group) {
    const int threads = 256, blocks = sms * 8, iters = 512;
    float ms = timeit([&] { k_red_group<K, ALIGN><<<blocks, threads>>>(grid, cells, iters, active); });
    const double groups = (double)threads * blocks / K * iters;
    printf("%-34s: %7.3f ms  %8.1f Gupd/s  %7.1f Ggroups/s  %7.1f Gsectors/s\n", tag, ms,
           groups * active / ms * 1e-6, groups / ms * 1e-6, groups * sectors_per_group / ms * 1e-6);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("device %s  SMs %d\n", p.name, sms);
    const uint32_t cells2d = 1028u * 1028u, cells3d = 258u * 258u * 258u, cells4 = 4u * 1028u * 1028u;
    double *grid; CK(cudaMalloc(&grid, (size_t)cells3d * 8)); CK(cudaMemset(grid, 0, (size_t)cells3d * 8));

    printf("--- 8.45 MB grid (1028^2) ---\n");
    run_group<1, 1>("L1  1 lane  / random double", grid, cells2d, sms, 1, 1);
    run_group<2, 2>("L2  2 lanes / 16 B aligned", grid, cells2d, sms, 2, 1);
    run_group<4, 4>("L4  4 lanes / one 32 B sector", grid, cells2d, sms, 4, 1);
    run_group<8, 8>("L8  8 lanes / 64 B", grid, cells2d, sms, 8, 2);
    run_group<16, 16>("L16 16 lanes / one 128 B line", grid, cells2d, sms, 16, 4);
    run_group<32, 32>("L32 32 lanes / 256 B", grid, cells2d, sms, 32, 8);
    run_group<4, 1>("U3  3 of 4 lanes, unaligned", grid, cells2d, sms, 3, 1.5);
    run_group<4, 4>("A3  3 of 4 lanes, sector aligned", grid, cells2d, sms, 3, 1);
    run_group<2, 1>("U2  2 lanes, unaligned (CIC row)", grid, cells2d, sms, 2, 1.25);
    run_group<4, 1>("U4  4 lanes, unaligned (spline row)", grid, cells2d, sms, 4, 1.75);
    {
        const int threads = 256, blocks = sms * 8, iters = 512;
        float ms = timeit([&] { k_red_block9<<<blocks, threads>>>(grid, cells2d, iters); });
        const double groups = (double)threads * blocks / 16 * iters;
        printf("%-34s: %7.3f ms  %8.1f Gupd/s  %7.1f Ggroups/s  %7.1f Gsectors/s\n", "B9  3x3 of a 4x4 block (one line)", ms,
               groups * 9 / ms * 1e-6, groups / ms * 1e-6, groups * 3 / ms * 1e-6);
    }
    printf("--- 33.8 MB (4 staggered copies of 1028^2) ---\n");
    run_group<4, 4>("A3  3 of 4 lanes, sector aligned", grid, cells4, sms, 3, 1);
    run_group<16, 16>("L16 16 lanes / one 128 B line", grid, cells4, sms, 16, 4);
    printf("--- 137 MB grid (258^3) ---\n");
    run_group<1, 1>("L1  1 lane  / random double", grid, cells3d, sms, 1, 1);
    run_group<2, 1>("U2  2 lanes, unaligned (CIC row)", grid, cells3d, sms, 2, 1.25);
    run_group<4, 4>("L4  4 lanes / one 32 B sector", grid, cells3d, sms, 4, 1);
    run_group<16, 16>("L16 16 lanes / one 128 B line", grid, cells3d, sms, 16, 4);

    printf("--- shared memory, 64 KB tile, 3 CTAs/SM x 256 threads ---\n");
    {
        const int tile_kb = 64, tile_cells = tile_kb * 1024 / 8, bl = sms * 3, th = 256, iters = 1024;
        CK(cudaFuncSetAttribute(k_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_kb * 1024));
        CK(cudaFuncSetAttribute(k_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_kb * 1024));
        CK(cudaFuncSetAttribute(k_smem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_kb * 1024));
        const double u = (double)bl * th * iters;
        float ms = timeit([&] { k_smem<0><<<bl, th, tile_kb * 1024>>>(grid, tile_cells, iters); });
        printf("S0  smem atomicAdd(double) CAS loop : %7.3f ms  %8.1f Gupd/s\n", ms, u / ms * 1e-6);
        ms = timeit([&] { k_smem<1><<<bl, th, tile_kb * 1024>>>(grid, tile_cells, iters); });
        printf("S1  smem plain load+add+store       : %7.3f ms  %8.1f Gupd/s\n", ms, u / ms * 1e-6);
        ms = timeit([&] { k_smem<2><<<bl, th, tile_kb * 1024>>>(grid, tile_cells, iters); });
        printf("S2  smem ATOMS.ADD u32              : %7.3f ms  %8.1f Gupd/s\n", ms, u / ms * 1e-6);
    }
    printf("--- TMA bulk reduce, 4 in flight per thread ---\n");
    {
        const int th = 128, bl = sms * 8, iters = 256;
        const double ops = (double)th * bl * iters;
        float ms = timeit([&] { k_tma_reduce<32, 4><<<bl, th, th * 32 * 4>>>(grid, cells2d, iters); });
        printf("T32  32 B rows  : %7.3f ms  %7.1f Gops/s  %7.1f Gsectors/s\n", ms, ops / ms * 1e-6, ops / ms * 1e-6);
        CK(cudaFuncSetAttribute(k_tma_reduce<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, th * 128 * 4));
        ms = timeit([&] { k_tma_reduce<128, 4><<<bl / 2, th, th * 128 * 4>>>(grid, cells2d, iters); });
        printf("T128 128 B rows : %7.3f ms  %7.1f Gops/s  %7.1f Gsectors/s\n", ms, ops / 2 / ms * 1e-6, ops / 2 * 4 / ms * 1e-6);
    }
    return 0;
}
