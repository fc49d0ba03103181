// atomics_mb3.cu -- third micro-benchmark: is the ~190 G sector-RED/s limit an L2 limit or an
// SM->L2 request limit, do dense 3-lane groups beat 3-of-4, and do TMA bulk reductions add to it?
//  SMS<k>  pattern L1 (one random double per lane) launched on k CTAs of 1024 threads (1 per SM)
//  D3      dense 3-lane groups: lanes 0..29 form 10 groups of 3 adjacent doubles (sector aligned)
//  D3U     the same, unaligned (plain C-order grid)
//  MIX     half of the warps issue REDs (U3 pattern), the other half TMA bulk reductions of 32 B
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t xs32(uint32_t &s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
__device__ __forceinline__ void red_f64(double *a, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(a), "d"(v) : "memory"); }

__global__ void __launch_bounds__(1024) k_l1(double *grid, uint32_t cells, int iters) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u; xs32(s);
    for (int i = 0; i < iters; ++i) red_f64(grid + __umulhi(xs32(s), cells), 1.0);
}
template <int ALIGN>
__global__ void __launch_bounds__(256) k_dense3(double *grid, uint32_t cells, int iters) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp = t >> 5;
    const uint32_t g = warp * 10 + lane / 3, j = lane % 3;
    uint32_t s = g * 2654435761u + 12345u; xs32(s);
    const uint32_t span = (cells - 64) / ALIGN;
    for (int i = 0; i < iters; ++i) {
        const uint32_t c = __umulhi(xs32(s), span) * ALIGN;
        if (lane < 30) red_f64(grid + c + j, 1.0);
    }
}
// MODE 0: all warps RED (U3); MODE 1: all warps TMA 32 B; MODE 2: even warps RED, odd warps TMA
template <int MODE>
__global__ void __launch_bounds__(256) k_mix(double *grid, uint32_t cells, int iters) {
    extern __shared__ __align__(128) double stage[];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const int warp = threadIdx.x >> 5;
    const bool tma = MODE == 1 || (MODE == 2 && (warp & 1));
    uint32_t s = (t / 4) * 2654435761u + 99u; xs32(s);
    if (!tma) {
        const uint32_t j = t & 3;
        for (int i = 0; i < iters; ++i) {
            const uint32_t c = __umulhi(xs32(s), cells - 64);
            if (j < 3) red_f64(grid + c + j, 1.0);
        }
    } else {
        // every 4th lane issues one 32-byte bulk reduction per iteration (same rows/s per warp as the RED warps)
        double *mine = stage + (threadIdx.x / 4) * 4 * 4;
        if ((t & 3) == 0) for (int k = 0; k < 16; ++k) mine[k] = 1.0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if ((t & 3) == 0) {
            for (int i = 0; i < iters; ++i) {
                const uint32_t c = __umulhi(xs32(s), (cells - 64) / 2) * 2;
                const uint32_t sa = (uint32_t)__cvta_generic_to_shared(mine + (i & 3) * 4);
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 32;" ::"l"(grid + c), "r"(sa) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }
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
    const uint32_t cells = 1028u * 1028u;
    double *grid; CK(cudaMalloc(&grid, (size_t)cells * 8 * 4)); CK(cudaMemset(grid, 0, (size_t)cells * 8 * 4));
    for (int k : {18, 37, 74, 111, 148, 296}) {
        const int iters = 2048;
        float ms = timeit([&] { k_l1<<<k, 1024>>>(grid, cells, iters); });
        printf("SMS%-3d L1 pattern on %3d CTAs x 1024 thr : %7.3f ms  %7.1f Gsectors/s  (%5.3f per clk per CTA @1.965GHz)\n", k, k, ms,
               (double)k * 1024 * iters / ms * 1e-6, (double)1024 * iters / ms * 1e-6 / 1.965);
    }
    {
        const int threads = 256, blocks = sms * 8, iters = 512;
        const double groups = (double)threads / 32 * 10 * blocks * iters;
        float ms = timeit([&] { k_dense3<4><<<blocks, threads>>>(grid, cells, iters); });
        printf("D3   dense 3-lane groups, sector aligned  : %7.3f ms  %7.1f Ggroups/s  %7.1f Gupd/s\n", ms, groups / ms * 1e-6, groups * 3 / ms * 1e-6);
        ms = timeit([&] { k_dense3<1><<<blocks, threads>>>(grid, cells, iters); });
        printf("D3U  dense 3-lane groups, unaligned       : %7.3f ms  %7.1f Ggroups/s  %7.1f Gupd/s\n", ms, groups / ms * 1e-6, groups * 3 / ms * 1e-6);
    }
    {
        const int threads = 256, blocks = sms * 8, iters = 512;
        const double rows = (double)threads / 4 * blocks * iters;
        float ms = timeit([&] { k_mix<0><<<blocks, threads, 8192>>>(grid, cells, iters); });
        printf("MIX0 all warps RED rows (U3)              : %7.3f ms  %7.1f Grows/s\n", ms, rows / ms * 1e-6);
        ms = timeit([&] { k_mix<1><<<blocks, threads, 8192>>>(grid, cells, iters); });
        printf("MIX1 all warps TMA 32 B rows              : %7.3f ms  %7.1f Grows/s\n", ms, rows / ms * 1e-6);
        ms = timeit([&] { k_mix<2><<<blocks, threads, 8192>>>(grid, cells, iters); });
        printf("MIX2 half RED, half TMA                   : %7.3f ms  %7.1f Grows/s\n", ms, rows / ms * 1e-6);
    }
    return 0;
}
