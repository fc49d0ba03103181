// atomics_mb.cu -- B200 micro-benchmarks that decide the deposit kernel design.
//
// Measures, with CUDA events, the update rates of the candidate accumulation paths
// for fp64 scatter-add (the inner operation of particle-to-grid deposition):
//   A  RED.E.ADD.F64 to random addresses in an L2-resident 8.45 MB grid (1 lane = 1 sector)
//   B  the same updates regrouped so 3 of every 4 adjacent lanes hit adjacent doubles
//      (sector-coalesced: what a TSC stencil row looks like)
//   C  4 of 4 lanes on a 32-byte aligned sector (what an aligned cubic-spline row looks like)
//   D  shared-memory atomicAdd(double) (lowered to a CAS loop on sm_100a), random in a tile
//   E  shared-memory plain load+add+store, random in a tile (bank-limit reference; racy)
//   F  RED.F32 random, and red.global.add.v2.f32
//   G  cp.reduce.async.bulk (TMA) .add.f64 of 32-byte rows from shared to random global rows
//   H  A on a 137 MB grid (the 258^3 3-D CIC grid, larger than L2)
//   I  streaming read of three f64 arrays (the load side alone)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo atomics_mb.cu -o atomics_mb
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ void red_f64(double *a, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void red_f32(float *a, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void red_v2f32(float *a, float v, float w) { asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(a), "f"(v), "f"(w) : "memory"); }

// A / H: every lane one random cell
__global__ void k_red_random(double *grid, uint64_t cells, int iters) {
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (int i = 0; i < iters; ++i) {
        uint64_t c = mix(t * 1315423911ull + i) % cells;
        red_f64(grid + c, 1.0);
    }
}
// B / C: groups of 4 lanes share a random row start; lanes j < active write cell start + j
__global__ void k_red_grouped(double *grid, uint64_t cells, int iters, int active, int aligned) {
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t g = t >> 2; int j = t & 3;
    for (int i = 0; i < iters; ++i) {
        uint64_t c = mix(g * 1315423911ull + i) % (cells - 8);
        if (aligned) c &= ~3ull;
        if (j < active) red_f64(grid + c + j, 1.0);
    }
}
// 8 lanes per group (two sectors when aligned to 64 B)
__global__ void k_red_grouped8(double *grid, uint64_t cells, int iters) {
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t g = t >> 3; int j = t & 7;
    for (int i = 0; i < iters; ++i) {
        uint64_t c = (mix(g * 1315423911ull + i) % (cells - 16)) & ~7ull;
        red_f64(grid + c + j, 1.0);
    }
}
// D: shared atomics (CAS loop); E: plain RMW
template <int MODE>
__global__ void k_smem(double *out, int tile_cells, int iters) {
    extern __shared__ double tile[];
    for (int i = threadIdx.x; i < tile_cells; i += blockDim.x) tile[i] = 0.0;
    __syncthreads();
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (int i = 0; i < iters; ++i) {
        uint32_t c = (uint32_t)(mix(t * 1315423911ull + i) % (uint64_t)tile_cells);
        if (MODE == 0) atomicAdd(&tile[c], 1.0);
        else if (MODE == 1) { double v = tile[c]; tile[c] = v + 1.0; }
        else { // MODE 2: native 32-bit integer atomics on the two halves (cost reference)
            atomicAdd(reinterpret_cast<unsigned *>(&tile[c]), 1u);
        }
    }
    __syncthreads();
    double s = 0; for (int i = threadIdx.x; i < tile_cells; i += blockDim.x) s += tile[i];
    if (s == -1.0) out[0] = s;
}
// F
__global__ void k_red_f32(float *grid, uint64_t cells, int iters, int v2) {
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (int i = 0; i < iters; ++i) {
        uint64_t c = mix(t * 1315423911ull + i) % (cells - 2);
        if (v2) red_v2f32(grid + (c & ~1ull), 1.f, 1.f); else red_f32(grid + c, 1.f);
    }
}
// G: TMA bulk reduce, one op of ROWB bytes per thread per iteration
template <int ROWB>
__global__ void k_tma_reduce(double *grid, uint64_t cells, int iters) {
    extern __shared__ __align__(128) double stage[];
    double *mine = stage + threadIdx.x * (ROWB / 8);
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (int i = 0; i < iters; ++i) {
        uint64_t c = (mix(t * 1315423911ull + i) % (cells - 16)) & ~1ull;   // 16-byte aligned
        for (int k = 0; k < ROWB / 8; ++k) mine[k] = 1.0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        uint32_t s = (uint32_t)__cvta_generic_to_shared(mine);
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                     ::"l"(grid + c), "r"(s), "n"(ROWB) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// I: streaming read
__global__ void k_stream3(const double2 *a, const double2 *b, const double2 *c, uint64_t n2, double *out) {
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, st = (uint64_t)gridDim.x * blockDim.x;
    double s = 0;
    for (uint64_t i = t; i < n2; i += st) {
        double2 x = __ldg(a + i), y = __ldg(b + i), z = __ldg(c + i);
        s += x.x + x.y + y.x + y.y + z.x + z.y;
    }
    if (s == -1.0) out[0] = s;
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

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s  SMs %d  L2 %d MB  smem/SM %zu KB  clock %d MHz\n", p.name, p.multiProcessorCount,
           p.l2CacheSize >> 20, p.sharedMemPerMultiprocessor >> 10, p.clockRate / 1000);
    const int sms = p.multiProcessorCount;
    const uint64_t cells2d = 1028ull * 1028ull, cells3d = 258ull * 258ull * 258ull;
    double *grid; CK(cudaMalloc(&grid, cells3d * 8)); CK(cudaMemset(grid, 0, cells3d * 8));
    const int threads = 256, blocks = sms * 8, iters = 256;
    const double upd = (double)threads * blocks * iters;

    float ms;
    ms = timeit([&] { k_red_random<<<blocks, threads>>>(grid, cells2d, iters); });
    printf("A  RED.F64 random 8.45MB          : %8.3f ms  %7.2f Gupd/s\n", ms, upd / ms * 1e-6);
    ms = timeit([&] { k_red_grouped<<<blocks, threads>>>(grid, cells2d, iters, 3, 0); });
    printf("B  RED.F64 3-of-4 lanes adjacent  : %8.3f ms  %7.2f Gupd/s  %7.2f Grows/s\n", ms, upd * 0.75 / ms * 1e-6, upd / 4 / ms * 1e-6);
    ms = timeit([&] { k_red_grouped<<<blocks, threads>>>(grid, cells2d, iters, 4, 0); });
    printf("B4 RED.F64 4-of-4 lanes unaligned : %8.3f ms  %7.2f Gupd/s  %7.2f Grows/s\n", ms, upd / ms * 1e-6, upd / 4 / ms * 1e-6);
    ms = timeit([&] { k_red_grouped<<<blocks, threads>>>(grid, cells2d, iters, 4, 1); });
    printf("C  RED.F64 4-of-4 lanes aligned   : %8.3f ms  %7.2f Gupd/s  %7.2f Grows/s\n", ms, upd / ms * 1e-6, upd / 4 / ms * 1e-6);
    ms = timeit([&] { k_red_grouped8<<<blocks, threads>>>(grid, cells2d, iters); });
    printf("C8 RED.F64 8 lanes 64B aligned    : %8.3f ms  %7.2f Gupd/s\n", ms, upd / ms * 1e-6);
    ms = timeit([&] { k_red_random<<<blocks, threads>>>(grid, cells3d, iters); });
    printf("H  RED.F64 random 137MB           : %8.3f ms  %7.2f Gupd/s\n", ms, upd / ms * 1e-6);
    ms = timeit([&] { k_red_grouped<<<blocks, threads>>>(grid, cells3d, iters, 2, 0); });
    printf("H2 RED.F64 2-of-4 adjacent 137MB  : %8.3f ms  %7.2f Gupd/s  %7.2f Grows/s\n", ms, upd * 0.5 / ms * 1e-6, upd / 4 / ms * 1e-6);
    ms = timeit([&] { k_red_f32<<<blocks, threads>>>((float *)grid, cells2d * 2, iters, 0); });
    printf("F  RED.F32 random                 : %8.3f ms  %7.2f Gupd/s\n", ms, upd / ms * 1e-6);
    ms = timeit([&] { k_red_f32<<<blocks, threads>>>((float *)grid, cells2d * 2, iters, 1); });
    printf("F2 RED.v2.F32 random              : %8.3f ms  %7.2f Gop/s\n", ms, upd / ms * 1e-6);

    for (int tile_kb : {16, 64, 200}) {
        int tile_cells = tile_kb * 1024 / 8;
        CK(cudaFuncSetAttribute(k_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_kb * 1024));
        CK(cudaFuncSetAttribute(k_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_kb * 1024));
        CK(cudaFuncSetAttribute(k_smem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_kb * 1024));
        int bl = sms * (tile_kb <= 16 ? 8 : tile_kb <= 64 ? 3 : 1), th = tile_kb <= 64 ? 256 : 1024;
        double u2 = (double)bl * th * iters;
        ms = timeit([&] { k_smem<0><<<bl, th, tile_kb * 1024>>>(grid, tile_cells, iters); });
        printf("D  smem atomicAdd f64 tile %3d KB  : %8.3f ms  %7.2f Gupd/s\n", tile_kb, ms, u2 / ms * 1e-6);
        ms = timeit([&] { k_smem<1><<<bl, th, tile_kb * 1024>>>(grid, tile_cells, iters); });
        printf("E  smem plain RMW f64 tile %3d KB  : %8.3f ms  %7.2f Gupd/s\n", tile_kb, ms, u2 / ms * 1e-6);
        ms = timeit([&] { k_smem<2><<<bl, th, tile_kb * 1024>>>(grid, tile_cells, iters); });
        printf("E2 smem ATOMS.ADD u32 tile %3d KB  : %8.3f ms  %7.2f Gupd/s\n", tile_kb, ms, u2 / ms * 1e-6);
    }
    {
        const int it2 = 64;
        double u3 = (double)threads * blocks * it2;
        ms = timeit([&] { k_tma_reduce<32><<<blocks, threads, threads * 32>>>(grid, cells2d, it2); });
        printf("G  TMA bulk reduce f64 32B rows   : %8.3f ms  %7.2f Grows/s\n", ms, u3 / ms * 1e-6);
        ms = timeit([&] { k_tma_reduce<64><<<blocks, threads, threads * 64>>>(grid, cells2d, it2); });
        printf("G2 TMA bulk reduce f64 64B rows   : %8.3f ms  %7.2f Grows/s\n", ms, u3 / ms * 1e-6);
    }
    {
        const uint64_t n = 1ull << 27;   // 3 x 1 GiB
        double *a, *b, *c; CK(cudaMalloc(&a, n * 8)); CK(cudaMalloc(&b, n * 8)); CK(cudaMalloc(&c, n * 8));
        CK(cudaMemset(a, 0, n * 8)); CK(cudaMemset(b, 0, n * 8)); CK(cudaMemset(c, 0, n * 8));
        for (int mult : {4, 8, 16, 32}) {
            ms = timeit([&] { k_stream3<<<sms * mult, 256>>>((double2 *)a, (double2 *)b, (double2 *)c, n / 2, grid); });
            printf("I  stream 3xf64 read grid=%2dxSMs   : %8.3f ms  %7.1f GB/s\n", mult, ms, 3.0 * n * 8 / ms * 1e-6);
        }
    }
    return 0;
}
