// atomics_mb5.cu -- round-2 micro-benchmark: what can replace the fp64 CAS loop for shared-memory
// privatised tiles, do shared atomics and L2 reductions add up when driven from one CTA, what does
// routing particles through L2-resident queues cost, and what does a grid-wide barrier cost?
//
//  F*  TSC-pattern accumulation (3 rows x 3 adjacent columns per particle, 9 updates) into a 9-row x
//      1040-column shared-memory band, one persistent CTA of 1024 threads per SM:
//        cas    atomicAdd(double)            (ATOMS.CAST.SPIN.64 loop)
//        plain  load + add + store           (racy, upper bound of a conflict-free exact scheme)
//        fx2    64-bit fixed point on two native ATOMS.ADD.U32: low word returning (carry), high word not
//        fx1    low word only                (upper bound: one native atomic per update)
//      lane mapping "p" = one particle per lane (9 updates each), "q" = 4 lanes per particle, lane j takes
//      column j (3 active), 3 updates each -- adjacent lanes on adjacent banks.
//  H*  hybrid: of the 32 warps of each CTA, `nred` do the L2 path (sector-aligned 3-lane RED.F64 rows into an
//      8.45 MB grid, the pattern of k_deposit_quads) while the rest do fx2 into shared memory; time-boxed, so
//      both rates are measured while the other path is running.
//  R*  routing: read 24 B/particle from HBM, write one 32 B record per particle to a scattered position of a
//      32 MB (L2-resident) queue buffer with one 256-bit store, and read the other buffer back coalesced.
//  B   grid-wide barrier (atomic counter + spin), 148 co-resident CTAs, cooperative launch.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t xs32(uint32_t &s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
__device__ __forceinline__ double ld_stream(const double *p) {
    double v; unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void red_f64(double *a, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(a), "d"(v) : "memory"); }

constexpr int kRows = 9, kPitch = 1040, kCells = kRows * kPitch;

__device__ __forceinline__ void fx2_add(unsigned *lo, unsigned *hi, int c, double v) {
    const long long q = __double2ll_rn(v * 70368744177664.0);          // 2^46 quanta per unit
    const unsigned l = (unsigned)q, h = (unsigned)((unsigned long long)q >> 32);
    const unsigned old = atomicAdd(lo + c, l);
    const unsigned carry = (old + l) < old ? 1u : 0u;
    atomicAdd(hi + c, h + carry);
}

template <int MODE>
__device__ __forceinline__ void upd(double *t, unsigned *lo, unsigned *hi, int c, double v) {
    if (MODE == 0) atomicAdd(t + c, v);
    else if (MODE == 1) { volatile double *p = t + c; *p = *p + v; }
    else if (MODE == 2) fx2_add(lo, hi, c, v);
    else atomicAdd(lo + c, (unsigned)__double2ll_rn(v * 1048576.0));
}

// MODE: 0 cas, 1 plain, 2 fx2, 3 fx1.  QUAD: lane mapping.
template <int MODE, bool QUAD>
__global__ void __launch_bounds__(1024, 1) k_fx(double *out, int iters) {
    extern __shared__ double tile[];
    unsigned *lo = reinterpret_cast<unsigned *>(tile), *hi = lo + kCells;
    for (int i = threadIdx.x; i < kCells; i += blockDim.x) tile[i] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t s = ((blockIdx.x * blockDim.x + threadIdx.x) / (QUAD ? 4 : 1)) * 2654435761u + 777u;
    xs32(s);
    for (int i = 0; i < iters; ++i) {
        const uint32_t r = xs32(s);
        const int row = __umulhi(r, kRows - 2), col = __umulhi(xs32(s), 1026);
        const double v = (double)(r & 0xffff) * 1.52587890625e-5;
        if (QUAD) {
            const int j = lane & 3;
            if (j < 3) {
#pragma unroll
                for (int a = 0; a < 3; ++a) upd<MODE>(tile, lo, hi, (row + a) * kPitch + col + j, v);
            }
        } else {
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) upd<MODE>(tile, lo, hi, (row + a) * kPitch + col + b, v);
        }
    }
    __syncthreads();
    double acc = 0;
    for (int i = threadIdx.x; i < kCells; i += blockDim.x) acc += tile[i];
    if (acc == -1.2345) out[0] = acc;
}

// hybrid: warps < nred -> L2 REDs (quad rows, sector aligned), others -> fx2 into shared memory (one particle per lane)
__global__ void __launch_bounds__(1024, 1) k_hybrid(double *grid, uint32_t cells, int nred, long long cycles,
                                                    unsigned long long *counts) {
    extern __shared__ double tile[];
    unsigned *lo = reinterpret_cast<unsigned *>(tile), *hi = lo + kCells;
    for (int i = threadIdx.x; i < kCells; i += blockDim.x) tile[i] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 4242u;
    xs32(s);
    const long long t0 = clock64();
    unsigned long long done = 0;
    if (warp < nred) {
        uint32_t sg = ((blockIdx.x * blockDim.x + threadIdx.x) / 4) * 2654435761u + 99u;
        xs32(sg);
        const uint32_t span = (cells - 4 * 1028 - 64) / 4;
        const int j = lane & 3;
        while (clock64() - t0 < cycles) {
#pragma unroll 1
            for (int k = 0; k < 8; ++k) {
                const uint32_t c = __umulhi(xs32(sg), span) * 4;
                if (j < 3) { red_f64(grid + c + j, 1.0); red_f64(grid + c + 1028 + j, 1.0); red_f64(grid + c + 2056 + j, 1.0); }
            }
            done += 8;           // particles per 4-lane group
        }
        if (j == 0) atomicAdd(counts + 0, done);
    } else {
        while (clock64() - t0 < cycles) {
#pragma unroll 1
            for (int k = 0; k < 8; ++k) {
                const uint32_t r = xs32(s);
                const int row = __umulhi(r, kRows - 2), col = __umulhi(xs32(s), 1026);
                const double v = (double)(r & 0xffff) * 1.52587890625e-5;
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) fx2_add(lo, hi, (row + a) * kPitch + col + b, v);
            }
            done += 8;
        }
        atomicAdd(counts + 1, done);
    }
    __syncthreads();
    double acc = 0;
    for (int i = threadIdx.x; i < kCells; i += blockDim.x) acc += tile[i];
    if (acc == -1.2345) grid[0] = acc;
}

// routing traffic: MODE 0 = stream only (HBM read of 24 B/particle), 1 = + scattered 32 B record stores into the
// L2-resident queue, 2 = + coalesced read-back of the other buffer
template <int MODE>
__global__ void __launch_bounds__(512, 2) k_route(const double *x, const double *y, const double *w, int64_t n,
                                                  double *q0, double *q1, uint32_t batch_mask, double *sink) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    double acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        double a, b, c;
        a = ld_stream(x + i);
        b = ld_stream(y + i);
        c = ld_stream(w + i);
        if (MODE == 0) { acc += a + b + c; continue; }
        const uint32_t batch = (uint32_t)(i / ((int64_t)batch_mask + 1));
        const uint32_t pos = ((uint32_t)i * 2654435761u) & batch_mask;          // a bijection of the batch
        double *dst = ((batch & 1) ? q1 : q0) + 4 * (size_t)pos;
        asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "l"(__double_as_longlong(a)),
                     "l"(__double_as_longlong(b)), "l"(__double_as_longlong(c)), "l"((long long)i) : "memory");
        if (MODE == 2) {
            const double *src = ((batch & 1) ? q0 : q1) + 4 * (size_t)((uint32_t)i & batch_mask);
            unsigned long long r0, r1, r2, r3;
            asm volatile("ld.global.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(r0), "=l"(r1), "=l"(r2), "=l"(r3) : "l"(src));
            acc += __longlong_as_double(r0) + __longlong_as_double(r1) + __longlong_as_double(r2) + (double)r3;
        }
    }
    if (acc == -1.2345) sink[0] = acc;
}

__global__ void __launch_bounds__(1024, 1) k_gridbar(unsigned *counter, int rounds) {
    for (int k = 0; k < rounds; ++k) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(counter, 1u);
            const unsigned target = (unsigned)(k + 1) * gridDim.x;
            unsigned v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter)); } while (v < target);
        }
        __syncthreads();
    }
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

template <int MODE, bool QUAD> void run_fx(const char *tag, double *out, int sms) {
    const int iters = 512;
    CK(cudaFuncSetAttribute(k_fx<MODE, QUAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCells * 8));
    float ms = timeit([&] { k_fx<MODE, QUAD><<<sms, 1024, kCells * 8>>>(out, iters); });
    const double parts = (double)sms * 1024 / (QUAD ? 4 : 1) * iters;
    printf("%-40s: %7.3f ms  %8.1f Gupd/s  %7.1f Gparticles/s (TSC, 9 updates)\n", tag, ms, parts * 9 / ms * 1e-6, parts / ms * 1e-6);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("device %s  SMs %d  clock %d kHz\n", p.name, sms, p.clockRate);
    const uint32_t cells2d = 1028u * 1028u;
    double *grid; CK(cudaMalloc(&grid, (size_t)cells2d * 8 * 4)); CK(cudaMemset(grid, 0, (size_t)cells2d * 8 * 4));

    printf("--- F: shared-memory TSC accumulation, 9 x 1040 band, 1 CTA x 1024 threads per SM ---\n");
    run_fx<0, false>("F cas/p   atomicAdd(double), lane=particle", grid, sms);
    run_fx<0, true>("F cas/q   atomicAdd(double), 4 lanes/particle", grid, sms);
    run_fx<1, false>("F plain/p load+add+store (racy)", grid, sms);
    run_fx<1, true>("F plain/q load+add+store (racy)", grid, sms);
    run_fx<2, false>("F fx2/p   2 x ATOMS.ADD.U32 with carry", grid, sms);
    run_fx<2, true>("F fx2/q   2 x ATOMS.ADD.U32 with carry", grid, sms);
    run_fx<3, false>("F fx1/p   1 x ATOMS.ADD.U32", grid, sms);
    run_fx<3, true>("F fx1/q   1 x ATOMS.ADD.U32", grid, sms);

    printf("--- H: hybrid, nred of 32 warps on L2 REDs, rest on fx2 shared atomics (time-boxed 1.5e6 cycles) ---\n");
    {
        unsigned long long *counts; CK(cudaMalloc(&counts, 16));
        CK(cudaFuncSetAttribute(k_hybrid, cudaFuncAttributeMaxDynamicSharedMemorySize, kCells * 8));
        for (int nred : {0, 4, 8, 12, 16, 24, 32}) {
            CK(cudaMemset(counts, 0, 16));
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            CK(cudaEventRecord(e0));
            k_hybrid<<<sms, 1024, kCells * 8>>>(grid, cells2d, nred, 1500000ll, counts);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            unsigned long long h[2]; CK(cudaMemcpy(h, counts, 16, cudaMemcpyDeviceToHost));
            printf("H nred=%2d : %7.3f ms  L2 path %7.1f Gp/s  + shared fx2 path %7.1f Gp/s  = %7.1f Gp/s\n", nred, ms,
                   h[0] / ms * 1e-6, h[1] / ms * 1e-6, (h[0] + h[1]) / ms * 1e-6);
        }
    }

    printf("--- R: routing through an L2-resident queue (2 x 32 MB), 2^27 particles ---\n");
    {
        const int64_t n = int64_t(1) << 27;
        double *x, *y, *w, *q0, *q1;
        CK(cudaMalloc(&x, n * 8)); CK(cudaMalloc(&y, n * 8)); CK(cudaMalloc(&w, n * 8));
        CK(cudaMemset(x, 0, n * 8)); CK(cudaMemset(y, 0, n * 8)); CK(cudaMemset(w, 0, n * 8));
        const uint32_t batch = 1u << 20;
        CK(cudaMalloc(&q0, (size_t)batch * 32)); CK(cudaMalloc(&q1, (size_t)batch * 32));
        CK(cudaMemset(q0, 0, (size_t)batch * 32)); CK(cudaMemset(q1, 0, (size_t)batch * 32));
        float ms = timeit([&] { k_route<0><<<sms * 2, 512>>>(x, y, w, n, q0, q1, batch - 1, grid); });
        printf("R0 stream 24 B/particle              : %7.3f ms  %7.1f Gp/s  %7.1f GB/s\n", ms, n / ms * 1e-6, n * 24.0 / ms * 1e-6);
        ms = timeit([&] { k_route<1><<<sms * 2, 512>>>(x, y, w, n, q0, q1, batch - 1, grid); });
        printf("R1 + scattered 32 B record stores    : %7.3f ms  %7.1f Gp/s\n", ms, n / ms * 1e-6);
        ms = timeit([&] { k_route<2><<<sms * 2, 512>>>(x, y, w, n, q0, q1, batch - 1, grid); });
        printf("R2 + coalesced read-back of records  : %7.3f ms  %7.1f Gp/s\n", ms, n / ms * 1e-6);
        cudaFree(x); cudaFree(y); cudaFree(w); cudaFree(q0); cudaFree(q1);
    }

    printf("--- B: grid barrier, %d CTAs x 1024 threads, cooperative launch ---\n", sms);
    {
        unsigned *counter; CK(cudaMalloc(&counter, 4));
        int rounds = 1000;
        void *args[] = {&counter, &rounds};
        float best = 1e30f;
        for (int r = 0; r < 4; ++r) {
            CK(cudaMemset(counter, 0, 4));
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            CK(cudaEventRecord(e0));
            CK(cudaLaunchCooperativeKernel((void *)k_gridbar, dim3(sms), dim3(1024), args, 0, 0));
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        printf("B  %d barriers: %7.3f ms  -> %6.2f us per barrier\n", rounds, best, best * 1e3 / rounds);
    }
    return 0;
}
