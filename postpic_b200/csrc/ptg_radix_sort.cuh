// ptg_radix_sort.cuh -- stable LSD radix sort of (uint32 key, uint32 value) pairs, hand-written for
// method SORTED (keys = stencil base cells, values = particle indices).
//
// 8-bit digits, ceil(bits / 8) passes, three kernels per pass:
//   k_rs_count    every CTA counts the digits of its tile of kRsTile consecutive pairs, per warp
//                 (a warp owns a contiguous kRsTile/8 slice and walks it 32 pairs at a time, so all
//                 loads are coalesced) and writes its 256 totals to counts[digit][cta]
//   k_rs_scan     one CTA per digit turns its row counts[digit][*] into an exclusive prefix and
//                 records the row total; CTA 0 then waits for nothing -- the 256 totals are scanned
//                 by k_rs_bases (one small CTA)
//   k_rs_scatter  every CTA recounts its tile per warp, turns the counts into tile-local positions and
//                 walks its slices again, placing every pair at its digit-ordered position inside
//                 the tile (64 KB of shared memory); the rank of a pair inside its round of 32 comes
//                 from __match_any_sync (lanes with the same digit, lower lanes first), so equal keys
//                 keep their input order -- the sort is STABLE, which is what makes method SORTED
//                 sum every cell's particles in original particle order.  The tile is then copied
//                 out digit run by digit run (coalesced stores).
// Traffic per pass: the pairs are read twice and written once (24 B per pair).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ptg {

constexpr int kRsThreads = 256;                    // 8 warps
constexpr int kRsRounds = 32;                      // rounds of 32 pairs per warp
constexpr int kRsTile = kRsThreads * kRsRounds;    // 8192 pairs per CTA

// digit of one lane, or a value no other lane has when the lane holds no pair
__device__ __forceinline__ unsigned rs_digit(uint32_t key, int shift, bool valid, int lane) {
    return valid ? ((key >> shift) & 255u) : 256u + (unsigned)lane;
}

// per-warp digit histogram of this warp's slice of the CTA's tile (shared: hist[8][256]).  Counting needs
// no order, so it uses the native 32-bit shared-memory atomic, four rounds of loads in flight at a time.
__device__ __forceinline__ void rs_count_slice(const uint32_t *keys, int64_t n, int64_t slice0, int shift,
                                               unsigned *hist /* this warp's 256 counters */) {
    const int lane = threadIdx.x & 31;
    for (int r = 0; r < kRsRounds; r += 4) {
        uint32_t key[4];
        bool valid[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = slice0 + (int64_t)(r + u) * 32 + lane;
            valid[u] = i < n;
            key[u] = valid[u] ? keys[i] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (valid[u]) atomicAdd(&hist[(key[u] >> shift) & 255u], 1u);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kRsThreads) k_rs_count(const uint32_t *keys, int64_t n, int shift,
                                                         unsigned *counts, int64_t nctas) {
    __shared__ unsigned hist[8][256];
    for (int k = threadIdx.x; k < 8 * 256; k += kRsThreads) (&hist[0][0])[k] = 0u;
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const int64_t tile0 = (int64_t)blockIdx.x * kRsTile;
    rs_count_slice(keys, n, tile0 + (int64_t)warp * (kRsTile / 8), shift, hist[warp]);
    __syncthreads();
    const int d = threadIdx.x;                       // 256 threads = 256 digits
    unsigned t = 0;
    for (int w = 0; w < 8; ++w) t += hist[w][d];
    counts[(int64_t)d * nctas + blockIdx.x] = t;
}

// row d of counts -> exclusive prefix along the CTAs; totals[d] = row sum
__global__ void __launch_bounds__(1024) k_rs_scan(unsigned *counts, int64_t nctas, unsigned long long *totals) {
    __shared__ unsigned long long warp_sums[32];
    __shared__ unsigned long long carry_s;
    unsigned *row = counts + (int64_t)blockIdx.x * nctas;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0ull;
    __syncthreads();
    for (int64_t base = 0; base < nctas; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const unsigned v = i < nctas ? row[i] : 0u;
        unsigned long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned long long s = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            warp_sums[lane] = s;                     // inclusive over warps
        }
        __syncthreads();
        const unsigned long long before = carry_s + (warp ? warp_sums[warp - 1] : 0ull) + (x - v);
        if (i < nctas) row[i] = (unsigned)before;    // < 2^32: n < 2^32 pairs
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry_s + warp_sums[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = carry_s;
}

// exclusive scan of the 256 digit totals
__global__ void __launch_bounds__(256) k_rs_bases(const unsigned long long *totals, unsigned long long *bases) {
    __shared__ unsigned long long s[256];
    s[threadIdx.x] = totals[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int d = 0; d < 256; ++d) { const unsigned long long t = s[d]; s[d] = run; run += t; }
    }
    __syncthreads();
    bases[threadIdx.x] = s[threadIdx.x];
}

// dynamic shared memory of k_rs_scatter: the tile's pairs in digit order
constexpr size_t kRsScatterSmem = (size_t)kRsTile * 2 * sizeof(uint32_t);

__global__ void __launch_bounds__(kRsThreads, 3) k_rs_scatter(const uint32_t *keys_in, const uint32_t *vals_in,
                                                              uint32_t *keys_out, uint32_t *vals_out, int64_t n,
                                                              int shift, const unsigned *counts, int64_t nctas,
                                                              const unsigned long long *bases) {
    extern __shared__ uint32_t sorted[];             // [kRsTile] keys, then [kRsTile] values
    uint32_t *skey = sorted, *sval = sorted + kRsTile;
    __shared__ unsigned hist[8][256];                // per-warp counts, then per-warp running local positions
    __shared__ unsigned local_start[257];            // first tile-local position of each digit
    __shared__ unsigned global_start[256];           // first output position of each digit of this CTA
    for (int k = threadIdx.x; k < 8 * 256; k += kRsThreads) (&hist[0][0])[k] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile0 = (int64_t)blockIdx.x * kRsTile;
    const int64_t slice0 = tile0 + (int64_t)warp * (kRsTile / 8);
    rs_count_slice(keys_in, n, slice0, shift, hist[warp]);
    __syncthreads();
    {   // thread d: digit total of the tile; its output base
        const int d = threadIdx.x;
        unsigned t = 0;
        for (int w = 0; w < 8; ++w) t += hist[w][d];
        local_start[d + 1] = t;
        global_start[d] = (unsigned)(bases[d] + counts[(int64_t)d * nctas + blockIdx.x]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {                          // 256-entry exclusive scan (once per tile of 8192 pairs)
        unsigned run = 0;
        local_start[0] = 0;
        for (int d = 1; d <= 256; ++d) { run += local_start[d]; local_start[d] = run; }
    }
    __syncthreads();
    {   // per-warp first local position of each digit: digit start + earlier warps of this CTA
        const int d = threadIdx.x;
        unsigned run = local_start[d];
        for (int w = 0; w < 8; ++w) { const unsigned c = hist[w][d]; hist[w][d] = run; run += c; }
    }
    __syncthreads();
    // pass 1: every pair goes to its digit-ordered place inside the tile (shared memory).  The rank of a
    // pair within its round of 32 comes from match_any, lower lanes first: stable.
    unsigned *pos = hist[warp];
    uint32_t nkey = 0u, nval = 0u;
    {
        const int64_t i = slice0 + lane;
        if (i < n) { nkey = keys_in[i]; nval = vals_in[i]; }
    }
    for (int r = 0; r < kRsRounds; ++r) {
        const int64_t i = slice0 + (int64_t)r * 32 + lane;
        const bool valid = i < n;
        const uint32_t key = nkey, val = nval;
        if (r + 1 < kRsRounds && i + 32 < n) { nkey = keys_in[i + 32]; nval = vals_in[i + 32]; }   // next round in flight
        const unsigned d = rs_digit(key, shift, valid, lane);
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned below = __popc(peers & ((1u << lane) - 1u));
        unsigned first = 0;
        if (valid) first = pos[d];                   // read by every lane of the group before the leader bumps it
        __syncwarp();
        if (valid) {
            skey[first + below] = key;
            sval[first + below] = val;
            if (below == 0) pos[d] = first + __popc(peers);
        }
        __syncwarp();
    }
    __syncthreads();
    // pass 2: copy out; consecutive tile positions of one digit are consecutive in the output, so the
    // stores of a warp fall into a few contiguous runs instead of 32 unrelated sectors
    const int64_t left = n - tile0;
    const int count = left < kRsTile ? (int)left : kRsTile;
    for (int p = threadIdx.x; p < count; p += kRsThreads) {
        const uint32_t key = skey[p];
        const unsigned d = (key >> shift) & 255u;
        const unsigned dest = global_start[d] + ((unsigned)p - local_start[d]);
        keys_out[dest] = key;
        vals_out[dest] = sval[p];
    }
}

// scratch: counts (256 * nctas u32) + totals (256 u64) + bases (256 u64)
inline size_t radix_sort_scratch_bytes(int64_t n) {
    const int64_t nctas = (n + kRsTile - 1) / kRsTile;
    return (size_t)(256 * nctas) * sizeof(unsigned) + 2 * 256 * sizeof(unsigned long long) + 256;
}

// Sorts by the low `bits` bits of the keys.  Returns in *keys_sorted / *vals_sorted which of the two
// buffers holds the result.  n < 2^32.
inline cudaError_t radix_sort_pairs(uint32_t *k0, uint32_t *k1, uint32_t *v0, uint32_t *v1, int64_t n, int bits,
                                    void *scratch, cudaStream_t st, uint32_t **keys_sorted, uint32_t **vals_sorted) {
    *keys_sorted = k0;
    *vals_sorted = v0;
    if (n <= 0) return cudaSuccess;
    const int64_t nctas = (n + kRsTile - 1) / kRsTile;
    unsigned *counts = static_cast<unsigned *>(scratch);
    char *p = static_cast<char *>(scratch) + (size_t)(256 * nctas) * sizeof(unsigned);
    p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(p) + 255) & ~uintptr_t(255));
    unsigned long long *totals = reinterpret_cast<unsigned long long *>(p);
    unsigned long long *bases = totals + 256;
    uint32_t *kin = k0, *kout = k1, *vin = v0, *vout = v1;
    cudaError_t e = cudaFuncSetAttribute(k_rs_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRsScatterSmem);
    if (e != cudaSuccess) return e;
    for (int shift = 0; shift < bits; shift += 8) {
        k_rs_count<<<(unsigned)nctas, kRsThreads, 0, st>>>(kin, n, shift, counts, nctas);
        k_rs_scan<<<256, 1024, 0, st>>>(counts, nctas, totals);
        k_rs_bases<<<1, 256, 0, st>>>(totals, bases);
        k_rs_scatter<<<(unsigned)nctas, kRsThreads, kRsScatterSmem, st>>>(kin, vin, kout, vout, n, shift, counts, nctas, bases);
        uint32_t *t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    *keys_sorted = kin;
    *vals_sorted = vin;
    return cudaGetLastError();
}

}  // namespace ptg
