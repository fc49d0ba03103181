// ptg_runs_common.cuh -- pieces shared by the run-length kernels (ptg_deposit_runs.cu: vector-load and
// bulk-copy kernels; ptg_deposit_runs_bulkwide.cu: bulk-copy kernel with widened register windows).
#pragma once
#include "ptg_internal.cuh"

namespace ptg {

#ifndef PTG_RUNS_MERGE_MAX_HEADS
#define PTG_RUNS_MERGE_MAX_HEADS 20      // warp merge only when the 32 lanes hold at most this many runs
#endif

// One run's window -> the grid.  key = padded linear index of the window's first cell.  In a widened window
// exact zeros are skipped: they carry nothing, and the outermost cells may lie beyond the grid (they can
// only ever hold zero there).
template <int MA, int MC, int MD, bool SKIP_ZERO, int MV>
__device__ __forceinline__ void runs_commit(double *grid, int64_t g1, int64_t g2, int key, const double (&acc)[MV]) {
    double *g = grid + key;
#pragma unroll
    for (int a = 0; a < MA; ++a)
#pragma unroll
        for (int c = 0; c < MC; ++c)
#pragma unroll
            for (int d = 0; d < MD; ++d) {
                const double v = acc[(a * MC + c) * MD + d];
                if (!SKIP_ZERO || v != 0.0) red_add(g + (a * g1 + c) * g2 + d, v);
            }
}

// Merge the runs still open at the end of each lane's particles across the warp (segmented shuffle
// reduction over lanes holding the same window); returns whether this lane has to commit its acc.
template <int MV>
__device__ __forceinline__ bool runs_warp_merge(int cur, double (&acc)[MV], int lane) {
    const int key = cur >= 0 ? cur : -1 - lane;
    const int prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = lane == 0 || prev != key;
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    bool mine = cur >= 0;
    if (__popc(heads) <= PTG_RUNS_MERGE_MAX_HEADS) {
        const unsigned above = lane == 31 ? 0u : heads & ~((2u << lane) - 1u);
        const int next_head = above ? __ffs(above) - 1 : 32;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const bool take = lane + dd < next_head;
#pragma unroll
            for (int k = 0; k < MV; ++k) {
                const double o = __shfl_down_sync(0xffffffffu, acc[k], dd);
                // predicated add (one @P DADD) instead of the add + two 32-bit selects nvcc emits
                asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p add.rn.f64 %0, %0, %1;\n\t}"
                    : "+d"(acc[k]) : "d"(o), "r"((int)take));
            }
        }
        mine = mine && head;
    }
    return mine;
}

#ifndef PTG_BULK_PER
#define PTG_BULK_PER 8
#endif
#ifndef PTG_BULK_WARPS
#define PTG_BULK_WARPS 16
#endif
#ifndef PTG_BULK_STAGES
#define PTG_BULK_STAGES 2
#endif
#ifndef PTG_BULK_INTERLEAVE
#define PTG_BULK_INTERLEAVE 0
#endif
constexpr int kBulkPer = PTG_BULK_PER;           // consecutive particles per lane and tile
constexpr int kBulkTile = 32 * kBulkPer;         // particles per warp tile
template <int ND, bool W> struct BulkCfg {
    static constexpr int NA = ND + (W ? 1 : 0);                         // arrays streamed
    static constexpr int STAGES = NA <= 2 ? PTG_BULK_STAGES + 1 : PTG_BULK_STAGES;
    static constexpr int TILE_BYTES = NA * kBulkTile * 8;               // one stage of one warp
    static constexpr int FIT = (200 * 1024) / (STAGES * TILE_BYTES) / 4 * 4;
    static constexpr int WARPS = FIT >= PTG_BULK_WARPS ? PTG_BULK_WARPS : FIT;
    static constexpr int THREADS = WARPS * 32;
    static constexpr size_t SMEM = (size_t)WARPS * STAGES * (TILE_BYTES + 16) + WARPS * 128;   // + full / empty barriers, + one word per lane
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned); the particle
// stream is touched once: evict-first in L2
__device__ __forceinline__ void bulk_load(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    const uint32_t a = smem_u32(bar);
    unsigned done = 0;
    for (unsigned spins = 0; !done; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (spins > (1u << 28)) __trap();            // a tile that never lands is a bug, not a reason to hang the GPU
    }
}

}  // namespace ptg
