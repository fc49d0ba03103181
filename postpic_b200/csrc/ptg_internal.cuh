// ptg_internal.cuh -- declarations shared by the translation units of libptg_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include "ptg_stencil.cuh"

namespace ptg {

// Decision record of the probe kernel (device memory): whether a shared-memory tile pays off
// for this launch and the window extents it chose.
enum : int { kModeRows = 0, kModeTiled = 1, kModeRuns = 2, kModeRunsWide = 3 };

constexpr int kProbeBlocks = 16;
constexpr int kTileCells = 25600;         // doubles in the 200 KB shared-memory tile of method TILED

struct TileConfig {
    int mode;          // which of the kernels launched for an AUTO call does the work
    int ext[3];        // tile window extents chosen for kModeTiled
    float est_hit;     // estimated fraction of particles that fall into such a window
    float run_frac;    // fraction of consecutive particle pairs with equal stencil base cell
    float near_frac;   // ... whose base cells differ by at most one cell forward on every axis
    unsigned ticket;   // probe CTAs that have finished (reset to 0 by the last one)
    int pad;
    // 64-bit fixed-point accumulation in shared memory (method TILED): a contribution v is added as
    // round(v * fx_inv_q) quanta on two native 32-bit shared atomics.  fx_q = 2^(e - 45) where 2^e is the
    // power of two above the largest |weight| the probe sampled; particles with |w| > fx_wcap = 2^e take
    // the exact RED.F64 path.  fx_inv_q == 0: fixed point unavailable for this launch (no sample, zero or
    // extreme weights) -- every particle takes the exact path.
    double fx_inv_q, fx_q, fx_wcap;
    // per-probe-CTA partial statistics
    struct Part { float mad[3]; int cnt; int eq_pairs; int all_pairs; int near_pairs; int pad; double wmax; } part[kProbeBlocks];
};

// ptg_deposit_global.cu -- `gate` (nullable): the kernels return at once unless gate->mode is
// kModeRows, i.e. when one of the kernels launched next to them is doing the work.
cudaError_t deposit_global(int ndim, int shape, const DepositParams &P, bool rows, const TileConfig *gate,
                           int sm_count, cudaStream_t st);

// ptg_deposit_tiled.cu -- probe + tiled kernel; `force` ignores the probe's verdict
// ptg_deposit_fixed.cu -- method FIXED: deterministic, order-independent integer accumulation
struct FixedState {
    unsigned long long wmax_bits;      // bit pattern of the largest finite |weight| (non-negative doubles order like integers)
    unsigned long long pad;
};
void fixed_set_scale(TileConfig *cfg, const FixedState *state, cudaStream_t st);   // cfg->fx_* := quantum from the exact max
cudaError_t deposit_fixed(int ndim, int shape, const DepositParams &P, TileConfig *cfg, int sm_count, cudaStream_t st);

// `exact`: accumulate the tile in fp64 (CAS loops) instead of 64-bit fixed point; `fixed` (method FIXED only): the
// probe's sampled quantum is replaced by the one derived from the exact maximum weight
// `after_probe` (nullable): an event recorded right behind the probe kernel, before the tile kernel is launched
cudaError_t deposit_tiled(int ndim, int shape, const DepositParams &P, TileConfig *cfg, bool force, bool exact,
                          int sm_count, cudaStream_t st, const FixedState *fixed = nullptr,
                          cudaEvent_t after_probe = nullptr);

// one launch for small calls on small grids (no memset, no probe); `zero_first`: the grid is overwritten, not added to
bool oneshot_applies(const DepositParams &P, int sm_count);
cudaError_t deposit_oneshot(int ndim, int shape, const DepositParams &P, TileConfig *cfg, bool zero_first, int sm_count,
                            cudaStream_t st);

// ptg_deposit_runs.cu -- run-length accumulation in registers for cell-ordered input
// (`wide`: the variant whose register window is one cell wider per axis, mode kModeRunsWide)
cudaError_t deposit_runs(int ndim, int shape, const DepositParams &P, const TileConfig *gate, bool wide, int sm_count,
                         cudaStream_t st);
bool runs_wide_available(int ndim, int shape);
// ptg_deposit_runs_bulkwide.cu -- the widened-window kernel with the particle stream on the copy engine
bool runs_bulk_wide_applies(int ndim, int shape, const DepositParams &P);
cudaError_t deposit_runs_bulk_wide(int ndim, int shape, const DepositParams &P, const TileConfig *gate, int sm_count,
                                   cudaStream_t st);

// ptg_deposit_sorted.cu -- deterministic sort-by-cell + segmented reduce; needs a workspace of
// sorted_workspace_bytes(n, cells) bytes
size_t sorted_workspace_bytes(int64_t n, int64_t cells);
cudaError_t deposit_sorted(int ndim, int shape, const DepositParams &P, void *ws, size_t ws_bytes, int sm_count,
                           cudaStream_t st);

cudaError_t selftest_radix_sort(int64_t n, int bits, uint32_t *keys, uint32_t *perm, cudaStream_t st);

// ptg_capi.cu -- stream-ordered scratch from the LIBRARY'S OWN memory pool (freed blocks stay cached there; the
// device's default pool, which other components of the process may share, is left alone)
cudaError_t scratch_alloc(void **ptr, size_t bytes, cudaStream_t st);

// ptg_util.cu
cudaError_t minmax_f64(int64_t n, const double *x, int64_t stride, double *out2, int sm_count,
                       cudaStream_t st);
cudaError_t scale_f64(double *grid, int64_t count, double factor, int sm_count, cudaStream_t st);
cudaError_t divide_f64(double *grid, int64_t count, double divisor, int sm_count, cudaStream_t st);
cudaError_t weights_f64(int kind, int64_t n, const double *px, const double *py, const double *pz,
                        const double *weight, const int64_t *strides, double mass, double c, double qe, double *out,
                        int sm_count, cudaStream_t st);
cudaError_t scalar_f64(const DepositParams &P, int kind, double *out, int sm_count, cudaStream_t st);
cudaError_t selftest_div6(int64_t n, const double *in, double *out, cudaStream_t st);

}  // namespace ptg
