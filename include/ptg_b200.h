/*
 * ptg_b200.h -- C ABI of the B200-native particle-to-grid deposition library
 *               (libptg_b200.so, sm_100a only).
 *
 * This is the drop-in boundary for postpic's one native component,
 *     postpic/particles/_particlestogrid.pyx
 *         histogram    :38-130     (1-D deposit)
 *         histogram2d  :135-278    (2-D deposit)
 *         histogram3d  :284-465    (3-D deposit)
 * and for the two full-array passes its only caller makes before it,
 *     postpic/particles/_routines.py:115-127   (np.min / np.max when a range end is None)
 * and the one pass the caller's caller makes after it,
 *     postpic/particles/particles.py:1101-1102  (h /= dV).
 *
 * Conventions
 *   - plain C types, no CUDA/torch types in any signature; `stream` is a cudaStream_t
 *     passed as void* (NULL = the legacy default stream);
 *   - every function returns PTG_OK (0) or a negative PTG_E* code, never throws, never
 *     prints; ptg_last_error() returns a thread-local human readable message;
 *   - "device" pointers are CUDA device pointers on the current device, "host"
 *     pointers are ordinary process memory (pageable or pinned);
 *   - element strides: strides[k] is the distance, in doubles, between consecutive
 *     particles of array k (k = 0,1,2 for x,y,z and 3 for the weights); NULL means
 *     all arrays are contiguous (stride 1).  The reference accepts arbitrarily strided
 *     1-D buffers (np.ndarray[np.double_t, ndim=1] without mode='c', pyx:38,135,284);
 *   - the grid is the reference's ghost-padded accumulator (pyx:81,96,114,193,219,249,
 *     354,387,426): C order, dims[a] = bins[a] + 2*shape, float64.  The histogram the
 *     reference returns is the interior view grid[shape:shape+bins[a], ...].
 *   - lo[a] and inv[a] are computed by the caller exactly as the reference does
 *     (pyx:53-60: degenerate-range widening, inv = 1.0/(hi-lo)*bins) -- or by
 *     ptg_axis_setup() below, which restates that arithmetic in C.
 *
 * There is NO CPU fallback: on a machine without an sm_100-class GPU every compute
 * entry point returns PTG_ENODEVICE.
 */
#ifndef PTG_B200_H
#define PTG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTG_ABI_VERSION 1

/* ---- return codes ------------------------------------------------------------- */
#define PTG_OK            0
#define PTG_EINVAL       -1   /* bad argument (ndim, shape, bins, NULL pointer, n < 0 ...)        */
#define PTG_ECUDA        -2   /* a CUDA runtime call or kernel failed; see ptg_last_error()       */
#define PTG_ENOMEM       -3   /* workspace too small / allocation failed                          */
#define PTG_ENODEVICE    -4   /* no usable sm_100 GPU                                             */
#define PTG_EUNSUPPORTED -5   /* valid request this build cannot serve (e.g. grid > 2^31 cells)   */

/* ---- flags for the deposit entry points ------------------------------------------ */
#define PTG_ZERO_GRID        0x1u   /* clear the padded grid before depositing                     */
/* deposition method (bits 4..7).  AUTO launches a one-CTA probe kernel that looks at 16 blocks of
 * consecutive particles and lets exactly one of RUNS / TILED / ROWS do the work (decided on
 * the device, no host round trip)                                                             */
#define PTG_METHOD_MASK      0xF0u
#define PTG_METHOD_AUTO      0x00u
#define PTG_METHOD_GLOBAL    0x10u  /* one RED.E.ADD.F64 per stencil cell, straight to L2          */
#define PTG_METHOD_ROWS      0x20u  /* lanes regrouped so one RED instruction covers whole stencil
                                       rows (sector-coalesced L2 reductions)                      */
#define PTG_METHOD_TILED     0x30u  /* shared-memory privatised sub-grid tile per CTA, warp-
                                       aggregated updates, tile merge to HBM                      */
#define PTG_METHOD_SORTED    0x40u  /* deterministic: sort by cell + segmented reduce, no atomics  */
#define PTG_METHOD_RUNS      0x50u  /* cell-ordered input: run-length accumulation in registers,
                                       one reduction per run instead of one per particle          */
#define PTG_METHOD_RUNS_WIDE 0x60u  /* RUNS with a register window one cell wider per axis: keeps
                                       runs alive when neighbours' base cells differ by one (odd
                                       shapes on cell-ordered input); (S+2)^ndim <= 27 only        */

#define PTG_METHOD_FIXED     0x70u  /* deterministic AND order-independent: every contribution is rounded once to a
                                       multiple of 2^-45 of the power of two above the exact largest |weight| and all
                                       sums are integer sums (64-bit in shared memory, 128-bit per cell in global
                                       memory); bit-identical across runs, launch geometries and input permutations,
                                       at the speed of the atomic paths                                           */

/* Shared-memory tiles (method TILED, also when AUTO picks it) accumulate in 64-bit fixed point on native
 * 32-bit shared atomics: each contribution is rounded to a multiple of 2^-45 of the power of two above the
 * largest sampled |weight| (<= 2^-46 relative rounding per contribution; integer-valued weights below 2^45
 * stay exact; larger / non-finite weights take the exact path), ~3x the rate of the fp64 CAS loop.  This flag
 * keeps the tile in fp64 (every contribution added in fp64, as all other methods do).                     */
#define PTG_EXACT_ACCUM      0x100u

/* ---- introspection ------------------------------------------------------------- */
int         ptg_abi_version(void);
const char *ptg_last_error(void);
const char *ptg_strerror(int code);
/* fills name (<= name_len bytes), SM count, L2 bytes, total HBM bytes of `device`. */
int ptg_device_info(int device, char *name, int name_len, int *sm_count,
                    int64_t *l2_bytes, int64_t *hbm_bytes, int *cc_major, int *cc_minor);

/* ---- host-side axis set-up (pyx:53-60, :155-169, :307-328) ----------------------- */
/* in: range ends and bin count.  out: widened lo/hi, inv = 1.0/(hi-lo)*bins. */
int ptg_axis_setup(double lo_in, double hi_in, int bins, double *lo, double *hi, double *inv);

/* ---- device-resident deposit (replaces the loops pyx:70-129, :181-276, :341-463) -- */
/*
 * x,y,z,w: device pointers (y,z ignored beyond ndim; w NULL = the reference's
 * `weights is None` branch).  grid: device pointer to the padded accumulator.
 * The call is asynchronous on `stream`.  Scratch memory (method SORTED's sort buffers, the
 * staggered staging grid of ROWS) is taken from `workspace` when ptg_workspace_bytes() bytes
 * are supplied for SORTED, and otherwise from the device's stream-ordered memory pool
 * (cudaMallocAsync on `stream`, no synchronisation); workspace may always be NULL.
 */
int ptg_deposit_f64(int ndim, int shape, int64_t n,
                    const double *x, const double *y, const double *z, const double *w,
                    const int64_t *strides,
                    const double *lo, const double *inv, const int *bins,
                    double *grid, unsigned flags,
                    void *workspace, size_t workspace_bytes, void *stream);

size_t ptg_workspace_bytes(int ndim, int shape, int64_t n, const int *bins, unsigned flags);

/* number of padded cells = prod(bins[a] + 2*shape); 0 on bad arguments */
int64_t ptg_padded_cells(int ndim, int shape, const int *bins);

/* ---- helpers on device buffers ------------------------------------------------------ */
/* out2[0] = min, out2[1] = max over x[0..n) (out2: device pointer to 2 doubles).  A NaN
 * anywhere yields NaN for both, like np.min/np.max; replaces _routines.py:117-126.     */
int ptg_minmax_f64(int64_t n, const double *x, int64_t stride, double *out2, void *stream);
/* grid[i] *= factor for i in [0,count): the h /= dV of particles.py:1101-1102 with
 * factor = 1/dV computed by the caller.                                               */
int ptg_scale_f64(double *grid, int64_t count, double factor, void *stream);
/* grid[i] /= divisor with an IEEE division: bit-identical to the reference's `h /= dV`. */
int ptg_divide_f64(double *grid, int64_t count, double divisor, void *stream);
int ptg_zero_f64(double *grid, int64_t count, void *stream);

/* ---- per-particle weight expressions in one pass (the step before the path) --------------- */
/* out[i] = weight[i] * f(px[i], py[i], pz[i]) (weight NULL = 1) for the kinetic scalars of postpic's
 * default scalar context (scalarproperties.py:183-205), which MultiSpecies._createHistgram
 * evaluates as `weight * (expr)` through numexpr (particles.py:1097).  Device pointers; strides[4]
 * (px, py, pz, weight) in doubles or NULL.  mass in kg; c and elementary_charge are passed in so
 * that the caller's constants (scipy.constants in the reference) are used.                       */
#define PTG_W_GAMMA_M1 1
#define PTG_W_GAMMA    2
#define PTG_W_EKIN     3   /* J   */
#define PTG_W_EKIN_MEV 4
#define PTG_W_EKIN_KEV 5
#define PTG_W_P        6   /* |p| */
/* The whole default scalar context of postpic (scalarproperties.py:159-222) as enumerated kinds; 1..6 are the
 * PTG_W_* above.  Usable as the fused weight factor (fuse_kind) and as fused AXIS expressions (axis_kind[]) of
 * ptg_deposit_ex_f64, and with ptg_scalar_f64.                                                             */
#define PTG_S_X 7
#define PTG_S_Y 8
#define PTG_S_Z 9
#define PTG_S_PX 10
#define PTG_S_PY 11
#define PTG_S_PZ 12
#define PTG_S_WEIGHT 13
#define PTG_S_MASS 14
#define PTG_S_CHARGE 15
#define PTG_S_TIME 16
#define PTG_S_MASS_U 17
#define PTG_S_CHARGE_E 18
#define PTG_S_X_UM 19
#define PTG_S_Y_UM 20
#define PTG_S_Z_UM 21
#define PTG_S_NP2 22
#define PTG_S_GAMMA_M 23
#define PTG_S_BETA 24
#define PTG_S_V 25
#define PTG_S_VX 26
#define PTG_S_VY 27
#define PTG_S_VZ 28
#define PTG_S_BETAX 29
#define PTG_S_BETAY 30
#define PTG_S_BETAZ 31
#define PTG_S_ERUHE 32
#define PTG_S_EKIN_MEV_AMU 33
#define PTG_S_EKIN_MEV_QM 34
#define PTG_S_EKIN_KEV_AMU 35
#define PTG_S_EKIN_KEV_QM 36
#define PTG_S_ANGLE_XAXIS 37
#define PTG_S_ANGLE_YAXIS 38
#define PTG_S_ANGLE_ZAXIS 39
#define PTG_S_ANGLE_XY 40
#define PTG_S_ANGLE_YX 41
#define PTG_S_ANGLE_YZ 42
#define PTG_S_ANGLE_ZY 43
#define PTG_S_ANGLE_ZX 44
#define PTG_S_ANGLE_XZ 45
#define PTG_S_R_XY 46
#define PTG_S_R_YZ 47
#define PTG_S_R_ZX 48
#define PTG_S_R_XYZ 49
#define PTG_S_LAST 49
int ptg_weights_f64(int kind, int64_t n, const double *px, const double *py, const double *pz,
                    const double *weight, const int64_t *strides, double mass, double c, double elementary_charge,
                    double *out, void *stream);

/* The same expression evaluated INSIDE the deposit kernels (no weights array is written or re-read):
 * ptg_deposit_f64 with the particle weight replaced by w[i] * f(px[i], py[i], pz[i]) (w NULL = 1).
 * This is `createField(..., weights='Ekin_MeV')` of particles.py:1097-1100 in one pass over
 * x, y(, z), px, py, pz, weight.  pstrides[3] in doubles or NULL; all pointers device pointers.      */
int ptg_deposit_fused_f64(int ndim, int shape, int64_t n,
                          const double *x, const double *y, const double *z, const double *w,
                          const int64_t *strides,
                          const double *lo, const double *inv, const int *bins,
                          double *grid, unsigned flags,
                          int kind, const double *px, const double *py, const double *pz,
                          const int64_t *pstrides, double mass, double c, double elementary_charge,
                          void *stream);

/* ---- the general device entry: everything above plus a selection mask, in one argument block ----- */
/* mask (device pointer, one byte per particle, stride in bytes; NULL = all): only particles with a
 * non-zero mask byte are deposited -- MultiSpecies.filter / compress (particles.py:219-232, :561-582) as
 * predicated deposition, no compressed copies of the particle arrays.  Set struct_size =
 * sizeof(ptg_deposit_args); unused members zero / NULL.                                           */
typedef struct ptg_deposit_args {
    size_t struct_size;
    int ndim, shape;
    int64_t n;
    const double *x, *y, *z, *w;
    const int64_t *strides;             /* [4] or NULL */
    const double *lo, *inv;
    const int *bins;
    double *grid;
    unsigned flags;
    const unsigned char *mask;
    int64_t mask_stride;
    int fuse_kind;                      /* 0 = no fused weight expression, else PTG_W_* */
    const double *px, *py, *pz;
    const int64_t *pstrides;            /* [3] or NULL */
    double mass, c, elementary_charge;
    void *workspace;
    size_t workspace_bytes;
    void *stream;
    /* ---- fused axis expressions (struct_size tells whether the caller knows these members) ----
     * axis_kind[a] != 0: coordinate a of every particle is the scalar PTG_S_<kind> evaluated in the kernel from
     * the atomic arrays pos_x/pos_y/pos_z (positions), px/py/pz (momenta), w (macro-particle weight) and the
     * species constants; the slot pointer x / y / z of that axis is ignored (may be NULL).  The axis and
     * `weight * (expr)` evaluations of particles.py:1071,1097 without a single N-sized temporary.          */
    int axis_kind[3];
    const double *pos_x, *pos_y, *pos_z;
    const int64_t *pos_strides;         /* [3] or NULL */
    double charge, time, atomic_mass;   /* species charge [C], dump time [s], atomic mass unit [kg] */
    double weight_const;                /* macro-particle weight used when w is NULL (0 is read as 1) */
} ptg_deposit_args;
int ptg_deposit_ex_f64(const ptg_deposit_args *args);
/* out[i] = scalar `kind` (1..PTG_S_LAST) of particle i, evaluated from the atomic arrays of `args` exactly as the
 * deposit kernels evaluate it (device pointers; only n, the atomic arrays, the constants and stream are read). */
int ptg_scalar_f64(const ptg_deposit_args *args, int kind, double *out);

/* ---- host-buffer drop-in: H2D + deposit + D2H in one call --------------------------- */
/*
 * Same meaning as ptg_deposit_f64 but x,y,z,w and grid are HOST pointers.  Particles
 * are streamed to the GPU in chunks through a small ring of device buffers so that the
 * PCIe copy of chunk k+1 overlaps the deposit of chunk k; the padded grid is copied
 * back to `grid` (which is overwritten, not accumulated into) before the call returns.
 * Page-locked unit-stride sources are DMA'd in place; pageable memory and any element
 * stride (negative and zero included) are gathered into a pinned ring by worker threads
 * first (PTG_HOST_THREADS overrides their number).
 * Synchronous.  `device` < 0 means the current device.
 */
int ptg_histogram_host_f64(int ndim, int shape, int64_t n,
                           const double *x, const double *y, const double *z, const double *w,
                           const int64_t *strides,
                           const double *lo, const double *inv, const int *bins,
                           double *grid, unsigned flags, int device);
/* The same pipeline with the result left on the device: particles from HOST arrays are streamed through the
 * staging ring and deposited into the caller's DEVICE grid `grid_device` (on `device`), which is cleared first
 * when flags has PTG_ZERO_GRID and accumulated into otherwise.  Synchronous: the grid is complete on return.
 * This is what a multi-GPU caller wants (postpic_b200.parallel.histogram_host_sharded): every rank streams its
 * own particle block with H2D overlapping the deposit, the grids are then summed over NVLink, and only the
 * sum crosses PCIe back.                                                                              */
int ptg_deposit_host_f64(int ndim, int shape, int64_t n,
                         const double *x, const double *y, const double *z, const double *w,
                         const int64_t *strides,
                         const double *lo, const double *inv, const int *bins,
                         double *grid_device, unsigned flags, int device);
int ptg_minmax_host_f64(int64_t n, const double *x, int64_t stride, double *out2, int device);
/* dst_device[i] = src[i * stride] for i in [0, n): host array (pageable or pinned, any element stride)
 * to a contiguous device array through the same threaded pinned staging ring.  Synchronous.  Used when
 * several passes need the particles (range completion followed by the deposit): one trip over PCIe. */
int ptg_upload_host_f64(int64_t n, const double *src, int64_t stride, double *dst_device, int device);
/* Release what the library caches per device: the staging ring of the host entry points (3 device slots of
 * 128 MiB and, once a pageable / strided source has been seen, 3 pinned host slots of 64 MiB, plus a device grid of
 * the largest padded grid deposited so far), their streams, the side streams / events of method AUTO, the ring of
 * probe verdicts, and the blocks cached in the library's private stream-ordered memory pool.  All of it is allocated
 * on first use and kept until this call (or process exit) so that repeated calls pay no allocation.            */
int ptg_release(void);

/* ---- test hooks (used by tests/ only) ------------------------------------------------- */
/* out[i] = the library's division-by-6 of in[i] (device pointers) -- checked bit-for-bit
 * against IEEE division by the tests.                                                   */
int ptg_selftest_div6(int64_t n, const double *in, double *out, void *stream);
/* the library's radix sort (method SORTED): keys[0..n) (device, uint32) sorted in place by their low
 * `bits` bits, stably; perm[i] = original index of the i-th sorted key.                           */
int ptg_selftest_radix_sort(int64_t n, int bits, unsigned *keys, unsigned *perm, void *stream);
/* the verdict of the most recent AUTO / TILED probe on the current device (synchronises):
 * mode 0 = ROWS, 1 = TILED, 2 = RUNS, 3 = RUNS_WIDE; ext[3] = tile window extents.          */
int ptg_debug_last_probe(int *mode, int *ext, float *est_hit, float *run_frac, float *near_frac);

#ifdef __cplusplus
}
#endif
#endif /* PTG_B200_H */
