/*
 * ptg_oracle.c -- CPU restatement of postpic's particle-to-grid deposition.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 * The product (postpic_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED.  oracle/_ref (the unmodified reference module, built by
 * oracle/build_ref.py) is executed against this file by tests/test_oracle.py on
 * seeded inputs, on the known-answer vectors and on the boundary probes; the
 * agreement required there is BIT-EXACT, because this restatement keeps the
 * reference's operation order (one particle at a time, in particle order, no FMA).
 *
 * What it restates (all line numbers: /root/reference/postpic/particles/_particlestogrid.pyx)
 *   - scaled coordinate u = (d - lo) * inv ............................ :71 :83 :98 :116 (1-D)
 *                                                                        :182-183 :196-197 :222-223 :252-253 (2-D)
 *                                                                        :342-344 :357-359 :390-392 :429-431 (3-D)
 *   - shape 0 (NGP): accept u>0 && u<bins on every axis, cell trunc(u)  :72-76 :184-188 :345-349
 *   - shape 1 (tophat): r=trunc(u+.5); weights (.5-u)+r, (.5+u)-r ....... :84-91 :198-214 :360-382
 *   - shape 2 (triangle): r=trunc(u); t=u-r; .5(1-t)^2, (.5+t)-t^2, .5t^2 :99-109 :224-244 :393-421
 *   - shape 3 (spline3): r=trunc(u+.5); t=(u-r)+.5; cubic B-spline ...... :117-129 :254-276 :432-463
 *   - the per-axis accept tests, which differ between axes (x on the truncated
 *     int, y on the raw double, z on the raw double only for 3-D shape 1) :85 :101 :119 :200 :228 :258 :363 :399 :438
 *   - ghost padding of `shape` cells per side, C order ................... :81 :96 :114 :193 :219 :249 :354 :387 :426
 *   - product order (wx*wy)*w, ((wx*wy)*wz)*w; the unweighted branch has no multiply by w
 *
 * The reference has undefined behaviour for some inputs (SURVEY.md section 8a "UB list"):
 * a raw-double-tested coordinate >= 2^31 converts to INT_MIN, passes `yr <= ybins`
 * and writes out of bounds.  Here (and in the CUDA path) such particles are DROPPED.
 * On an int-tested axis x86's cvttsd2si yields INT_MIN for NaN/inf/huge values, the
 * `r >= 0` test then rejects the particle; the double-domain tests below give the
 * same outcome without relying on an out-of-range conversion.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/build_oracle.py).
 */
#include <stdint.h>
#include <stddef.h>

#define PTG_MAXW 4

/* Which kind of acceptance test an axis uses for shapes >= 1. */
enum { TEST_ON_INT = 0, TEST_ON_DOUBLE = 1 };

/* One axis of the stencil.  Returns 0 when the particle is rejected by this axis.
 * On success: *base is the first padded cell index, wt[0..shape] the stencil weights. */
static int axis_stencil(int shape, int test_kind, double d, double lo, double inv, int bins,
                        int *base, double wt[PTG_MAXW])
{
    const double u = (d - lo) * inv;
    /* v is the value that gets truncated: u for shape 2, u + 0.5 for shapes 1 and 3. */
    const double v = (shape == 2) ? u : u + 0.5;
    /* trunc(v) <= bins  <=>  v < bins + 1   (also rejects NaN, +inf and anything >= 2^31) */
    if (!(v < (double)bins + 1.0)) return 0;
    if (test_kind == TEST_ON_INT) {
        /* trunc(v) >= 0  <=>  v > -1   (truncation toward zero lets (-1,0) through as r = 0) */
        if (!(v > -1.0)) return 0;
    } else {
        if (!(u >= 0.0)) return 0;
    }
    const int r = (int)v;
    const double rd = (double)r;
    if (shape == 1) {
        wt[0] = (0.5 - u) + rd;
        wt[1] = (0.5 + u) - rd;
        *base = r;                 /* r + shape - 1 */
    } else if (shape == 2) {
        const double t = u - rd;
        const double omt = 1.0 - t;
        wt[0] = 0.5 * (omt * omt);
        wt[1] = (0.5 + t) - t * t;
        wt[2] = 0.5 * (t * t);
        *base = r + 1;             /* r + shape - 1 */
    } else {
        const double t = (u - rd) + 0.5;
        const double sixth = 1.0 / 6.0, twothirds = 2.0 / 3.0;
        wt[0] = sixth + t * (-0.5 + (0.5 - t / 6.0) * t);
        wt[1] = twothirds + ((-1.0 + t / 2.0) * t) * t;
        wt[2] = sixth + t * (0.5 + (0.5 - t / 2.0) * t);
        wt[3] = ((t * t) * t) / 6.0;
        *base = r + 1;             /* r + shape - 2 */
    }
    return 1;
}

/* Kind of test per axis: x always int; y always double; z double only for shape 1. */
static int axis_test_kind(int axis, int ndim, int shape)
{
    if (axis == 0) return TEST_ON_INT;
    if (axis == 1) return TEST_ON_DOUBLE;
    (void)ndim;
    return (shape == 1) ? TEST_ON_DOUBLE : TEST_ON_INT;
}

/*
 * grid: C-order padded array with dims bins[a] + 2*shape, caller-zeroed (or holding a
 * partial result to accumulate onto).  coords[a] has element stride strides[a].
 * w may be NULL (unweighted branch).  Returns the number of accepted particles, or -1.
 */
int64_t ptg_oracle_deposit(int ndim, int shape, int64_t n,
                           const double *const coords[3], const int64_t strides[3],
                           const double *w, int64_t wstride,
                           const double lo[3], const double inv[3], const int bins[3],
                           double *grid)
{
    if (ndim < 1 || ndim > 3 || shape < 0 || shape > 3) return -1;
    int64_t dims[3] = {1, 1, 1};
    for (int a = 0; a < ndim; ++a) dims[a] = (int64_t)bins[a] + 2 * shape;
    int64_t accepted = 0;

    if (shape == 0) {
        for (int64_t i = 0; i < n; ++i) {
            int64_t cell = 0;
            int ok = 1;
            for (int a = 0; a < ndim; ++a) {
                const double u = (coords[a][i * strides[a]] - lo[a]) * inv[a];
                if (!(u > 0.0 && u < (double)bins[a])) { ok = 0; break; }
                cell = cell * dims[a] + (int64_t)(int)u;
            }
            if (!ok) continue;
            grid[cell] += w ? w[i * wstride] : 1.0;
            ++accepted;
        }
        return accepted;
    }

    const int nw = shape + 1;
    for (int64_t i = 0; i < n; ++i) {
        int base[3] = {0, 0, 0};
        double wt[3][PTG_MAXW];
        int ok = 1;
        for (int a = 0; a < ndim && ok; ++a)
            ok = axis_stencil(shape, axis_test_kind(a, ndim, shape),
                              coords[a][i * strides[a]], lo[a], inv[a], bins[a], &base[a], wt[a]);
        if (!ok) continue;
        ++accepted;
        if (ndim == 1) {
            for (int ix = 0; ix < nw; ++ix) {
                const double c = wt[0][ix];
                grid[base[0] + ix] += w ? c * w[i * wstride] : c;
            }
        } else if (ndim == 2) {
            for (int ix = 0; ix < nw; ++ix)
                for (int iy = 0; iy < nw; ++iy) {
                    const double c = wt[0][ix] * wt[1][iy];
                    grid[(int64_t)(base[0] + ix) * dims[1] + (base[1] + iy)] +=
                        w ? c * w[i * wstride] : c;
                }
        } else {
            for (int ix = 0; ix < nw; ++ix)
                for (int iy = 0; iy < nw; ++iy)
                    for (int iz = 0; iz < nw; ++iz) {
                        const double c = (wt[0][ix] * wt[1][iy]) * wt[2][iz];
                        grid[((int64_t)(base[0] + ix) * dims[1] + (base[1] + iy)) * dims[2]
                             + (base[2] + iz)] += w ? c * w[i * wstride] : c;
                    }
        }
    }
    return accepted;
}

/* Sum of the weights of accepted particles (the "total deposited weight" the
 * north-star tolerance is stated against), in particle order. */
double ptg_oracle_accepted_weight(int ndim, int shape, int64_t n,
                                  const double *const coords[3], const int64_t strides[3],
                                  const double *w, int64_t wstride,
                                  const double lo[3], const double inv[3], const int bins[3],
                                  double *abs_sum_out)
{
    double total = 0.0, atotal = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        int ok = 1;
        for (int a = 0; a < ndim && ok; ++a) {
            const double d = coords[a][i * strides[a]];
            if (shape == 0) {
                const double u = (d - lo[a]) * inv[a];
                ok = (u > 0.0 && u < (double)bins[a]);
            } else {
                int b; double wt[PTG_MAXW];
                ok = axis_stencil(shape, axis_test_kind(a, ndim, shape), d, lo[a], inv[a], bins[a], &b, wt);
            }
        }
        if (!ok) continue;
        const double wi = w ? w[i * wstride] : 1.0;
        total += wi;
        atotal += wi < 0 ? -wi : wi;
    }
    if (abs_sum_out) *abs_sum_out = atotal;
    return total;
}
