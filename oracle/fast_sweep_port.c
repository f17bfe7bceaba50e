/*
 * oracle/fast_sweep_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the eikonal fast-sweeping rupture-onset solver that
 * the reference runs once per chain per Metropolis step
 * (reference: beat/fast_sweeping/fast_sweep_ext.c:120-206 `fast_sweep`, cell
 * update :77-118 `upwind`, local solver :65-75 `eq_solve`; called through
 * beat/pytensorf.py:443-482 `Sweeper.perform`).
 *
 * It exists so the CUDA kernel can be checked on machines where /root/reference
 * is absent (the GPU box).  It is itself pinned against the reference's own
 * compiled C (oracle/_ref/fast_sweep_ext*.so, built by oracle/Makefile from the
 * sources where they lie) by tests/test_oracle_fast_sweep.py and against the
 * committed golden vectors under tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this file's shared object.
 *
 * Restatement notes (all deliberate):
 *   - the reference writes sqrt as pow(x, 0.5) and the square as pow(d, 2.0);
 *     here sqrt(x) and d*d are used (the CUDA kernel does the same, so CUDA vs
 *     this port is bit-exact).  glibc's pow(x, 0.5) differs from the correctly
 *     rounded sqrt by 1 ulp for ~0.09 % of arguments (measured in this
 *     container), so against the compiled reference this port is bit-identical
 *     on ~98.5 % of random grids and within 2 ulp on the rest; the start-time
 *     INDICES derived from it (beat/ffi/base.py:506-517) are identical in all
 *     tested cases.  The reference's own cross-implementation test only asks
 *     for atol 1e-6 (test/test_fastsweep.py:131-133);
 *   - grid convention: `n_rows` x `n_cols`, flat index row*n_cols + col.  The
 *     production caller passes rows = dip, cols = strike
 *     (beat/pytensorf.py:475-482 swaps the C argument names on purpose);
 *   - initial field is +inf with 0 at the hypocentre (:140-149), NaNs from
 *     inf-inf fall through the same comparison directions as the reference
 *     (`x < y ? x : y` keeps y when x is NaN);
 *   - convergence: sum over cells, in flat order, of (T - T_prev)^2 <= 0.1
 *     (:127,151,198-201).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline double lesser(double x, double y) { return (x < y) ? x : y; }

/* one Gauss-Seidel cell update; returns the value to store at (r, c) */
static double relax_cell(const double *T, const double *slow, double h,
                         long r, long c, long n_rows, long n_cols)
{
    long rm = r > 0 ? r - 1 : 0;
    long rp = r + 1 < n_rows ? r + 1 : n_rows - 1;
    long cm = c > 0 ? c - 1 : 0;
    long cp = c + 1 < n_cols ? c + 1 : n_cols - 1;

    double a = lesser(T[rm * n_cols + c], T[rp * n_cols + c]);  /* along rows    */
    double b = lesser(T[r * n_cols + cm], T[r * n_cols + cp]);  /* along columns */
    double f = slow[r * n_cols + c];
    double cur = T[r * n_cols + c];
    double cand;

    if (fabs(a - b) >= f * h) {
        cand = lesser(a, b);
        cand += f * h;
    } else {
        double d = a - b;
        cand = a + b + sqrt(2.0 * f * f * h * h - d * d);
        cand /= 2.0;
    }
    return (cand < cur) ? cand : cur;
}

/* returns the number of outer iterations performed */
int fsport_sweep(const double *slow, double *T, double h,
                 long hyp_row, long hyp_col, long n_rows, long n_cols)
{
    const double eps = 0.1;
    long n = n_rows * n_cols;
    long r, c, k;
    int iters = 0;
    double err = 1.0e6;
    double *prev = (double *)malloc((size_t)n * sizeof(double));

    for (k = 0; k < n; k++) T[k] = INFINITY;
    T[hyp_row * n_cols + hyp_col] = 0.0;

    while (err > eps) {
        memcpy(prev, T, (size_t)n * sizeof(double));

        for (r = 0; r < n_rows; r++)            /* rows up,   cols up   */
            for (c = 0; c < n_cols; c++)
                T[r * n_cols + c] = relax_cell(T, slow, h, r, c, n_rows, n_cols);
        for (r = n_rows - 1; r >= 0; r--)       /* rows down, cols up   */
            for (c = 0; c < n_cols; c++)
                T[r * n_cols + c] = relax_cell(T, slow, h, r, c, n_rows, n_cols);
        for (r = n_rows - 1; r >= 0; r--)       /* rows down, cols down */
            for (c = n_cols - 1; c >= 0; c--)
                T[r * n_cols + c] = relax_cell(T, slow, h, r, c, n_rows, n_cols);
        for (r = 0; r < n_rows; r++)            /* rows up,   cols down */
            for (c = n_cols - 1; c >= 0; c--)
                T[r * n_cols + c] = relax_cell(T, slow, h, r, c, n_rows, n_cols);

        err = 0.0;
        for (k = 0; k < n; k++) {
            double d = T[k] - prev[k];
            err += d * d;
        }
        iters++;
    }
    free(prev);
    return iters;
}

/* batch driver: slow[B, n], hyp_row[B], hyp_col[B] -> T[B, n]; iters[B] optional */
void fsport_sweep_batch(const double *slow, double *T, double h,
                        const long *hyp_row, const long *hyp_col,
                        long n_rows, long n_cols, long B, int *iters)
{
    long n = n_rows * n_cols, b;
    for (b = 0; b < B; b++) {
        int it = fsport_sweep(slow + b * n, T + b * n, h, hyp_row[b], hyp_col[b],
                              n_rows, n_cols);
        if (iters) iters[b] = it;
    }
}
