/* CPU restatement of ToeplitzMatrices.jl's direct solvers -- TEST INFRASTRUCTURE.
 *
 * Follows /root/reference/src/directLinearSolvers.jl line by line:
 *   reverse_dot        :137-145
 *   reverse_increment! :148-168   (alias-safe pairwise path when x === y)
 *   durbin!            :15-28
 *   levinson!          :100-121
 *   levinson(T, b)     :123-134   (normalise by the diagonal, un-normalise x)
 * plus batched drivers (columns = independent systems) that the reference
 * does not have: they only loop the single-system routine, with
 * OpenMP over systems so the CPU baseline can use every host core.
 *
 * Parity unpinned: see oracle/__init__.py.  Only tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 * Build: oracle/build_oracle.py (gcc -O2 -fopenmp -shared).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* :137-145 -- sequential left-to-right accumulation like the @simd-less form */
static double reverse_dot(const double *x, const double *y, int64_t n)
{
    double d = 0.0;
    for (int64_t i = 0; i < n; ++i)
        d += x[i] * y[n - 1 - i];
    return d;
}

/* :148-168 */
static void reverse_increment(double *x, const double *y, double alpha, int64_t n)
{
    if (x == y) {
        for (int64_t i = 0; i < n / 2; ++i) {
            double y_i = y[i];
            double z_i = y[n - 1 - i];
            x[i] += alpha * z_i;
            x[n - 1 - i] += alpha * y_i;
        }
        if (n & 1) {
            int64_t i = n / 2;
            x[i] += alpha * y[i];
        }
    } else {
        for (int64_t i = 0; i < n; ++i)
            x[i] += alpha * y[n - 1 - i];
    }
}

/* durbin!(y, r), length(r) == length(y) == n   (:15-28) */
void oracle_durbin(double *y, const double *r, int64_t n)
{
    y[0] = -r[0];
    double alpha = -r[0], beta = 1.0;
    for (int64_t k = 1; k < n; ++k) {
        beta *= (1 - alpha * alpha);
        alpha = -(r[k] + reverse_dot(r, y, k)) / beta;
        reverse_increment(y, y, alpha, k);
        y[k] = alpha;
    }
}

/* levinson!(x, r, b, y), n = length(r)+1 = length(x) = length(b) = length(y)  (:100-121) */
void oracle_levinson(double *x, const double *r, const double *b, double *y, int64_t n)
{
    y[0] = -r[0];
    x[0] = b[0];
    double alpha = -r[0], beta = 1.0;
    for (int64_t k = 1; k < n; ++k) {
        beta *= (1 - alpha * alpha);
        double mu = (b[k] - reverse_dot(r, x, k)) / beta;
        reverse_increment(x, y, mu, k);
        x[k] = mu;
        if (k < n - 1) {
            alpha = -(r[k] + reverse_dot(r, y, k)) / beta;
            reverse_increment(y, y, alpha, k);
            y[k] = alpha;
        }
    }
}

/* levinson(T::SymmetricToeplitz, b): vc has n entries  (:123-134) */
void oracle_levinson_sym(double *x, const double *vc, const double *b, int64_t n)
{
    double r0 = vc[0];
    double *r = (double *)malloc(sizeof(double) * (size_t)(n > 1 ? n - 1 : 1));
    double *y = (double *)calloc((size_t)n, sizeof(double));
    for (int64_t i = 0; i + 1 < n; ++i)
        r[i] = (r0 != 1.0) ? vc[i + 1] / r0 : vc[i + 1];
    memset(x, 0, sizeof(double) * (size_t)n);
    if (n == 1) {
        x[0] = b[0];
    } else {
        oracle_levinson(x, r, b, y, n);
    }
    if (r0 != 1.0)
        for (int64_t i = 0; i < n; ++i)
            x[i] /= r0;
    free(r);
    free(y);
}

/* batched: column j of R ((n-1) x nsys, ld ldr), B, X (n x nsys) is system j */
void oracle_levinson_batched(double *X, int64_t ldx, const double *R, int64_t ldr,
                             const double *B, int64_t ldb, int64_t n, int64_t nsys,
                             int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        double *y = (double *)malloc(sizeof(double) * (size_t)n);
#pragma omp for schedule(static)
        for (int64_t j = 0; j < nsys; ++j) {
            memset(y, 0, sizeof(double) * (size_t)n);
            memset(X + j * ldx, 0, sizeof(double) * (size_t)n);
            oracle_levinson(X + j * ldx, R + j * ldr, B + j * ldb, y, n);
        }
        free(y);
    }
}

void oracle_durbin_batched(double *Y, int64_t ldy, const double *R, int64_t ldr,
                           int64_t n, int64_t nsys, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < nsys; ++j) {
        memset(Y + j * ldy, 0, sizeof(double) * (size_t)n);
        oracle_durbin(Y + j * ldy, R + j * ldr, n);
    }
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
