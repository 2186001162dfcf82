/* libtoeplitz_b200.so -- C ABI of the B200-native ToeplitzMatrices.jl hot path.
 *
 * This is the drop-in boundary: the entry points a Julia package extension
 * (ext/ToeplitzMatricesB200Ext.jl) binds with `ccall` to dispatch
 * LinearAlgebra.mul!/factorize/ldiv!/... on device-resident vectors.  Each entry
 * cites the reference method it replaces (paths relative to the reference repo,
 * ToeplitzMatrices.jl v0.8.5).  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *  - every function returns an int status (TB200_OK == 0); tb200_last_error() gives
 *    the thread-local message for the last non-zero status.
 *  - pointers are DEVICE pointers unless the name ends in _host or the parameter is
 *    documented as host.  Matrices are column-major with leading dimension ld* in
 *    ELEMENTS of their own dtype.
 *  - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it except
 *    where stated (tb200_cgs/tb200_cg read convergence scalars back every iteration).
 *  - dtype: element type of a vector/matrix, TB200_F32/F64/C32/C64.  The working
 *    precision of a handle is that of its matrix (S = promote_type(float(T),
 *    ComplexF32), src/linearalgebra.jl:125): vectors passed to a handle must have the
 *    same precision (real or complex); the host layer casts mixed inputs.
 *  - a handle owns its (read-only) spectrum plus one workspace set PER CALLING STREAM: it
 *    may be used concurrently from several streams / host threads (one stream per thread).
 *    The reference's factorization cannot -- every product shares its single `tmp` vector
 *    (src/ToeplitzMatrices.jl:97-101).  A product on a stream other than the one that
 *    factorized waits for the spectrum by event; no host synchronisation is needed.
 *  - CUDA graphs: after one warm-up call on the capturing stream (workspaces allocated, the
 *    factorization seen complete) tb200_mul / tb200_circ_ldiv enqueue kernels and event
 *    fork / joins only and can be captured; a capturing stream cannot wait for a factorization
 *    still running on another stream (TB200_CUDA_ERR with a message saying so).
 *  - device memory comes from the stream-ordered pool; tb200_destroy orders the release
 *    after everything already queued on the streams that used the handle, so destroying
 *    right after an asynchronous call is safe.  The caller's vectors must outlive the
 *    work queued on them, as with any asynchronous CUDA call.
 *  - there is NO CPU fallback: unsupported inputs return TB200_UNSUPPORTED.
 */
#ifndef TOEPLITZ_B200_H
#define TOEPLITZ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum tb200_status {
    TB200_OK = 0,
    TB200_DIM_MISMATCH = 1,   /* -> Julia DimensionMismatch */
    TB200_BAD_ARG = 2,        /* -> ArgumentError */
    TB200_UNSUPPORTED = 3,    /* -> ErrorException */
    TB200_CUDA_ERR = 4,
    TB200_OOM = 5,
    TB200_NCCL_ERR = 6
};

enum tb200_dtype { TB200_F32 = 0, TB200_F64 = 1, TB200_C32 = 2, TB200_C64 = 3 };

enum tb200_kind {
    TB200_TOEPLITZ = 0,   /* Toeplitz(vc, vr)              src/toeplitz.jl:7-18        */
    TB200_SYMMETRIC = 1,  /* SymmetricToeplitz(v)          src/special.jl:44-76        */
    TB200_CIRCULANT = 2,  /* Circulant(v)                  src/special.jl:44-76        */
    TB200_LOWER = 3,      /* LowerTriangularToeplitz(v)                                */
    TB200_UPPER = 4,      /* UpperTriangularToeplitz(v)                                */
    TB200_HANKEL = 5      /* Hankel(v, (m,n))              src/hankel.jl:2-12          */
};

enum tb200_specop { TB200_OP_INV = 0, TB200_OP_PINV = 1, TB200_OP_SQRT = 2, TB200_OP_CONJ = 3, TB200_OP_COPY = 4 };

typedef struct tb200_fact* tb200_handle_t;

const char* tb200_version(void);
const char* tb200_last_error(void);

/* ---- factorize ------------------------------------------------------------------
 * Replaces factorize(::Toeplitz) src/linearalgebra.jl:122-131, (::SymmetricToeplitz)
 * :140-150, (::Circulant) :184-192, (::Lower/UpperTriangularToeplitz) :415-435, and for
 * TB200_HANKEL the composition factorize(reverse(H, dims=2)) src/hankel.jl:112-121,133-137
 * (the x-reversal is then applied inside tb200_mul).
 * vc: first column (m entries); vr: first row (n entries, TOEPLITZ only, else NULL).
 * SYMMETRIC/CIRCULANT/LOWER/UPPER: vc = v (n entries), m == n.  HANKEL: vc = v (m+n-1).
 * The handle stores the spectrum of the zero-padded power-of-two circulant embedding
 * (N' >= m+n-1); CIRCULANT keeps exact n-point semantics (n a power of two is the fast
 * path; other n use a chirp-z (Bluestein) plan built on the same kernels).            */
int tb200_factorize(int kind, int dtype, int64_t m, int64_t n, const void* vc, const void* vr,
                    void* stream, tb200_handle_t* out);
int tb200_factorize_host(int kind, int dtype, int64_t m, int64_t n, const void* vc_host,
                         const void* vr_host, void* stream, tb200_handle_t* out);
int tb200_destroy(tb200_handle_t h);
/* size(F) src/special.jl:262-263; embedding length; dtype of the matrix */
int tb200_info(tb200_handle_t h, int64_t* m, int64_t* n, int64_t* n_embed, int* dtype, int* kind);

/* ---- mul! ------------------------------------------------------------------------
 * y[:, j] = alpha * maybereal(A * x[:, j]) + beta * y[:, j],  j < ncols.
 * Replaces mul!(y, ::ToeplitzFactorization, x, alpha, beta) src/linearalgebra.jl:42-80 and
 * its matrix form :88-99 (the serial column loop becomes the batch axis), and the Hankel
 * forms src/hankel.jl:133-137.  alpha/beta: (re, im) pairs, host.  beta == 0 means y is
 * not read (:68-71).  A real ydtype requires a real matrix, real x and real alpha/beta
 * (then the real part is taken, src/ToeplitzMatrices.jl:114-115).                      */
int tb200_mul(tb200_handle_t h, void* y, int64_t ldy, int ydtype, const void* x, int64_t ldx, int xdtype,
              int64_t ncols, const double alpha[2], const double beta[2], void* stream);
/* same, HOST buffers: column chunks travel host->device->host on three internal streams
 * (upload / compute / download of consecutive chunks overlap); waits for the handle's
 * spectrum by event.  Copies are issued straight from / to the caller's memory: pinned
 * (page-locked) host matrices run at PCIe speed, pageable ones at the driver's staged-copy
 * speed (both measured in bench.py).  Synchronous.                                     */
int tb200_mul_host(tb200_handle_t h, void* y_host, int64_t ldy, int ydtype, const void* x_host, int64_t ldx,
                   int xdtype, int64_t ncols, const double alpha[2], const double beta[2]);
/* N = m+n-1 < 512 direct O(mn) product, no factorization: src/linearalgebra.jl:19-34   */
int tb200_mul_direct(int kind, int dtype, int64_t m, int64_t n, const void* vc, const void* vr,
                     void* y, int64_t ldy, int ydtype, const void* x, int64_t ldx, int xdtype,
                     int64_t ncols, const double alpha[2], const double beta[2], void* stream);

/* ---- Circulant spectral family ----------------------------------------------------
 * ldiv!(C::CirculantFactorization, b) src/linearalgebra.jl:225-242 (+ matrix form :102-110,
 * 3-arg ldiv!(x, F, b)): x = maybereal(ifft(fft(b) ./ spectrum)); x may alias b.         */
int tb200_circ_ldiv(tb200_handle_t h, void* x, int64_t ldx, const void* b, int64_t ldb, int dtype,
                    int64_t ncols, void* stream);
/* eigvals(C) :286-287 -- natural-order spectrum, n complex values of the handle precision */
int tb200_circ_spectrum(tb200_handle_t h, void* out, void* stream);
/* inv(F) :250-253, adjoint(F) :214-215, and the spectral halves of pinv :276-284 /
 * sqrt :311-315 -> a NEW handle (own storage; the reference shares tmp/dft).            */
int tb200_circ_map(tb200_handle_t h, int op, double tol, void* stream, tb200_handle_t* out);
/* A*B on factorizations :197-211 -> new handle holding specA .* specB                   */
int tb200_circ_mul(tb200_handle_t a, tb200_handle_t b, void* stream, tb200_handle_t* out);
/* vc = maybereal(ifft(spectrum)) -> first column of a Circulant (:208-210,247-248,282-283,313) */
int tb200_circ_to_vc(tb200_handle_t h, void* vc_out, int out_dtype, void* stream);

/* strang(A) :255-266 (which = 0) / chan(A) :267-274 (which = 1): v_out has n entries     */
int tb200_precond_vector(int which, int kind, int dtype, int64_t n, const void* vc, const void* vr,
                         void* v_out, void* stream);

/* ---- batched Levinson / Durbin ------------------------------------------------------
 * Column j of R/B/X/Y is one independent system (the batch axis is new).
 * levinson!(x, r, b) src/directLinearSolvers.jl:100-121: T = SymToeplitz([1; r]),
 * r has n-1 entries, b and x have n.  diag (nsys entries) or NULL: the normalising
 * wrapper levinson(T::SymmetricToeplitz, b) :123-134 (r /= diag, x /= diag).
 * durbin!(y, r) :15-28: r and y have n entries.  dtype TB200_F32 / TB200_F64.            */
int tb200_levinson_batched(int dtype, int64_t n, int64_t nsys, const void* r, int64_t ldr,
                           const void* b, int64_t ldb, void* x, int64_t ldx, const void* diag,
                           void* stream);
int tb200_durbin_batched(int dtype, int64_t n, int64_t nsys, const void* r, int64_t ldr,
                         void* y, int64_t ldy, void* stream);
/* Many right-hand sides, ONE matrix: X = SymmetricToeplitz(vc) \ B, the column loop of
 * ext/ToeplitzMatricesStatsBaseExt.jl:15-24.  The y-recursion of levinson! (src/directLinearSolvers.jl:114-118)
 * is independent of b: it runs once, the per-column work drops from 5k to 2k FMAs per step.  vc: n entries
 * (vc[0] = diagonal, :123-134); B, X: n x nrhs; 2 <= n <= 512 (larger n: tb200_levinson_batched).           */
int tb200_levinson_multirhs(int dtype, int64_t n, int64_t nrhs, const void* vc, const void* B, int64_t ldb,
                            void* X, int64_t ldx, void* stream);

/* trench(T::SymmetricToeplitz) src/directLinearSolvers.jl:36-85: B (n x n, column-major, ldb) = inv(T) for the
 * SPD matrix T = diag * SymToeplitz([1; r]); r has n-1 entries (already divided by diag, :51-53), real dtype.
 * Durbin + the anti-diagonal recurrence fill; both triangles are written (the reference fills the upper one and
 * wraps it in Symmetric).                                                                                     */
int tb200_trench(int dtype, int64_t n, const void* r, void* B, int64_t ldb, double diag, void* stream);
/* smallinv(::LowerTriangularToeplitz) src/linearalgebra.jl:337-349: out = first column of inv(L(v)), n <= 64.
 * Base case of the recursive-doubling triangular inverse (:350-396), whose products go through tb200_mul*.    */
int tb200_tri_smallinv(int dtype, int64_t n, const void* v, void* out, void* stream);

/* ---- iterative solvers ---------------------------------------------------------------
 * IterativeLinearSolvers.cgs src/iterativeLinearSolvers.jl:99-215 and cg :9-97 with A a
 * factorized structured matrix and M a Circulant factorization (factorize(strang(A)) /
 * factorize(chan(A)), src/linearalgebra.jl:133-136,152,399-402) or NULL (identity).
 * x (in: start vector, out: solution) and b have n entries of `dtype`.  Outputs (host):
 * err = final relative residual, iter, flag (0 converged, 1 max_it, -1 breakdown rho == 0).
 * Synchronous on `stream`.                                                               */
int tb200_cgs(tb200_handle_t A, tb200_handle_t M, void* x, const void* b, int dtype, int64_t n,
              int max_it, double tol, double* err, int* iter, int* flag, void* stream);
int tb200_cg(tb200_handle_t A, tb200_handle_t M, void* x, const void* b, int dtype, int64_t n,
             int max_it, double tol, double* err, int* iter, int* flag, void* stream);

/* The matrix forms ldiv!(A, B::Matrix) src/linearalgebra.jl:102-110 (a serial column loop over the
 * solvers above) with all columns iterating together: X (in: start vectors, out: solutions) and
 * B are n x nrhs, column-major.  Per column the arithmetic is that of tb200_cgs / tb200_cg; a
 * column that converges, breaks down (cgs, rho == 0) or returns early is frozen while the others
 * go on.  err, iter, flag: host arrays of nrhs entries (flag 2 = cg's value-less early return,
 * src/iterativeLinearSolvers.jl:57-59).  Synchronous on `stream`.                              */
int tb200_cgs_batched(tb200_handle_t A, tb200_handle_t M, void* X, int64_t ldx, const void* B, int64_t ldb,
                      int dtype, int64_t n, int64_t nrhs, int max_it, double tol, double* err, int* iter,
                      int* flag, void* stream);
int tb200_cg_batched(tb200_handle_t A, tb200_handle_t M, void* X, int64_t ldx, const void* B, int64_t ldb,
                     int dtype, int64_t n, int64_t nrhs, int max_it, double tol, double* err, int* iter,
                     int* flag, void* stream);

/* ---- multi-GPU plumbing ----------------------------------------------------------------
 * Batches shard by columns / systems with no data-path collective; the only
 * exchange is the spectrum, once per factorization: the root factorizes, the other ranks
 * create an empty handle of the same shape and the spectrum (`bytes` bytes at `ptr`) is
 * broadcast over NCCL / NVLink -- by tb200_bcast_spectrum, or by the host's own collective
 * followed by tb200_spectrum_ready(h, stream_that_wrote_it).                            */
int tb200_factorize_empty(int kind, int dtype, int64_t m, int64_t n, void* stream, tb200_handle_t* out);
int tb200_spectrum_buffer(tb200_handle_t h, void** ptr, int64_t* bytes);
int tb200_spectrum_ready(tb200_handle_t h, void* stream);

/* NCCL communicator (libnccl.so.2 is bound at run time with dlopen; TB200_NCCL_LIB overrides the path).
 *  tb200_comm_init       one process driving ndev GPUs (devs = NULL: 0..ndev-1), ncclCommInitAll
 *  tb200_comm_unique_id  128-byte id made on rank 0 and shipped to the other processes by the host
 *  tb200_comm_init_rank  one process per GPU (the current device), ncclCommInitRank
 *  tb200_bcast_spectrum  handles[i] lives on local device i of the communicator (nhandles = ndev for
 *                        tb200_comm_init, 1 for tb200_comm_init_rank); root = rank owning the spectrum;
 *                        streams[i] (or NULL = default stream) carries the ncclBroadcast and the
 *                        handle's readiness event.  Asynchronous.                                     */
typedef struct tb200_comm* tb200_comm_t;
int tb200_comm_init(int ndev, const int* devs, tb200_comm_t* out);
int tb200_comm_unique_id(void* id128);
int tb200_comm_init_rank(int nranks, int rank, const void* id128, tb200_comm_t* out);
int tb200_comm_size(tb200_comm_t c, int* nranks, int* nlocal);
int tb200_bcast_spectrum(tb200_comm_t c, tb200_handle_t* handles, int nhandles, int root, void** streams);
int tb200_comm_destroy(tb200_comm_t c);

/* Container algebra on the stored coefficient vectors (the reference's generic broadcasts on A.vc / A.vr / A.v:
 * src/toeplitz.jl:66-140 tril, triu, +, -, scalar product, conj, copyto!, fill!, lmul!, rmul!, src/special.jl:13-33,77-84,223-248,
 * src/hankel.jl:75-130 incl. reverse).  op: 0 out = alpha*a + beta*b (b may be NULL), 1 conj(a), 2 real(a), 3 imag(a),
 * 4 fill(alpha).  Every operand has its own TB200_* dtype (promotion happens in the kernel); rev_a / rev_b read the
 * operand back to front.  Asynchronous on `stream`.                                                                */
int tb200_vec_op(int op, int64_t n, const void* a, int dtype_a, int rev_a, const void* b, int dtype_b, int rev_b, void* out,
                 int dtype_out, const double alpha[2], const double beta[2], void* stream);
/* `==` on two stored vectors (src/toeplitz.jl:114-121, src/hankel.jl:96): *equal = 1 iff all n entries agree.
 * b == NULL compares against zero (iszero / isdiag).  Synchronises `stream` (returns a host scalar).               */
int tb200_vec_equal(int64_t n, const void* a, int dtype_a, const void* b, int dtype_b, int* equal, void* stream);

/* tuning / introspection: number of kernel launches issued by this library so far      */
int64_t tb200_launch_count(void);
/* process-wide tuning knobs (defaults are what the measurements in DESIGN.md were taken with):
 *   "group_bytes"      two-level workspace budget per column group (default 48 MiB: one 32-MiB column pair at n = 2^20 Float64)
 *   "overlap_streams"  internal streams that alternate column groups (default 2; 0 / 1 = off)
 *   "l2_persist_mb"    persisting-L2 carve-out for the workspaces of products with several column groups whose
 *                      workspaces and spectrum fit L2 together (default 64; 0 = off)
 *   "prune_half"       skip the padding half of the embedding in the strided passes (default 1)
 *   "pdl"              programmatic dependent launch of passes B / C: 0 off, 1 products without group overlap (default), 2 always
 *   "pair_columns"     real columns of a matrix product share one complex transform two by two (default 1).  The pair
 *                      is coupled numerically: a NaN / Inf in one column reaches its partner, and each column's error is
 *                      relative to the larger norm of the pair.  0 transforms every real column alone (the reference's
 *                      independent column loop, src/linearalgebra.jl:95-97) at half the throughput.
 *   "rows_direct"      pass B addresses its workspace rows directly instead of through the column views (default 1)
 *   "tma"              TMA tensor-box staging of the strided tiles: bit 0 pass A, bit 1 pass C (default 0)
 *   "lev_gen"          Levinson / Durbin in generator form (default 1; 0 = the dot-product kernels)
 *   "time_passes"      bracket every pass with CUDA events for tb200_pass_times (default 0)
 *   "single_max_f64/f32", "l2max_f64/f32", "force_l1", "ta_cap"   plan shape overrides (experiments)
 *   "host_chunk_bytes", "krylov_work_bytes"                       staging / workspace budgets of tb200_mul_host and the batched solvers */
int tb200_set_option(const char* name, int64_t value);
/* with option "time_passes" = 1 every FFT pass is bracketed by CUDA events on its stream;
 * this drains them: summed milliseconds and launch counts for {cols A, rows B, cols C,
 * single-kernel rows}.  Synchronises on the recorded events.                            */
int tb200_pass_times(double ms[4], int64_t counts[4]);
/* same with n <= 5 entries: entry 4 = the fused "pass C of one group + pass A of the next" launches
 * (option "ca_fuse"); tb200_pass_times reports those together with cols C.             */
int tb200_pass_times_ex(double* ms, int64_t* counts, int n);

#ifdef __cplusplus
}
#endif
#endif /* TOEPLITZ_B200_H */
