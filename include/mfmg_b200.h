/*
 * mfmg_b200.h -- C ABI of the B200-native (sm_100a) multigrid V-cycle apply.
 *
 * This is the drop-in boundary for mfmg's device operator path: every entry point names the
 * reference interface it replaces (paths relative to the ORNL-CEES/mfmg tree).  Plain pointers and
 * sizes only; opaque handles; every function returns an int status (0 = MFMGB_OK) and never
 * throws; work is stream-ordered on the context's stream; results are visible to the host after
 * mfmgb_ctx_synchronize() (the *_host entry points synchronise before returning, matching the
 * reference's synchronous semantics).  No cuSPARSE / cuSOLVER / AMGX behind any of them.
 *
 * Vectors are raw device pointers to double (the storage of
 * dealii::LinearAlgebra::distributed::Vector<double, MemorySpace::CUDA>::get_values()).
 * Matrices are CSR with 64-bit row offsets on the host side of the ABI (the reference's int
 * row_ptr overflows at 513^3 Q1) and int32 local column indices.
 */
#ifndef MFMG_B200_H
#define MFMG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

#if defined(_WIN32)
#define MFMGB_API
#else
#define MFMGB_API __attribute__((visibility("default")))
#endif

  typedef struct mfmgb_ctx mfmgb_ctx;
  typedef struct mfmgb_csr mfmgb_csr;
  typedef struct mfmgb_jacobi mfmgb_jacobi;
  typedef struct mfmgb_dense mfmgb_dense;
  typedef struct mfmgb_mf mfmgb_mf;
  typedef struct mfmgb_hierarchy mfmgb_hierarchy;

  enum
  {
    MFMGB_OK = 0,
    MFMGB_ERR_INVALID = 1,         /* bad argument (reference: ASSERT / ASSERT_THROW, exceptions.hpp:44-63) */
    MFMGB_ERR_CUDA = 2,            /* a CUDA call failed (reference: ASSERT_CUDA, exceptions.hpp:204-228; checked here in ALL build types) */
    MFMGB_ERR_NOT_IMPLEMENTED = 3, /* reference: NotImplementedExc, exceptions.hpp:65-84 */
    MFMGB_ERR_SINGULAR = 4,        /* zero pivot in the dense factorisation / missing diagonal */
    MFMGB_ERR_NCCL = 5,
    MFMGB_ERR_NOT_CONVERGED = 6    /* PCG hit max_it (dealii::SolverControl::NoConvergence) */
  };

  /* ---- context: replaces mfmg::CudaHandle (source/cuda/cuda_handle.cu:17-56), minus the library handles ---- */
  /* stream: a cudaStream_t to launch on, or NULL to create a private non-blocking stream. */
  MFMGB_API int mfmgb_ctx_create(int device, void *stream, mfmgb_ctx **out);
  MFMGB_API int mfmgb_ctx_destroy(mfmgb_ctx *ctx);
  MFMGB_API int mfmgb_ctx_synchronize(mfmgb_ctx *ctx);
  MFMGB_API void *mfmgb_ctx_stream(mfmgb_ctx *ctx);
  /* last error message of this context (or of the calling thread when ctx == NULL) */
  MFMGB_API const char *mfmgb_last_error(mfmgb_ctx *ctx);
  /* number of kernels this context has launched so far (bench.py's gpu_launches claim) */
  MFMGB_API int64_t mfmgb_ctx_launch_count(mfmgb_ctx *ctx);
  MFMGB_API const char *mfmgb_version(void);

  /* ---- raw device memory: replaces cuda_malloc / cuda_free / cuda_mem_copy_to_dev / cuda_mem_copy_to_host
   *      (include/mfmg/cuda/utils.cuh:66-99), the helpers the reference's callers use to fill the arrays they hand to
   *      SparseMatrixDevice's take-ownership constructor.  Plain cudaMalloc memory: mfmgb_csr_adopt_device accepts it.
   *      ctx may be NULL here (the reference's helpers take no handle). ---- */
  MFMGB_API int mfmgb_dev_malloc(mfmgb_ctx *ctx, int64_t bytes, void **out);
  MFMGB_API int mfmgb_dev_free(mfmgb_ctx *ctx, void *ptr);
  MFMGB_API int mfmgb_dev_upload(mfmgb_ctx *ctx, void *dst_dev, const void *src_host, int64_t bytes);
  MFMGB_API int mfmgb_dev_download(mfmgb_ctx *ctx, const void *src_dev, void *dst_host, int64_t bytes);

  /* ---- device vectors: replaces cuda_malloc/cuda_free/cuda_mem_copy_to_{dev,host} (include/mfmg/cuda/utils.cuh:66-99)
   *      and the deal.II CUDA-vector ops used on the path (hierarchy.hpp:258,286,302; cuda_smoother.cu:50-59) ---- */
  MFMGB_API int mfmgb_vec_alloc(mfmgb_ctx *ctx, int64_t n, double **out);
  MFMGB_API int mfmgb_vec_free(mfmgb_ctx *ctx, double *v);
  MFMGB_API int mfmgb_vec_upload(mfmgb_ctx *ctx, double *dst_dev, const double *src_host, int64_t n);
  MFMGB_API int mfmgb_vec_download(mfmgb_ctx *ctx, const double *src_dev, double *dst_host, int64_t n);
  MFMGB_API int mfmgb_vec_fill(mfmgb_ctx *ctx, double *v, double value, int64_t n);
  MFMGB_API int mfmgb_vec_copy(mfmgb_ctx *ctx, double *dst, const double *src, int64_t n);
  MFMGB_API int mfmgb_vec_axpy(mfmgb_ctx *ctx, double *y, double a, const double *x, int64_t n); /* y += a x  (Vector::add) */
  MFMGB_API int mfmgb_vec_dot(mfmgb_ctx *ctx, const double *a, const double *b, int64_t n, double *result_host);

  /* ---- CSR matrix: replaces mfmg::SparseMatrixDevice<double> (include/mfmg/cuda/sparse_matrix_device.cuh:28-104)
   *      and convert_matrix (source/cuda/utils.cu:39-168) ---- */
  /* copies host arrays to the device; the caller keeps its arrays (convert_matrix semantics). */
  MFMGB_API int mfmgb_csr_upload(mfmgb_ctx *ctx, int64_t n_rows, int64_t n_cols, const int64_t *rowptr,
                                 const int32_t *col, const double *val, mfmgb_csr **out);
  /* same, with the reference's int row_ptr (sparse_matrix_device.cuh:94). */
  MFMGB_API int mfmgb_csr_upload_i32(mfmgb_ctx *ctx, int64_t n_rows, int64_t n_cols, const int32_t *rowptr,
                                     const int32_t *col, const double *val, mfmgb_csr **out);
  /* TAKES OWNERSHIP of three cudaMalloc'ed arrays, exactly like SparseMatrixDevice's ctor
   * (include/mfmg/cuda/sparse_matrix_device.templates.cuh:244-272); freed in mfmgb_csr_destroy. */
  MFMGB_API int mfmgb_csr_adopt_device(mfmgb_ctx *ctx, int64_t n_rows, int64_t n_cols, int64_t nnz, double *val_dev,
                                       int32_t *col_dev, int32_t *rowptr_dev, mfmgb_csr **out);
  MFMGB_API int mfmgb_csr_destroy(mfmgb_ctx *ctx, mfmgb_csr *A);
  MFMGB_API int mfmgb_csr_info(const mfmgb_csr *A, int64_t *n_rows, int64_t *n_cols, int64_t *nnz);
  /* raw device arrays (SparseMatrixDevice::val_dev / column_index_dev, public members cuh:92-94) */
  MFMGB_API int mfmgb_csr_device_arrays(const mfmgb_csr *A, double **val_dev, int32_t **col_dev, void **rowptr_dev,
                                        int *rowptr_is_64);
  /* device -> host copy (round-trip check, tests/test_utils_device.cu:221-263) */
  MFMGB_API int mfmgb_csr_download(mfmgb_ctx *ctx, const mfmgb_csr *A, int64_t *rowptr, int32_t *col, double *val);
  /* explicit transpose, ascending columns, deterministic; replaces CudaMatrixOperator::transpose
   * (source/cuda/cuda_matrix_operator.cu:93-130), done once at setup. */
  MFMGB_API int mfmgb_csr_transpose(mfmgb_ctx *ctx, const mfmgb_csr *A, mfmgb_csr **out);
  /* lanes-per-row override for the vector-CSR kernels: 0 = choose from the mean row length */
  MFMGB_API int mfmgb_csr_set_lanes_per_row(mfmgb_csr *A, int lanes);
  MFMGB_API int mfmgb_csr_get_lanes_per_row(const mfmgb_csr *A);
  /* kernel family behind mfmgb_spmv & co.: -1 = automatic, 0 = vector-CSR with direct global loads, 1 = tile-streamed
   * (col/val/rowptr staged in shared memory by TMA bulk copies, csrc/csr_tile.cu); both give bit-identical results */
  MFMGB_API int mfmgb_csr_set_kernel(mfmgb_csr *A, int kernel);
  MFMGB_API int mfmgb_csr_get_kernel(const mfmgb_csr *A);

  /* y = A x : SparseMatrixDevice::vmult (sparse_matrix_device.templates.cuh:351-371), CudaMatrixOperator::apply
   * NO_TRANS (source/cuda/cuda_matrix_operator.cu:80-91) */
  MFMGB_API int mfmgb_spmv(mfmgb_ctx *ctx, const mfmgb_csr *A, const double *x, double *y);
  /* r = A x - b : the negative residual of Hierarchy::apply (include/mfmg/common/hierarchy.hpp:282-286), fused */
  MFMGB_API int mfmgb_residual_neg(mfmgb_ctx *ctx, const mfmgb_csr *A, const double *x, const double *b, double *r);
  /* b_c = R r : restrictor->apply (hierarchy.hpp:289-290) */
  MFMGB_API int mfmgb_restrict(mfmgb_ctx *ctx, const mfmgb_csr *R, const double *r, double *b_c);
  /* x -= P x_c with P = R^T stored explicitly: restrictor->apply(TRANS) + x.add(-1, x_corr) (hierarchy.hpp:297-302), fused */
  MFMGB_API int mfmgb_prolong_correct(mfmgb_ctx *ctx, const mfmgb_csr *P, const double *x_c, double *x);

  /* ---- Jacobi smoother: replaces mfmg::CudaSmoother (source/cuda/cuda_smoother.cu:39-60,99-172) ---- */
  /* builds D^-1 (kernel extract_inv_diag, cuda_smoother.cu:86-96). omega = 1 is the reference. */
  MFMGB_API int mfmgb_jacobi_setup(mfmgb_ctx *ctx, const mfmgb_csr *A, double omega, mfmgb_jacobi **out);
  /* same, from a given diagonal (matrix-free operators: LaplaceOperator::compute_diagonal, tests/laplace_matrix_free.hpp:75-98) */
  MFMGB_API int mfmgb_jacobi_setup_diag(mfmgb_ctx *ctx, const double *diag_dev, int64_t n, double omega, mfmgb_jacobi **out);
  MFMGB_API int mfmgb_jacobi_destroy(mfmgb_ctx *ctx, mfmgb_jacobi *J);
  MFMGB_API const double *mfmgb_jacobi_inv_diag(const mfmgb_jacobi *J);
  /* x <- x - omega D^-1 (A x - b), one fused kernel; x is updated in place (an internal
   * ping-pong buffer holds the previous iterate). */
  MFMGB_API int mfmgb_jacobi_apply(mfmgb_ctx *ctx, const mfmgb_jacobi *J, const mfmgb_csr *A, const double *b, double *x);
  /* out-of-place form: x_out = x_in - omega D^-1 (A x_in - b); x_out must not alias x_in */
  MFMGB_API int mfmgb_jacobi_apply_oop(mfmgb_ctx *ctx, const mfmgb_jacobi *J, const mfmgb_csr *A, const double *b,
                                       const double *x_in, double *x_out);
  /* x <- x - omega D^-1 r for a residual r = A x - b the caller formed with its own operator (user matrix-free
   * operators behind CudaMatrixFreeOperator): the update half of cuda_smoother.cu:52-59 */
  MFMGB_API int mfmgb_jacobi_apply_residual(mfmgb_ctx *ctx, const mfmgb_jacobi *J, const double *r, double *x);
  /* x = omega D^-1 b : the sweep when x == 0 on entry (hierarchy.hpp:253-259 then :277-279) */
  MFMGB_API int mfmgb_jacobi_apply_zero_guess(mfmgb_ctx *ctx, const mfmgb_jacobi *J, const double *b, double *x);

  /* ---- dense coarse solver: replaces mfmg::CudaSolver "lu_dense" (source/cuda/cuda_solver.cu:496-515 ->
   *      source/cuda/dealii_operator_device_helpers.cu:169-228), factorised ONCE instead of per apply ---- */
  MFMGB_API int mfmgb_dense_factor(mfmgb_ctx *ctx, const mfmgb_csr *A_c, mfmgb_dense **out);
  MFMGB_API int mfmgb_dense_destroy(mfmgb_ctx *ctx, mfmgb_dense *D);
  MFMGB_API int mfmgb_dense_solve(mfmgb_ctx *ctx, const mfmgb_dense *D, const double *b, double *x);
  MFMGB_API int64_t mfmgb_dense_size(const mfmgb_dense *D);
  /* number of row interchanges the factorisation performed (diagnostic) */
  MFMGB_API int64_t mfmgb_dense_num_swaps(const mfmgb_dense *D);
  /* how mfmgb_dense_solve applies the factorisation: 0 = one GEMV with A^-1 = U^-1 L^-1 P formed at setup (well-conditioned
   * operators: a substitution solve would be ~n dependent steps per cycle), 1 = forward / backward substitution with the
   * kept factors, getrs' operation order (chosen when min |u_kk| / max |u_kk| < 1e-12, returned in *pivot_ratio: the
   * reference's own 4^3-cell gold configuration has a numerically singular coarse operator).  MFMGB_DENSE_SOLVE=
   * substitution | inverse overrides. */
  MFMGB_API int mfmgb_dense_solve_mode(const mfmgb_dense *D, double *pivot_ratio);

  /* ---- matrix-free Laplace/diffusion operator: fills the CudaMatrixFreeOperator slot
   *      (source/cuda/cuda_matrix_free_operator.cu:32-37,60-70; the operator itself is defined by
   *      tests/laplace_matrix_free.hpp:75-199) on a uniform Cartesian grid, lexicographic DoFs ---- */
  /* coef: host array [n_cells][(degree+1)^dim]; constrained: host uint8[n]; h: cell size per direction */
  MFMGB_API int mfmgb_mf_laplace_create(mfmgb_ctx *ctx, int dim, int degree, const int64_t *cells, const double *h,
                                        const double *coef, const uint8_t *constrained, mfmgb_mf **out);
  /* the same operator on one z-slab of a row-partitioned grid (3D Q1): `cells` is the local box (the owned cell layers
   * plus one layer below when there is a lower neighbour), node planes [own_plane_begin, own_plane_end) of the box are
   * owned, the others are ghost planes; vectors, `constrained` and the Jacobi data use the layout
   * [owned planes | ghost planes below | ghost planes above] -- the [owned | ghost] layout of mfmgb_halo_create */
  MFMGB_API int mfmgb_mf_laplace_create_slab(mfmgb_ctx *ctx, int dim, int degree, const int64_t *cells, const double *h,
                                             const double *coef, const uint8_t *constrained, int64_t own_plane_begin,
                                             int64_t own_plane_end, mfmgb_mf **out);
  MFMGB_API int mfmgb_mf_destroy(mfmgb_ctx *ctx, mfmgb_mf *M);
  MFMGB_API int64_t mfmgb_mf_size(const mfmgb_mf *M);        /* rows = owned nodes */
  MFMGB_API int64_t mfmgb_mf_vector_size(const mfmgb_mf *M); /* owned + ghost nodes */
  /* which kernel serves the operator: 0 = generic colour-phase cell kernel (2D, Q2), 1 = 3D Q1 node-owner z-sweep with a
   * per-cell coefficient, 2 = the same with the per-quadrature-point table, 3 = 3D Q1 with ONE coefficient for the whole
   * grid: factorised 27-point stencil, persistent z-sweep (csrc/mf_q1_sweep.cuh) */
  MFMGB_API int mfmgb_mf_kernel(const mfmgb_mf *M);
  MFMGB_API int mfmgb_mf_apply(mfmgb_ctx *ctx, const mfmgb_mf *M, const double *x, double *y);
  /* diagonal with constrained entries set to 1 (compute_diagonal) into a device vector */
  MFMGB_API int mfmgb_mf_diagonal(mfmgb_ctx *ctx, const mfmgb_mf *M, double *diag_dev);

  /* ---- hierarchy / V-cycle / PCG: replaces mfmg::Hierarchy<V>::vmult/apply
   *      (include/mfmg/common/hierarchy.hpp:238-309) and the dealii::SolverCG loop around it
   *      (tests/hierarchy_driver.cc:200-213) ---- */
  /* n_smoothing_steps = "smoother.n_smoothing_steps" (hierarchy.hpp:169); is_preconditioner = "is preconditioner" (:168) */
  MFMGB_API int mfmgb_hierarchy_create(mfmgb_ctx *ctx, int n_levels, int n_smoothing_steps, int is_preconditioner,
                                       double omega, mfmgb_hierarchy **out);
  /* level operators are BORROWED (they must outlive the hierarchy). level 0 = finest. */
  MFMGB_API int mfmgb_hierarchy_set_operator(mfmgb_hierarchy *H, int level, const mfmgb_csr *A);
  MFMGB_API int mfmgb_hierarchy_set_mf_operator(mfmgb_hierarchy *H, const mfmgb_mf *M); /* level 0 only */
  /* R maps level-1 -> level (n_level x n_{level-1}); P = R^T explicit (may be NULL: built internally) */
  MFMGB_API int mfmgb_hierarchy_set_restrictor(mfmgb_hierarchy *H, int level, const mfmgb_csr *R, const mfmgb_csr *P);
  /* "smoother.type" Chebyshev instead of Jacobi: mfmg::DealIIMatrixFreeSmoother (source/dealii/dealii_matrix_free_smoother.cc:
   * 34-79), i.e. dealii::PreconditionChebyshev with the inverse diagonal as inner preconditioner, applied as
   * x -= p(D^-1 A) D^-1 (A x - b).  Arguments = the reference's "smoother.degree" / "smoother.smoothing_range" /
   * "smoother.max_eigenvalue" and deal.II's eig_cg_n_iterations (reference defaults: 0, 0., 1., 8).  degree counts the
   * operator applications AFTER the first damped-Jacobi step (deal.II @89057dff semantics: degree 0 == Jacobi with
   * omega = 1 / theta, fused into the SpMV like the Jacobi smoother).  Before finalize; single-GPU hierarchies. */
  MFMGB_API int mfmgb_hierarchy_set_smoother_chebyshev(mfmgb_hierarchy *H, int degree, double smoothing_range,
                                                       double max_eigenvalue, int eig_cg_n_iterations);
  /* out4 = lambda_min, lambda_max (incl. deal.II's safety factor 1.2), theta, delta of a level's smoother */
  MFMGB_API int mfmgb_hierarchy_chebyshev_info(const mfmgb_hierarchy *H, int level, double *out4);
  /* builds smoothers, the coarse factorisation and all level workspaces (no allocation afterwards) */
  MFMGB_API int mfmgb_hierarchy_finalize(mfmgb_ctx *ctx, mfmgb_hierarchy *H);
  MFMGB_API int mfmgb_hierarchy_destroy(mfmgb_ctx *ctx, mfmgb_hierarchy *H);
  /* Hierarchy::vmult(x, b): one V-cycle, device vectors */
  MFMGB_API int mfmgb_vcycle(mfmgb_ctx *ctx, mfmgb_hierarchy *H, const double *b, double *x);
  /* Hierarchy::apply(b, x, level) */
  MFMGB_API int mfmgb_hierarchy_apply(mfmgb_ctx *ctx, mfmgb_hierarchy *H, const double *b, double *x, int level);
  /* Hierarchy<Vector<double,Host>>::vmult: HOST vectors, H2D + V-cycle + D2H inside, synchronous
   * (the host-vector specialisations, source/cuda/cuda_matrix_operator.cu:51-70, cuda_smoother.cu:62-84) */
  MFMGB_API int mfmgb_vcycle_host(mfmgb_ctx *ctx, mfmgb_hierarchy *H, const double *b_host, double *x_host);
  /* The same for n_rhs independent right-hand sides (preconditioner mode; block / multi-right-hand-side Krylov drivers):
   * x_host[j] = V-cycle(b_host[j]).  Every right-hand side makes its own host -> device -> host trip; the trips are
   * pipelined over four staging buffers -- H2D of j+1, the cycle of j and D2H of j-1 overlap (PCIe is full duplex) --
   * so the throughput is bound by the slowest of the three stages instead of their sum.  Synchronous: all results are
   * on the host on return.  Pointers may repeat (the same pinned buffer for several j). */
  MFMGB_API int mfmgb_vcycle_host_batch(mfmgb_ctx *ctx, mfmgb_hierarchy *H, int n_rhs, const double *const *b_host,
                                        double *const *x_host);
  /* one un-captured V-cycle with CUDA events between the level-0 stages (measurement aid for bench.py):
   * stage_ms[6] = pre-smoothing, residual, restriction, coarse levels (recursion), prolongation+correction,
   * post-smoothing -- the "Apply: fine levels" / "Apply: coarsest level" timer sections of hierarchy.hpp:263,271. */
  MFMGB_API int mfmgb_vcycle_profile(mfmgb_ctx *ctx, mfmgb_hierarchy *H, const double *b, double *x, double *stage_ms);
  /* finer timeline of one V-cycle: CUDA events between the pieces of the cycle (stages, interior / boundary rows of a
   * partitioned level, the steps of the coarse solve).  use_graph = 1: the events are captured as event-record nodes and
   * the times are those of a CUDA-graph REPLAY (what mfmgb_vcycle runs).  names: ';'-separated piece names, ms[k] =
   * duration of piece k.  (The "Apply: ..." timer sections of hierarchy.hpp:263,271, resolved per kernel.) */
  MFMGB_API int mfmgb_vcycle_timeline(mfmgb_ctx *ctx, mfmgb_hierarchy *H, const double *b, double *x, int use_graph,
                                      char *names, int names_len, double *ms, int max_marks, int *n_marks);
  /* capture the V-cycle into a CUDA graph and replay it on later mfmgb_vcycle calls (1 = on) */
  MFMGB_API int mfmgb_hierarchy_use_graph(mfmgb_hierarchy *H, int on);
  /* kernels launched by one V-cycle of this hierarchy */
  MFMGB_API int mfmgb_hierarchy_launches_per_cycle(const mfmgb_hierarchy *H);

  /* Preconditioned CG with deal.II's SolverCG recurrence; A = level-0 operator of H.  Device vectors.
   * Stops when |A x - b|_2 <= tol (absolute) or after max_it iterations.  res_hist_host (may be NULL)
   * receives max_it + 1 residual norms.  H == NULL: unpreconditioned CG on A.
   * Row-partitioned hierarchy: b and x hold mfmgb_hierarchy_vector_size(H, 0) entries -- the owned rows followed by
   * the ghost tail the halo exchange writes (mfmgb_pcg_host sizes its staging vectors that way itself). */
  MFMGB_API int mfmgb_pcg(mfmgb_ctx *ctx, mfmgb_hierarchy *H, const mfmgb_csr *A, const double *b, double *x, double tol,
                          int max_it, int *iterations, double *res_hist_host);
  MFMGB_API int mfmgb_pcg_host(mfmgb_ctx *ctx, mfmgb_hierarchy *H, const mfmgb_csr *A, const double *b_host,
                               double *x_host, double tol, int max_it, int *iterations, double *res_hist_host);

  /* ---- multi-GPU (one process per GPU, NCCL over NVLink): replaces the per-SpMV host MPI_Allgatherv of the whole
   *      source vector (include/mfmg/cuda/sparse_matrix_device.templates.cuh:104-138, source/cuda/utils.cu:305-482)
   *      by a halo exchange of boundary entries, and deal.II's MPI reductions in SolverCG by ncclAllReduce ---- */
  typedef struct mfmgb_halo mfmgb_halo;
  /* rank 0 creates the 128-byte NCCL id; the launcher (MPI_Bcast / torch.distributed) hands it to every rank */
  MFMGB_API int mfmgb_comm_unique_id(char *out128);
  MFMGB_API int mfmgb_comm_init(mfmgb_ctx *ctx, const char *id128, int nranks, int rank);
  MFMGB_API int mfmgb_comm_finalize(mfmgb_ctx *ctx);
  /* how ghost entries and small reductions travel: NVLink peer-memory stores by the library's own kernels (CUDA IPC
   * window mapped at mfmgb_comm_init; MFMGB_PEER=0 switches it off) or NCCL */
  MFMGB_API const char *mfmgb_comm_transport(mfmgb_ctx *ctx);
  /* MFMGB_ERR_NCCL when a kernel of this context gave up waiting for a peer GPU (MFMGB_PEER_TIMEOUT_MS, default 20 s);
   * call after mfmgb_ctx_synchronize */
  MFMGB_API int mfmgb_comm_check(mfmgb_ctx *ctx);
  MFMGB_API int mfmgb_comm_rank(mfmgb_ctx *ctx);
  MFMGB_API int mfmgb_comm_size(mfmgb_ctx *ctx);
  /* Halo plan of a row-partitioned level.  Vectors that are gathered from have n_owned + n_ghost entries; the ghost
   * tail is ordered by neighbour (neighbor_ranks order, recv_counts entries each).  send_indices: concatenated LOCAL
   * owned indices sent to each neighbour (send_counts each), in the order the receiver stores them.
   * COLLECTIVE when the context has an initialised communicator (the peer-memory mailboxes are allocated symmetrically:
   * every rank creates its plans in the same order). */
  MFMGB_API int mfmgb_halo_create(mfmgb_ctx *ctx, int64_t n_owned, int64_t n_ghost, int n_neighbors,
                                  const int *neighbor_ranks, const int64_t *send_counts, const int32_t *send_indices,
                                  const int64_t *recv_counts, mfmgb_halo **out);
  MFMGB_API int mfmgb_halo_destroy(mfmgb_ctx *ctx, mfmgb_halo *halo);
  /* fill the ghost tail of v (blocking form, on the context stream) */
  MFMGB_API int mfmgb_halo_exchange(mfmgb_ctx *ctx, const mfmgb_halo *halo, double *v);
  MFMGB_API int mfmgb_allreduce_sum(mfmgb_ctx *ctx, double *dev, int n);
  /* attach the plan to a (non-coarsest) level whose operator is n_owned x (n_owned + n_ghost); rows
   * [boundary_lo, boundary_hi) reference owned columns only and overlap the exchange */
  MFMGB_API int mfmgb_hierarchy_set_halo(mfmgb_hierarchy *H, int level, const mfmgb_halo *halo, int64_t boundary_lo,
                                         int64_t boundary_hi);
  /* rows [0, first_boundary_row) of the level's restrictor reference owned fine columns only: they are computed while
   * the residual's halo is exchanged (after mfmgb_hierarchy_set_restrictor; default 0 = no overlap) */
  MFMGB_API int mfmgb_hierarchy_set_restrict_split(mfmgb_hierarchy *H, int level, int64_t first_boundary_row);
  /* Halo-free restriction for hierarchies that use mfmgb_coarse_dd (call after set_restrictor / set_coarse_dd, on EVERY
   * rank): the level's R must hold no ghost columns; R_below (NULL on rank 0) = the restrictor rows of the lower
   * neighbour's separator restricted to this rank's fine entries.  Its product with the residual joins the all-reduce of
   * the coarse solve, so the residual's halo is not exchanged.  R_below is borrowed. */
  MFMGB_API int mfmgb_hierarchy_set_restrict_no_halo(mfmgb_ctx *ctx, mfmgb_hierarchy *H, int level,
                                                     const mfmgb_csr *R_below);
  /* offsets (nranks + 1) of the rank-owned rows of the replicated coarsest level */
  MFMGB_API int mfmgb_hierarchy_set_coarse_offsets(mfmgb_hierarchy *H, const int64_t *offsets, int nranks);
  /* Domain-decomposed form of the dense coarse solve for a row-partitioned hierarchy whose coarse operator is block
   * tridiagonal in the ranks' row blocks (z-slab partitions are): one level of nested dissection with the last
   * agglomerate layer of every rank but the last as separator.  Exactly the reference's direct solve
   * (source/cuda/cuda_solver.cu:496-515) reorganised so that a rank reads 8 n_I^2 bytes per cycle instead of
   * 8 n_c^2 / N, with one all-reduce of n_S doubles as the only communication (csrc/coarse_dd.cu).
   * Index conventions: this rank's interior = coarse rows [own_begin, own_begin + A_II.n_rows); separators are numbered
   * globally 0..n_S-1 (sep_index[k] = coarse row of separator k); A_IS / A_SI couple the interior with the adjacent
   * separators [adj_begin, adj_begin + A_IS.n_cols); the rank's own separator is [own_sep_begin, +own_sep_n).
   * A_SI is BORROWED (must outlive the solver); the others are only read during creation.  Collective call. */
  typedef struct mfmgb_coarse_dd mfmgb_coarse_dd;
  MFMGB_API int mfmgb_coarse_dd_create(mfmgb_ctx *ctx, int64_t n_c, int64_t own_begin, int64_t n_S, int64_t adj_begin,
                                       int64_t own_sep_begin, int64_t own_sep_n, const mfmgb_csr *A_II,
                                       const mfmgb_csr *A_IS, const mfmgb_csr *A_SI, const mfmgb_csr *A_SS,
                                       const int32_t *sep_index, mfmgb_coarse_dd **out);
  MFMGB_API int mfmgb_coarse_dd_destroy(mfmgb_ctx *ctx, mfmgb_coarse_dd *dd);
  /* b_c: valid on this rank's coarse rows; x_c: valid on return on this rank's rows and on all separator rows */
  MFMGB_API int mfmgb_coarse_dd_solve(mfmgb_ctx *ctx, const mfmgb_coarse_dd *dd, const double *b_c, double *x_c);
  /* use it as the coarsest-level solver of a partitioned hierarchy (before finalize; borrowed) */
  MFMGB_API int mfmgb_hierarchy_set_coarse_dd(mfmgb_hierarchy *H, const mfmgb_coarse_dd *dd);
  /* length device vectors of this level must have (n_owned + n_ghost) */
  MFMGB_API int64_t mfmgb_hierarchy_vector_size(const mfmgb_hierarchy *H, int level);
  /* Process-wide launch parameters of the exchange kernels (measurement aid; results do not depend on them):
   * "halo_push_ctas" (16): CTAs of the fused compute + exchange A-kernel that store the boundary plane(s) into the
   * neighbours' mailboxes; "halo_push_penalty" (0): interior tiles such a CTA is spared.  Defaults can also be given
   * as MFMGB_HALO_PUSH_CTAS / MFMGB_HALO_PUSH_PENALTY.  A change applies to launches (and graph captures) made after it.
   * set: MFMGB_ERR_INVALID for an unknown name; get: -1 for an unknown name. */
  MFMGB_API int mfmgb_tunable_set(const char *name, long long value);
  MFMGB_API long long mfmgb_tunable_get(const char *name);

#ifdef __cplusplus
}
#endif
#endif /* MFMG_B200_H */
