// hierarchy.cu -- level storage, the V-cycle (mfmg::Hierarchy::apply) and the PCG driver loop.
//
// V-cycle: include/mfmg/common/hierarchy.hpp:246-309, statement by statement, with
//   * every level vector preallocated at finalize (the reference allocates 4 vectors per level per
//     cycle, hierarchy.hpp:284,289,293,297),
//   * smoother / residual / prolongation-correction fused into SpMV epilogues,
//   * the "x == 0" pre-smoothing sweep reduced to x = omega D^-1 b,
//   * optional CUDA-graph replay of the whole cycle (launch-latency bound on small levels).
// PCG: the recurrence of dealii::SolverCG (tests/hierarchy_driver.cc:200-213), scalars kept on
// the device, one host read-back per iteration for the reference's stopping test.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "comm.cuh"
#include "dense.cuh"
#include "jacobi.cuh"
#include "mf.cuh"
#include "vecops.cuh"

using namespace mfmgb;

struct mfmgb_level
{
  int64_t n = 0;
  const mfmgb_csr *A = nullptr;
  const mfmgb_mf *M = nullptr;
  const mfmgb_csr *R = nullptr; // maps level-1 -> this level
  const mfmgb_csr *P = nullptr; // explicit transpose of R
  mfmgb_csr *P_owned = nullptr;
  mfmgb_jacobi *J = nullptr;
  mfmgb_dense *D = nullptr;
  double *res = nullptr, *xtmp = nullptr; // fine-side work vectors (size n + ghosts)
  double *bc = nullptr, *xc = nullptr;    // this level's rhs / solution when it is the coarse side
  // row-partitioned (multi-GPU) level: ghost entries live in the tail [n, n + n_ghost) of every gathered vector
  const mfmgb_halo *halo = nullptr;
  int64_t blo = 0, bhi = 0; // rows [0, blo) and [bhi, n) reference ghost columns, [blo, bhi) is the interior
  int64_t r_split = 0;      // (as coarse side) rows [0, r_split) of R reference owned fine columns only
  // halo-free restriction (with the domain-decomposed coarse solve): R holds no ghost columns; R_below = the rows of
  // the lower neighbour's separator restricted to this rank's entries (NULL on rank 0), gb = R_below r
  bool restrict_no_halo = false;
  const mfmgb_csr *R_below = nullptr;
  double *gb = nullptr;
  // [R ; R_below] as ONE matrix (built at setup) writing [own coarse rows | separator-below share] into rs: one launch
  // instead of two on every rank but the first (those ranks are the critical path into the coarse all-reduce)
  mfmgb_csr *R_stack = nullptr;
  double *rs = nullptr;
  // Chebyshev smoother (dealii::PreconditionChebyshev as DealIIMatrixFreeSmoother uses it): estimate and work vectors
  double cheb_lmin = 1., cheb_lmax = 1., cheb_theta = 1., cheb_delta = 0.;
  double *c_r = nullptr, *c_dst = nullptr, *c_u1 = nullptr, *c_u2 = nullptr;
};

struct mfmgb_hierarchy
{
  int n_levels = 0;
  int nu = 1;
  bool is_preconditioner = true;
  double omega = 1.;
  bool finalized = false;
  // smoother.type: Jacobi (source/cuda/cuda_smoother.cu) or Chebyshev (source/dealii/dealii_matrix_free_smoother.cc:34-60)
  bool chebyshev = false;
  int cheb_degree = 0;          // smoother.degree: matrix-vector products after the first damped-Jacobi step
  double cheb_range = 0.;       // smoother.smoothing_range
  double cheb_max_ev = 1.;      // smoother.max_eigenvalue (used when cheb_cg_its == 0)
  int cheb_cg_its = 8;          // eig_cg_n_iterations
  std::vector<mfmgb_level> lev;
  // graph replay
  bool use_graph = false;
  // instantiated cycles, keyed by the (b, x) pair they were captured with (PCG, the host-vector entry point and the
  // pipelined batch entry point each bring their own buffers); least recently used slot is replaced
  struct GraphSlot
  {
    const double *b = nullptr;
    double *x = nullptr;
    cudaGraphExec_t exec = nullptr;
    uint64_t last_use = 0;
  };
  GraphSlot graphs[6];
  uint64_t graph_clock = 0;
  // pipelined host-vector batches (mfmgb_vcycle_host_batch): staging buffers, copy streams, events
  static constexpr int kBatchBuffers = 4; // one more than the three stages: slack for jitter between the copies and the cycle
  double *bb_dev[kBatchBuffers] = {}, *bx_dev[kBatchBuffers] = {};
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_h2d[kBatchBuffers] = {}, ev_comp[kBatchBuffers] = {}, ev_d2h[kBatchBuffers] = {};
  int launches_per_cycle = 0;
  // device staging for the *_host entry points and PCG work vectors
  double *b_dev = nullptr, *x_dev = nullptr;
  double *g = nullptr, *h = nullptr, *d = nullptr;
  double *scal = nullptr;      // device scalars: [0]=gh_old [1]=gh_new [2]=dh [3]=res2
  double *scal_host = nullptr; // pinned
  // multi-GPU: offsets of the rank-owned slices of the (replicated) coarsest-level vectors
  std::vector<int64_t> coarse_offsets;
  bool distributed = false;
  const mfmgb_coarse_dd *dd = nullptr; // domain-decomposed coarse solve (coarse_dd.cu) instead of the dense inverse
  // stage profiling (mfmgb_vcycle_profile): events between the level-0 stages
  bool profiling = false;
  cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

namespace
{
constexpr int kBlock = 256;

void drop_graph(void *h)
{
  mfmgb_hierarchy *H = static_cast<mfmgb_hierarchy *>(h);
  for (auto &g : H->graphs)
    if (g.exec)
    {
      cudaGraphExecDestroy(g.exec);
      g = mfmgb_hierarchy::GraphSlot();
    }
}

int level_apply_A(mfmgb_ctx *ctx, const mfmgb_level &l, double *x, Epi epi, const EpiArgs &e)
{
  if (l.M)
  {
    if (l.halo)
    {
      // the middle z chunks read owned planes only: they run while the ghost planes are exchanged
      const int nc = mf_num_chunks(l.M);
      MFMGB_CHECK(halo_start(ctx, l.halo, x));
      if (nc >= 3)
        MFMGB_CHECK(mf_apply_chunks(ctx, l.M, x, epi, e, 1, nc - 1));
      MFMGB_CHECK(halo_wait(ctx, l.halo, x));
      if (nc < 3)
        return mf_apply(ctx, l.M, x, epi, e);
      MFMGB_CHECK(mf_apply_chunks(ctx, l.M, x, epi, e, 0, 1));
      return mf_apply_chunks(ctx, l.M, x, epi, e, nc - 1, nc);
    }
    return mf_apply(ctx, l.M, x, epi, e);
  }
  if (!l.halo)
    return csr_apply(ctx, l.A, x, epi, e);
  // row-partitioned level: the halo exchange of x runs on the communication stream while the interior rows
  // are computed; the rows that reference ghost columns follow once the ghosts have landed
  // (MFMGB_HALO_OVERLAP=0: exchange first, then one launch over all rows -- measurement aid)
  static const bool overlap = [] {
    const char *v = getenv("MFMGB_HALO_OVERLAP");
    return !(v && v[0] == '0');
  }();
  if (halo_can_fuse(ctx, l.halo) && csr_can_fuse_ghost(l.A))
  {
    // ONE launch per application, like on one GPU -- the fused compute + exchange kernel: its first CTAs store this
    // rank's boundary entries into the neighbours' mailboxes over NVLink and raise their flags, all CTAs compute rows,
    // those that reach a boundary tile wait for the neighbours' flags inside the kernel and gather the ghost columns
    // straight from the mailbox
    GhostArgs g;
    halo_ghost_args(ctx, l.halo, l.blo, l.bhi, &g);
    if (g.n_push_ctas == 0)
    {
      MFMGB_CHECK(halo_push_inline(ctx, l.halo, x));
      prof_mark(ctx, "A push boundary plane(s) to the neighbours (NVLink)");
    }
    MFMGB_CHECK(csr_apply_tile(ctx, l.A, x, epi, e, 0, l.n, 0, 0, &g));
    prof_mark(ctx, "A all rows: push + in-kernel wait + ghost columns from the mailbox (one launch)");
    return MFMGB_OK;
  }
  MFMGB_CHECK(halo_start(ctx, l.halo, x));
  if (!overlap)
  {
    MFMGB_CHECK(halo_wait(ctx, l.halo, x));
    return csr_apply(ctx, l.A, x, epi, e);
  }
  MFMGB_CHECK(csr_apply(ctx, l.A, x, epi, e, l.blo, l.bhi));
  prof_mark(ctx, "A interior rows");
  MFMGB_CHECK(halo_wait(ctx, l.halo, x));
  prof_mark(ctx, "A halo wait");
  MFMGB_CHECK(csr_apply2(ctx, l.A, x, epi, e, 0, l.blo, l.bhi, l.n)); // both boundary blocks in one launch
  prof_mark(ctx, "A boundary rows");
  return MFMGB_OK;
}

static const char *const kStageNames[7] = {"(start)", "pre-smooth", "residual", "restrict", "coarse solve",
                                           "prolong+correct", "post-smooth"};
#define STAGE_MARK(k)                                                                              \
  do                                                                                                \
  {                                                                                                 \
    if (H->profiling && li == 0)                                                                    \
      MFMGB_CUDA(ctx, cudaEventRecord(H->ev[k], ctx->stream));                                      \
    if (li == 0)                                                                                    \
      prof_mark(ctx, kStageNames[k]);                                                               \
  } while (0)

// ---- Chebyshev smoother kernels (elementwise; the operator applications go through level_apply_A) ----------------
__global__ void __launch_bounds__(kBlock) negate_into_kernel(int64_t n, const double *__restrict__ a, double *__restrict__ out)
{
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    out[i] = -a[i];
}
// vector_updates(start_zero): dst = D^-1 (f2 src), u1 = -dst
__global__ void __launch_bounds__(kBlock)
    cheb_start_kernel(int64_t n, double f2, const double *__restrict__ dinv, const double *__restrict__ src,
                      double *__restrict__ dst, double *__restrict__ u1)
{
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
  {
    const double d = __dmul_rn(dinv[i], __dmul_rn(f2, src[i]));
    dst[i] = d;
    u1[i] = -d;
  }
}
// vector_updates(!start_zero) with u2 = A dst - src already formed: u1 = f1 u1 + f2 D^-1 u2, dst -= u1
__global__ void __launch_bounds__(kBlock)
    cheb_step_kernel(int64_t n, double f1, double f2, const double *__restrict__ dinv, const double *__restrict__ u2,
                     double *__restrict__ u1, double *__restrict__ dst)
{
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
  {
    const double t = __dmul_rn(dinv[i], u2[i]);
    const double u = __dadd_rn(__dmul_rn(f1, u1[i]), __dmul_rn(f2, t));
    u1[i] = u;
    dst[i] = __dsub_rn(dst[i], u);
  }
}
// x = (zero ? 0 : x) - dst      (DealIIMatrixFreeSmoother::apply: x.add(-1., tmp))
__global__ void __launch_bounds__(kBlock)
    cheb_finish_kernel(int64_t n, int zero, const double *__restrict__ dst, double *__restrict__ x)
{
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    x[i] = __dsub_rn(zero ? 0. : x[i], dst[i]);
}
__global__ void __launch_bounds__(kBlock)
    scale_rows_kernel(int64_t n, const double *__restrict__ dinv, const double *__restrict__ g, double *__restrict__ h)
{
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    h[i] = dinv[i] * g[i];
}
__global__ void __launch_bounds__(kBlock)
    cg_direction_host_beta_kernel(int64_t n, double beta, const double *__restrict__ h, double *__restrict__ d)
{
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    d[i] = fma(beta, d[i], -h[i]);
}
__global__ void __launch_bounds__(kBlock) mod11_guess_kernel(int64_t n, double mean, double *__restrict__ g)
{
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    g[i] = -((double)(i % 11) - mean); // g = A 0 - v
}

inline unsigned ew_grid(const mfmgb_ctx *ctx, int64_t n)
{
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, kBlock), (int64_t)ctx->num_sms * 16));
}

// ascending eigenvalues of the symmetric tridiagonal matrix (d, e), k <= 64: cyclic Jacobi rotations (setup, host)
void tridiag_eigenvalues(int k, const double *d, const double *e, double *w)
{
  std::vector<double> a((size_t)k * (size_t)k, 0.);
  for (int i = 0; i < k; ++i)
  {
    a[(size_t)i * k + i] = d[i];
    if (i + 1 < k)
      a[(size_t)i * k + i + 1] = a[(size_t)(i + 1) * k + i] = e[i];
  }
  for (int sweep = 0; sweep < 100; ++sweep)
  {
    double off = 0.;
    for (int i = 0; i < k; ++i)
      for (int j = i + 1; j < k; ++j)
        off += a[(size_t)i * k + j] * a[(size_t)i * k + j];
    if (off < 1e-300)
      break;
    for (int p = 0; p < k; ++p)
      for (int q = p + 1; q < k; ++q)
      {
        const double apq = a[(size_t)p * k + q];
        if (apq == 0.)
          continue;
        const double tau = (a[(size_t)q * k + q] - a[(size_t)p * k + p]) / (2. * apq);
        const double t = (tau >= 0. ? 1. : -1.) / (std::fabs(tau) + std::sqrt(1. + tau * tau));
        const double c = 1. / std::sqrt(1. + t * t), sn = t * c;
        for (int r = 0; r < k; ++r)
        {
          const double arp = a[(size_t)r * k + p], arq = a[(size_t)r * k + q];
          a[(size_t)r * k + p] = c * arp - sn * arq;
          a[(size_t)r * k + q] = sn * arp + c * arq;
        }
        for (int r = 0; r < k; ++r)
        {
          const double apr = a[(size_t)p * k + r], aqr = a[(size_t)q * k + r];
          a[(size_t)p * k + r] = c * apr - sn * aqr;
          a[(size_t)q * k + r] = sn * apr + c * aqr;
        }
      }
  }
  for (int i = 0; i < k; ++i)
    w[i] = a[(size_t)i * k + i];
  std::sort(w, w + k);
}

// dealii::PreconditionChebyshev::estimate_eigenvalues (deal.II @89057dff, pinned by mfmg's ci/Dockerfile; restated from
// its published algorithm, see oracle/mfmg_oracle.c cheb_estimate): eig_cg_n_iterations steps of D^-1-preconditioned CG
// on A x = v from x = 0 give the Lanczos tridiagonal matrix whose extreme eigenvalues estimate those of D^-1 A.
// Setup-time: the operator applications and dots run on the device, the scalar recurrence on the host.
int cheb_estimate(mfmgb_ctx *ctx, mfmgb_hierarchy *H, mfmgb_level &l)
{
  const int64_t n = l.n;
  double lmin = 1., lmax = 1.;
  if (H->cheb_cg_its > 0 && n > 0)
  {
    double *g = l.c_r, *hh = l.c_u1, *d = l.c_dst, *Ad = l.c_u2;
    cudaStream_t st = ctx->stream;
    const unsigned grid = ew_grid(ctx, n);
    auto dot = [&](const double *a, const double *b, double *out) {
      MFMGB_CHECK(mfmgb_vec_dot(ctx, a, b, n, out));
      return (int)MFMGB_OK;
    };
    const double mean = [&] { // mean of (i mod 11), i < n
      const int64_t full = n / 11, rem = n % 11;
      return ((double)full * 55. + (double)(rem * (rem - 1) / 2)) / (double)n;
    }();
    mod11_guess_kernel<<<grid, kBlock, 0, st>>>(n, mean, g);
    MFMGB_LAUNCHED(ctx);
    double gg = 0.;
    MFMGB_CHECK(dot(g, g, &gg));
    double res = std::sqrt(gg);
    const double res0 = res, tol = std::sqrt(2.220446049250313e-16), reduce = 1e-10;
    double diag[64], offd[64];
    int k = 0;
    if (res > tol)
    {
      scale_rows_kernel<<<grid, kBlock, 0, st>>>(n, l.J->dinv, g, hh);
      MFMGB_LAUNCHED(ctx);
      negate_into_kernel<<<grid, kBlock, 0, st>>>(n, hh, d);
      MFMGB_LAUNCHED(ctx);
      double gh = 0., beta_alpha = 0.;
      MFMGB_CHECK(dot(g, hh, &gh));
      const int max_it = std::min(H->cheb_cg_its, 64);
      for (int it = 1; it <= max_it; ++it)
      {
        EpiArgs e;
        e.y = Ad;
        MFMGB_CHECK(level_apply_A(ctx, l, d, Epi::Spmv, e));
        double dAd = 0.;
        MFMGB_CHECK(dot(d, Ad, &dAd));
        const double alpha = gh / dAd;
        MFMGB_CHECK(vec_axpy(ctx, g, alpha, Ad, n));
        MFMGB_CHECK(dot(g, g, &gg));
        res = std::sqrt(gg);
        scale_rows_kernel<<<grid, kBlock, 0, st>>>(n, l.J->dinv, g, hh);
        MFMGB_LAUNCHED(ctx);
        double beta = gh;
        MFMGB_CHECK(dot(g, hh, &gh));
        beta = gh / beta;
        diag[k] = 1. / alpha + beta_alpha;
        beta_alpha = beta / alpha;
        offd[k] = std::sqrt(beta) / alpha;
        ++k;
        if (res <= tol || res <= reduce * res0)
          break;
        cg_direction_host_beta_kernel<<<grid, kBlock, 0, st>>>(n, beta, hh, d);
        MFMGB_LAUNCHED(ctx);
      }
    }
    if (k > 0)
    {
      double w[64];
      tridiag_eigenvalues(k, diag, offd, w);
      lmin = w[0];
      lmax = 1.2 * w[k - 1]; // safety factor: the CG is not converged in general
    }
  }
  else
  {
    lmax = H->cheb_max_ev;
    lmin = H->cheb_range != 0. ? H->cheb_max_ev / H->cheb_range : H->cheb_max_ev;
  }
  const double alpha = H->cheb_range > 1. ? lmax / H->cheb_range : std::min(0.9 * lmax, lmin);
  l.cheb_lmin = lmin;
  l.cheb_lmax = lmax;
  l.cheb_delta = (lmax - alpha) * 0.5;
  l.cheb_theta = (lmax + alpha) * 0.5;
  return MFMGB_OK;
}

// One Chebyshev smoothing step in place (degree >= 1), DealIIMatrixFreeSmoother::apply
// (source/dealii/dealii_matrix_free_smoother.cc:67-79): r = A x - b; tmp = p(D^-1 A) D^-1 r; x -= tmp.
// zero: x == 0 on entry (r = -b without an operator application).
int cheb_sweep(mfmgb_ctx *ctx, mfmgb_hierarchy *H, mfmgb_level &l, const double *b, double *x, bool zero)
{
  const int64_t n = l.n;
  cudaStream_t st = ctx->stream;
  const unsigned grid = ew_grid(ctx, n);
  if (zero)
  {
    negate_into_kernel<<<grid, kBlock, 0, st>>>(n, b, l.c_r);
    MFMGB_LAUNCHED(ctx);
  }
  else
  {
    EpiArgs e;
    e.y = l.c_r;
    e.b = b;
    MFMGB_CHECK(level_apply_A(ctx, l, x, Epi::Resid, e));
  }
  const double theta = l.cheb_theta, delta = l.cheb_delta;
  cheb_start_kernel<<<grid, kBlock, 0, st>>>(n, 1. / theta, l.J->dinv, l.c_r, l.c_dst, l.c_u1);
  MFMGB_LAUNCHED(ctx);
  if (std::fabs(delta) >= 1e-40)
  {
    double rhok = delta / theta;
    const double sigma = theta / delta;
    for (int k = 0; k < H->cheb_degree; ++k)
    {
      EpiArgs e;
      e.y = l.c_u2;
      e.b = l.c_r;
      MFMGB_CHECK(level_apply_A(ctx, l, l.c_dst, Epi::Resid, e)); // u2 = A dst - src
      const double rhokp = 1. / (2. * sigma - rhok);
      const double f1 = rhokp * rhok, f2 = 2. * rhokp / delta;
      rhok = rhokp;
      cheb_step_kernel<<<grid, kBlock, 0, st>>>(n, f1, f2, l.J->dinv, l.c_u2, l.c_u1, l.c_dst);
      MFMGB_LAUNCHED(ctx);
    }
  }
  cheb_finish_kernel<<<grid, kBlock, 0, st>>>(n, zero ? 1 : 0, l.c_dst, x);
  MFMGB_LAUNCHED(ctx);
  return MFMGB_OK;
}

int apply_level(mfmgb_ctx *ctx, mfmgb_hierarchy *H, const double *b, double *x, int li)
{
  mfmgb_level &fine = H->lev[li];
  const int64_t n = fine.n;
  const bool x_is_zero = li > 0 || H->is_preconditioner; // hierarchy.hpp:253-259
  if (li == H->n_levels - 1) // hierarchy.hpp:261-268 (the solve overwrites x)
  {
    if (!H->dd)
      return dense_solve_async(ctx, fine.D, b, x);
    if (fine.R_stack && b == fine.bc)
    {
      // the stacked restriction wrote [own rows | share below] into rs: the solver addresses its right-hand side by
      // GLOBAL coarse index and only touches this rank's rows, so hand it rs shifted back by the rank's offset
      mfmgb_comm *c = ctx_comm(ctx);
      return coarse_dd_solve_async(ctx, H->dd, fine.rs - H->coarse_offsets[c->rank], x, fine.rs + fine.R->n_rows);
    }
    return coarse_dd_solve_async(ctx, H->dd, b, x, fine.restrict_no_halo ? fine.gb : nullptr);
  }

  mfmgb_level &coarse = H->lev[li + 1];
  const int nu = H->nu;
  // Chebyshev of degree >= 1: every sweep works in place on x (1 + degree operator applications, own work vectors).
  // Degree 0 is damped Jacobi with omega = 1 / theta (set at finalize) and takes the fused path below.
  const bool cheb_in_place = H->chebyshev && H->cheb_degree > 0;
  // sweeps that run as an out-of-place fused SpMV (each flips the buffer the iterate lives in)
  const int n_oop = cheb_in_place ? 0 : (x_is_zero && nu > 0 ? nu - 1 : nu) + nu;
  double *cur = x, *other = fine.xtmp;
  const double omega = fine.J->omega;
  STAGE_MARK(0);
  if (cheb_in_place)
  {
    if (x_is_zero && nu == 0)
      MFMGB_CHECK(vec_fill(ctx, cur, 0., n));
    for (int s = 0; s < nu; ++s)
      MFMGB_CHECK(cheb_sweep(ctx, H, fine, b, cur, x_is_zero && s == 0));
  }
  else if (x_is_zero)
  {
    if (n_oop & 1)
      std::swap(cur, other); // start in xtmp so the last sweep lands in x
    if (nu > 0)
      MFMGB_CHECK(mfmgb_jacobi_apply_zero_guess(ctx, fine.J, b, cur)); // x = omega D^-1 b
    else
      MFMGB_CHECK(vec_fill(ctx, cur, 0., n));
  }
  EpiArgs e;
  for (int s = (x_is_zero && nu > 0 ? 1 : 0); s < nu && !cheb_in_place; ++s) // pre-smoothing, hierarchy.hpp:277-279
  {
    e = EpiArgs();
    e.y = other;
    e.b = b;
    e.dinv = fine.J->dinv;
    e.xin = cur;
    e.omega = omega;
    MFMGB_CHECK(level_apply_A(ctx, fine, cur, Epi::Jacobi, e));
    std::swap(cur, other);
  }
  STAGE_MARK(1);
  // negative residual res = A x - b, hierarchy.hpp:282-286
  e = EpiArgs();
  e.y = fine.res;
  e.b = b;
  MFMGB_CHECK(level_apply_A(ctx, fine, cur, Epi::Resid, e));
  STAGE_MARK(2);
  // b_c = R res, hierarchy.hpp:289-290
  e = EpiArgs();
  if (fine.halo && coarse.restrict_no_halo)
  {
    // no exchange at all: the ghost-plane entries of R were dropped at setup; the neighbour above computes them
    // (its R_below) and they are summed by the all-reduce of the domain-decomposed coarse solve
    mfmgb_comm *c = ctx_comm(ctx);
    if (coarse.R_stack)
    {
      e.y = coarse.rs; // [own coarse rows | this rank's share of the separator rows below]
      MFMGB_CHECK(csr_apply(ctx, coarse.R_stack, fine.res, Epi::Spmv, e));
    }
    else
    {
      e.y = coarse.bc + H->coarse_offsets[c->rank];
      MFMGB_CHECK(csr_apply(ctx, coarse.R, fine.res, Epi::Spmv, e));
      if (coarse.R_below)
      {
        e.y = coarse.gb;
        MFMGB_CHECK(csr_apply(ctx, coarse.R_below, fine.res, Epi::Spmv, e));
      }
    }
  }
  else if (fine.halo)
  {
    // R's columns include the ghost plane of the residual: exchange it (rows [0, r_split) of R that only read owned
    // entries may overlap the exchange); owned coarse rows go to this rank's slice of b_c, then -- for the dense
    // coarse solve -- every rank gathers the whole (small) coarse right-hand side
    MFMGB_CHECK(halo_start(ctx, fine.halo, fine.res));
    mfmgb_comm *c = ctx_comm(ctx);
    e.y = coarse.bc + H->coarse_offsets[c->rank];
    if (coarse.r_split > 0)
      MFMGB_CHECK(csr_apply(ctx, coarse.R, fine.res, Epi::Spmv, e, 0, coarse.r_split));
    MFMGB_CHECK(halo_wait(ctx, fine.halo, fine.res));
    if (coarse.r_split < coarse.R->n_rows)
      MFMGB_CHECK(csr_apply(ctx, coarse.R, fine.res, Epi::Spmv, e, coarse.r_split, coarse.R->n_rows));
    if (!H->dd) // (the domain-decomposed coarse solve reads only this rank's slice)
      MFMGB_CHECK(allgather_slices(ctx, coarse.bc, H->coarse_offsets));
  }
  else
  {
    e.y = coarse.bc;
    MFMGB_CHECK(csr_apply(ctx, coarse.R, fine.res, Epi::Spmv, e));
  }
  STAGE_MARK(3);
  // recurse, hierarchy.hpp:293-294
  MFMGB_CHECK(apply_level(ctx, H, coarse.bc, coarse.xc, li + 1));
  STAGE_MARK(4);
  // x -= R^T x_c, hierarchy.hpp:297-302 (explicit transpose, fused subtraction, in place: row-local)
  e = EpiArgs();
  e.y = cur;
  MFMGB_CHECK(csr_apply(ctx, coarse.P, coarse.xc, Epi::Sub, e));
  STAGE_MARK(5);
  for (int s = 0; s < nu; ++s) // post-smoothing, hierarchy.hpp:305-306
  {
    if (cheb_in_place)
    {
      MFMGB_CHECK(cheb_sweep(ctx, H, fine, b, cur, false));
      continue;
    }
    e = EpiArgs();
    e.y = other;
    e.b = b;
    e.dinv = fine.J->dinv;
    e.xin = cur;
    e.omega = omega;
    MFMGB_CHECK(level_apply_A(ctx, fine, cur, Epi::Jacobi, e));
    std::swap(cur, other);
  }
  if (cur != x) // only in solver mode with an odd number of sweeps
    MFMGB_CUDA(ctx, cudaMemcpyAsync(x, cur, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
  STAGE_MARK(6);
  return MFMGB_OK;
}

int run_vcycle(mfmgb_ctx *ctx, mfmgb_hierarchy *H, const double *b, double *x)
{
  // The partitioned cycle is captured too: NCCL operations and the fork/join onto the communication stream (events)
  // are capturable, every rank replays the same sequence, and ~25 eager launches per cycle are most of the
  // multi-GPU overhead at 2 M DoFs per GPU.  (The graph must be dropped before ncclCommDestroy: ctx->graph_owners.)
  // MFMGB_DIST_GRAPH=0 falls back to eager launches.
  static const bool dist_graph = [] {
    const char *v = getenv("MFMGB_DIST_GRAPH");
    return !(v && v[0] == '0');
  }();
  if (!H->use_graph || (H->distributed && !dist_graph))
    return apply_level(ctx, H, b, x, 0);
  mfmgb_hierarchy::GraphSlot *slot = nullptr, *victim = &H->graphs[0];
  for (auto &g : H->graphs)
  {
    if (g.exec && g.b == b && g.x == x)
      slot = &g;
    if (!g.exec || (victim->exec && g.last_use < victim->last_use))
      victim = &g;
  }
  if (!slot)
  {
    slot = victim;
    if (slot->exec)
    {
      cudaGraphExecDestroy(slot->exec);
      slot->exec = nullptr;
    }
    cudaGraph_t graph = nullptr;
    const int64_t before = ctx->launches;
    MFMGB_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = apply_level(ctx, H, b, x, 0);
    cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
    if (rc != MFMGB_OK)
    {
      if (graph)
        cudaGraphDestroy(graph);
      return rc;
    }
    MFMGB_CUDA(ctx, ce);
    H->launches_per_cycle = (int)(ctx->launches - before);
    ctx->launches = before;
    MFMGB_CUDA(ctx, cudaGraphInstantiate(&slot->exec, graph, 0));
    cudaGraphDestroy(graph);
    slot->b = b;
    slot->x = x;
  }
  slot->last_use = ++H->graph_clock;
  MFMGB_CUDA(ctx, cudaGraphLaunch(slot->exec, ctx->stream));
  ctx->launches += H->launches_per_cycle;
  return MFMGB_OK;
}

// ---- PCG kernels (device-resident scalars) ---------------------------------------------------
// g += alpha h; x += alpha d with alpha = gh / dh; partial sums of g.g
__global__ void __launch_bounds__(kBlock)
    pcg_update_kernel(int64_t n, const double *__restrict__ scal, const double *__restrict__ h,
                      const double *__restrict__ d, double *__restrict__ g, double *__restrict__ x,
                      double *__restrict__ partials)
{
  __shared__ double sm[kBlock / 32];
  const double alpha = scal[0] / scal[2];
  double s = 0.;
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
  {
    const double gi = fma(alpha, h[i], g[i]);
    g[i] = gi;
    x[i] = fma(alpha, d[i], x[i]);
    s = fma(gi, gi, s);
  }
  s = block_sum<kBlock>(s, sm);
  if (threadIdx.x == 0)
    partials[blockIdx.x] = s;
}

// d = beta d - h with beta = gh_new / gh_old
__global__ void __launch_bounds__(kBlock)
    pcg_direction_kernel(int64_t n, const double *__restrict__ scal, const double *__restrict__ h, double *__restrict__ d)
{
  const double beta = scal[1] / scal[0];
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    d[i] = fma(beta, d[i], -h[i]);
}

__global__ void negate_kernel(int64_t n, const double *__restrict__ h, double *__restrict__ d)
{
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    d[i] = -h[i];
}

__global__ void rotate_scalar_kernel(double *scal) { scal[0] = scal[1]; }

int ensure_pcg_work(mfmgb_ctx *ctx, mfmgb_hierarchy *H, int64_t n)
{
  if (H->g)
    return MFMGB_OK;
  if (!H->lev.empty() && H->lev[0].halo)
    n += H->lev[0].halo->n_ghost; // gathered vectors carry their ghost tail
  MFMGB_CUDA(ctx, cudaMalloc(&H->g, sizeof(double) * (size_t)(n + 2)));
  MFMGB_CUDA(ctx, cudaMalloc(&H->h, sizeof(double) * (size_t)(n + 2)));
  MFMGB_CUDA(ctx, cudaMalloc(&H->d, sizeof(double) * (size_t)(n + 2)));
  MFMGB_CUDA(ctx, cudaMalloc(&H->scal, sizeof(double) * 8));
  MFMGB_CUDA(ctx, cudaMallocHost(&H->scal_host, sizeof(double) * 8));
  return MFMGB_OK;
}

int ensure_host_staging(mfmgb_ctx *ctx, mfmgb_hierarchy *H, int64_t n)
{
  if (H->b_dev)
    return MFMGB_OK;
  if (H->lev[0].halo)
    n += H->lev[0].halo->n_ghost;
  MFMGB_CUDA(ctx, cudaMalloc(&H->b_dev, sizeof(double) * (size_t)(n + 2)));
  MFMGB_CUDA(ctx, cudaMalloc(&H->x_dev, sizeof(double) * (size_t)(n + 2)));
  return MFMGB_OK;
}
} // namespace

extern "C"
{
  MFMGB_API int mfmgb_hierarchy_create(mfmgb_ctx *ctx, int n_levels, int n_smoothing_steps, int is_preconditioner,
                                       double omega, mfmgb_hierarchy **out)
  {
    MFMGB_REQUIRE(ctx, ctx && out && n_levels >= 1 && n_smoothing_steps >= 0, "mfmgb_hierarchy_create: bad arguments");
    mfmgb_hierarchy *H = new mfmgb_hierarchy();
    H->n_levels = n_levels;
    H->nu = n_smoothing_steps;
    H->is_preconditioner = is_preconditioner != 0;
    H->omega = omega;
    H->lev.resize((size_t)n_levels);
    ctx->graph_owners.emplace_back(H, &drop_graph);
    *out = H;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_hierarchy_set_operator(mfmgb_hierarchy *H, int level, const mfmgb_csr *A)
  {
    if (!H || !A || level < 0 || level >= H->n_levels || H->finalized)
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_hierarchy_set_operator: bad arguments");
    if (A->n_cols < A->n_rows)
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_hierarchy_set_operator: level operator must be square "
                                                "(or n_rows x (n_rows + ghosts) for a row-partitioned level)");
    H->lev[level].A = A;
    H->lev[level].n = A->n_rows;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_hierarchy_set_mf_operator(mfmgb_hierarchy *H, const mfmgb_mf *M)
  {
    if (!H || !M || H->finalized)
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_hierarchy_set_mf_operator: bad arguments");
    H->lev[0].M = M;
    H->lev[0].n = M->n;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_hierarchy_set_restrictor(mfmgb_hierarchy *H, int level, const mfmgb_csr *R, const mfmgb_csr *P)
  {
    if (!H || !R || level < 1 || level >= H->n_levels || H->finalized)
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_hierarchy_set_restrictor: bad arguments");
    if (P && !H->distributed && (P->n_rows != R->n_cols || P->n_cols != R->n_rows || P->nnz != R->nnz))
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_hierarchy_set_restrictor: P is not shaped like R^T");
    H->lev[level].R = R;
    H->lev[level].P = P;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_hierarchy_set_halo(mfmgb_hierarchy *H, int level, const mfmgb_halo *halo, int64_t boundary_lo,
                                         int64_t boundary_hi)
  {
    if (!H || !halo || level < 0 || level >= H->n_levels - 1 || H->finalized)
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_hierarchy_set_halo: bad arguments");
    mfmgb_level &l = H->lev[level];
    const int64_t n_vec = l.A ? l.A->n_cols : (l.M ? l.M->n_local : -1);
    if (halo->n_owned != l.n || n_vec != halo->n_owned + halo->n_ghost || boundary_lo < 0 || boundary_hi > l.n ||
        boundary_lo > boundary_hi)
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_hierarchy_set_halo: plan does not match the level operator");
    l.halo = halo;
    l.blo = boundary_lo;
    l.bhi = boundary_hi;
    H->distributed = true;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_hierarchy_set_restrict_split(mfmgb_hierarchy *H, int level, int64_t first_boundary_row)
  {
    if (!H || level < 1 || level >= H->n_levels || H->finalized || !H->lev[level].R || first_boundary_row < 0 ||
        first_boundary_row > H->lev[level].R->n_rows)
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_hierarchy_set_restrict_split: bad arguments");
    H->lev[level].r_split = first_boundary_row;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_hierarchy_set_restrict_no_halo(mfmgb_ctx *ctx, mfmgb_hierarchy *H, int level,
                                                     const mfmgb_csr *R_below)
  {
    MFMGB_REQUIRE(ctx, ctx && H && level >= 1 && level < H->n_levels && !H->finalized && H->lev[level].R && H->dd,
                  "mfmgb_hierarchy_set_restrict_no_halo: needs a restrictor and the domain-decomposed coarse solve");
    mfmgb_level &l = H->lev[level];
    const int64_t n_below = coarse_dd_n_sep_below(H->dd);
    if (R_below ? (R_below->n_rows != n_below || R_below->n_cols != l.R->n_cols) : n_below != 0)
      return fail(ctx, MFMGB_ERR_INVALID, "mfmgb_hierarchy_set_restrict_no_halo: R_below must be %lld x %lld",
                  (long long)n_below, (long long)l.R->n_cols);
    l.restrict_no_halo = true;
    l.R_below = R_below;
    if (R_below)
    {
      MFMGB_CHECK(mfmgb_vec_alloc(ctx, n_below, &l.gb));
      const char *st = getenv("MFMGB_RESTRICT_STACK");
      if (!(st && st[0] == '0'))
      {
        // stack the two restrictors (setup: download, concatenate, upload)
        const mfmgb_csr *parts[2] = {l.R, R_below};
        std::vector<int64_t> rp(1, 0);
        std::vector<int32_t> col;
        std::vector<double> val;
        for (const mfmgb_csr *M : parts)
        {
          std::vector<int64_t> prp((size_t)M->n_rows + 1);
          std::vector<int32_t> pc((size_t)std::max<int64_t>(M->nnz, 1));
          std::vector<double> pv((size_t)std::max<int64_t>(M->nnz, 1));
          MFMGB_CHECK(mfmgb_csr_download(ctx, M, prp.data(), pc.data(), pv.data()));
          const int64_t base = rp.back();
          for (int64_t i = 1; i <= M->n_rows; ++i)
            rp.push_back(base + prp[(size_t)i]);
          col.insert(col.end(), pc.begin(), pc.begin() + M->nnz);
          val.insert(val.end(), pv.begin(), pv.begin() + M->nnz);
        }
        if (col.empty())
        {
          col.push_back(0);
          val.push_back(0.);
        }
        MFMGB_CHECK(mfmgb_csr_upload(ctx, l.R->n_rows + n_below, l.R->n_cols, rp.data(), col.data(), val.data(), &l.R_stack));
        l.R_stack->lanes = l.R->lanes; // same row structure: keep the restrictor's lanes-per-row choice
        csr_plan_tile(l.R_stack);
        MFMGB_CHECK(mfmgb_vec_alloc(ctx, l.R->n_rows + n_below, &l.rs));
      }
    }
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_hierarchy_set_coarse_offsets(mfmgb_hierarchy *H, const int64_t *offsets, int nranks)
  {
    if (!H || !offsets || nranks < 1 || H->finalized)
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_hierarchy_set_coarse_offsets: bad arguments");
    H->coarse_offsets.assign(offsets, offsets + nranks + 1);
    H->distributed = true;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_hierarchy_set_smoother_chebyshev(mfmgb_hierarchy *H, int degree, double smoothing_range,
                                                       double max_eigenvalue, int eig_cg_n_iterations)
  {
    if (!H || H->finalized || degree < 0 || eig_cg_n_iterations < 0 || eig_cg_n_iterations > 64 ||
        (eig_cg_n_iterations > 0 && eig_cg_n_iterations <= 2))
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_hierarchy_set_smoother_chebyshev: bad arguments (degree >= 0; "
                                                "eig_cg_n_iterations 0 or 3..64, as deal.II asserts)");
    H->chebyshev = true;
    H->cheb_degree = degree;
    H->cheb_range = smoothing_range;
    H->cheb_max_ev = max_eigenvalue;
    H->cheb_cg_its = eig_cg_n_iterations;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_hierarchy_chebyshev_info(const mfmgb_hierarchy *H, int level, double *out4)
  {
    if (!H || !out4 || level < 0 || level >= H->n_levels - 1 || !H->chebyshev || !H->finalized)
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_hierarchy_chebyshev_info: bad arguments");
    const mfmgb_level &l = H->lev[level];
    out4[0] = l.cheb_lmin;
    out4[1] = l.cheb_lmax;
    out4[2] = l.cheb_theta;
    out4[3] = l.cheb_delta;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_hierarchy_set_coarse_dd(mfmgb_hierarchy *H, const mfmgb_coarse_dd *dd)
  {
    if (!H || !dd || H->finalized)
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_hierarchy_set_coarse_dd: bad arguments");
    H->dd = dd;
    return MFMGB_OK;
  }

  MFMGB_API int64_t mfmgb_hierarchy_vector_size(const mfmgb_hierarchy *H, int level)
  {
    if (!H || level < 0 || level >= H->n_levels)
      return 0;
    const mfmgb_level &l = H->lev[level];
    return l.n + (l.halo ? l.halo->n_ghost : 0);
  }

  MFMGB_API int mfmgb_hierarchy_finalize(mfmgb_ctx *ctx, mfmgb_hierarchy *H)
  {
    NvtxRange nvtx_range("mfmgb: hierarchy setup (finalize)");
    MFMGB_REQUIRE(ctx, ctx && H && !H->finalized, "mfmgb_hierarchy_finalize: bad arguments");
    for (int li = 0; li < H->n_levels; ++li)
    {
      mfmgb_level &l = H->lev[li];
      if (!l.A && !l.M)
        return fail(ctx, MFMGB_ERR_INVALID, "mfmgb_hierarchy_finalize: level %d has no operator", li);
      if (li > 0)
      {
        if (!l.R)
          return fail(ctx, MFMGB_ERR_INVALID, "mfmgb_hierarchy_finalize: level %d has no restrictor", li);
        const mfmgb_level &lf = H->lev[li - 1];
        const int64_t fine_cols = lf.n + (lf.halo ? lf.halo->n_ghost : 0);
        if (H->distributed)
        {
          mfmgb_comm *c = ctx_comm(ctx);
          if (!c || (int)H->coarse_offsets.size() != c->nranks + 1)
            return fail(ctx, MFMGB_ERR_INVALID, "mfmgb_hierarchy_finalize: partitioned hierarchy needs mfmgb_comm_init "
                                                "and mfmgb_hierarchy_set_coarse_offsets");
          if (li != H->n_levels - 1 || !l.P)
            return fail(ctx, MFMGB_ERR_INVALID, "mfmgb_hierarchy_finalize: partitioned mode supports a replicated "
                                                "coarsest level right below the partitioned one, with an explicit P");
          const int64_t nc_own = H->coarse_offsets[c->rank + 1] - H->coarse_offsets[c->rank];
          if (l.R->n_rows != nc_own || l.R->n_cols != fine_cols || l.P->n_rows != lf.n || l.P->n_cols != l.n ||
              H->coarse_offsets.back() != l.n)
            return fail(ctx, MFMGB_ERR_INVALID, "mfmgb_hierarchy_finalize: partitioned R / P shapes are inconsistent");
        }
        else if (l.R->n_rows != l.n || l.R->n_cols != H->lev[li - 1].n)
          return fail(ctx, MFMGB_ERR_INVALID,
                      "mfmgb_hierarchy_finalize: restrictor of level %d is %lld x %lld, expected %lld x %lld", li,
                      (long long)l.R->n_rows, (long long)l.R->n_cols, (long long)l.n, (long long)H->lev[li - 1].n);
        if (!l.P)
        {
          MFMGB_CHECK(mfmgb_csr_transpose(ctx, l.R, &l.P_owned));
          l.P = l.P_owned;
        }
        MFMGB_CHECK(mfmgb_vec_alloc(ctx, l.n, &l.bc));
        MFMGB_CHECK(mfmgb_vec_alloc(ctx, l.n, &l.xc));
      }
      if (li < H->n_levels - 1)
      {
        // build_smoother, hierarchy.hpp:204
        if (l.M)
        {
          double *diag = nullptr;
          MFMGB_CHECK(mfmgb_vec_alloc(ctx, l.n, &diag));
          MFMGB_CHECK(mfmgb_mf_diagonal(ctx, l.M, diag));
          MFMGB_CHECK(mfmgb_jacobi_setup_diag(ctx, diag, l.n, H->omega, &l.J));
          MFMGB_CHECK(mfmgb_vec_free(ctx, diag));
        }
        else
          MFMGB_CHECK(mfmgb_jacobi_setup(ctx, l.A, H->omega, &l.J));
        const int64_t ng = l.halo ? l.halo->n_ghost : 0;
        MFMGB_CHECK(mfmgb_vec_alloc(ctx, l.n + ng, &l.res));
        MFMGB_CHECK(mfmgb_vec_alloc(ctx, l.n + ng, &l.xtmp));
        if (H->chebyshev)
        {
          if (H->distributed)
            return fail(ctx, MFMGB_ERR_NOT_IMPLEMENTED, "mfmgb_hierarchy_finalize: the Chebyshev smoother is implemented "
                                                        "for single-GPU hierarchies (the reference pairs it with its "
                                                        "host matrix-free path only)");
          for (double **v : {&l.c_r, &l.c_dst, &l.c_u1, &l.c_u2})
            MFMGB_CHECK(mfmgb_vec_alloc(ctx, l.n + ng, v));
          MFMGB_CHECK(cheb_estimate(ctx, H, l));
          // degree 0 == damped Jacobi with omega = 1 / theta: served by the fused Jacobi sweep
          if (H->cheb_degree == 0)
            l.J->omega = 1. / l.cheb_theta;
        }
      }
      else
      {
        // build_coarse_solver, hierarchy.hpp:194
        if (!l.A)
          return fail(ctx, MFMGB_ERR_INVALID, "mfmgb_hierarchy_finalize: the coarsest level needs an assembled operator");
        if (H->dd)
          continue;
        MFMGB_CHECK(mfmgb_dense_factor(ctx, l.A, &l.D));
        // partitioned hierarchy without the domain-decomposed solve: from n_c = 8192 on, splitting the GEMV by rows
        // across the ranks (one all-gather) beats every rank streaming the whole inverse (MFMGB_DENSE_SPLIT_MIN overrides)
        if (H->distributed)
        {
          int64_t split_min = 8192;
          if (const char *env = getenv("MFMGB_DENSE_SPLIT_MIN"))
            split_min = atoll(env);
          if (l.n >= split_min)
            MFMGB_CHECK(dense_enable_distributed(ctx, l.D));
        }
      }
    }
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // count the launches of one cycle (dry capture, nothing executes)
    if (!H->distributed)
    {
      double *tb = nullptr, *tx = nullptr;
      MFMGB_CHECK(mfmgb_vec_alloc(ctx, H->lev[0].n, &tb));
      MFMGB_CHECK(mfmgb_vec_alloc(ctx, H->lev[0].n, &tx));
      MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      cudaGraph_t graph = nullptr;
      const int64_t before = ctx->launches;
      MFMGB_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
      const int rc = apply_level(ctx, H, tb, tx, 0);
      cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
      if (graph)
        cudaGraphDestroy(graph);
      H->launches_per_cycle = (int)(ctx->launches - before);
      ctx->launches = before;
      MFMGB_CHECK(rc);
      MFMGB_CUDA(ctx, ce);
      MFMGB_CHECK(mfmgb_vec_free(ctx, tb));
      MFMGB_CHECK(mfmgb_vec_free(ctx, tx));
    }
    H->finalized = true;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_hierarchy_destroy(mfmgb_ctx *ctx, mfmgb_hierarchy *H)
  {
    if (!H)
      return MFMGB_OK;
    MFMGB_REQUIRE(ctx, ctx, "ctx is NULL");
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    drop_graph(H);
    for (int k = 0; k < mfmgb_hierarchy::kBatchBuffers; ++k)
    {
      cudaFree(H->bb_dev[k]);
      cudaFree(H->bx_dev[k]);
      if (H->ev_h2d[k])
        cudaEventDestroy(H->ev_h2d[k]);
      if (H->ev_comp[k])
        cudaEventDestroy(H->ev_comp[k]);
      if (H->ev_d2h[k])
        cudaEventDestroy(H->ev_d2h[k]);
    }
    if (H->s_in)
      cudaStreamDestroy(H->s_in);
    if (H->s_out)
      cudaStreamDestroy(H->s_out);
    for (size_t k = 0; k < ctx->graph_owners.size(); ++k)
      if (ctx->graph_owners[k].first == H)
      {
        ctx->graph_owners.erase(ctx->graph_owners.begin() + (long)k);
        break;
      }
    for (int k = 0; k < 7; ++k)
      if (H->ev[k])
        cudaEventDestroy(H->ev[k]);
    for (auto &l : H->lev)
    {
      if (l.P_owned)
        mfmgb_csr_destroy(ctx, l.P_owned);
      mfmgb_jacobi_destroy(ctx, l.J);
      mfmgb_dense_destroy(ctx, l.D);
      cudaFree(l.res);
      cudaFree(l.xtmp);
      cudaFree(l.bc);
      cudaFree(l.xc);
      cudaFree(l.gb);
      cudaFree(l.rs);
      if (l.R_stack)
        mfmgb_csr_destroy(ctx, l.R_stack);
      cudaFree(l.c_r);
      cudaFree(l.c_dst);
      cudaFree(l.c_u1);
      cudaFree(l.c_u2);
    }
    cudaFree(H->b_dev);
    cudaFree(H->x_dev);
    cudaFree(H->g);
    cudaFree(H->h);
    cudaFree(H->d);
    cudaFree(H->scal);
    if (H->scal_host)
      cudaFreeHost(H->scal_host);
    delete H;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_hierarchy_use_graph(mfmgb_hierarchy *H, int on)
  {
    if (!H)
      return fail(nullptr, MFMGB_ERR_INVALID, "mfmgb_hierarchy_use_graph: H is NULL");
    H->use_graph = on != 0;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_hierarchy_launches_per_cycle(const mfmgb_hierarchy *H) { return H ? H->launches_per_cycle : 0; }

  MFMGB_API int mfmgb_hierarchy_apply(mfmgb_ctx *ctx, mfmgb_hierarchy *H, const double *b, double *x, int level)
  {
    MFMGB_REQUIRE(ctx, ctx && H && H->finalized && b && x, "mfmgb_hierarchy_apply: bad arguments");
    MFMGB_REQUIRE(ctx, level >= 0 && level < H->n_levels, "mfmgb_hierarchy_apply: level out of range");
    MFMGB_REQUIRE(ctx, b != x, "mfmgb_hierarchy_apply: b and x must not alias");
    return apply_level(ctx, H, b, x, level);
  }

  MFMGB_API int mfmgb_vcycle(mfmgb_ctx *ctx, mfmgb_hierarchy *H, const double *b, double *x)
  {
    NvtxRange nvtx_range("mfmgb: V-cycle apply");
    MFMGB_REQUIRE(ctx, ctx && H && H->finalized && b && x, "mfmgb_vcycle: bad arguments");
    MFMGB_REQUIRE(ctx, b != x, "mfmgb_vcycle: b and x must not alias");
    return run_vcycle(ctx, H, b, x);
  }

  MFMGB_API int mfmgb_vcycle_profile(mfmgb_ctx *ctx, mfmgb_hierarchy *H, const double *b, double *x, double *stage_ms)
  {
    MFMGB_REQUIRE(ctx, ctx && H && H->finalized && b && x && stage_ms, "mfmgb_vcycle_profile: bad arguments");
    MFMGB_REQUIRE(ctx, H->n_levels >= 2, "mfmgb_vcycle_profile: needs at least two levels");
    for (int k = 0; k < 7; ++k)
      if (!H->ev[k])
        MFMGB_CUDA(ctx, cudaEventCreate(&H->ev[k]));
    H->profiling = true;
    const int rc = apply_level(ctx, H, b, x, 0);
    H->profiling = false;
    MFMGB_CHECK(rc);
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 6; ++k)
    {
      float ms = 0.f;
      MFMGB_CUDA(ctx, cudaEventElapsedTime(&ms, H->ev[k], H->ev[k + 1]));
      stage_ms[k] = (double)ms;
    }
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_vcycle_timeline(mfmgb_ctx *ctx, mfmgb_hierarchy *H, const double *b, double *x, int use_graph,
                                      char *names, int names_len, double *ms, int max_marks, int *n_marks)
  {
    MFMGB_REQUIRE(ctx, ctx && H && H->finalized && b && x && names && names_len > 0 && ms && n_marks,
                  "mfmgb_vcycle_timeline: bad arguments");
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    ctx->prof_on = true;
    ctx->prof_n = 0;
    int rc = MFMGB_OK;
    const int64_t launches_before = ctx->launches;
    if (use_graph)
    {
      MFMGB_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
      ctx->prof_capturing = true;
      rc = apply_level(ctx, H, b, x, 0);
      cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
      ctx->prof_on = ctx->prof_capturing = false;
      if (rc == MFMGB_OK && ce == cudaSuccess)
        ce = cudaGraphInstantiate(&exec, graph, 0);
      if (graph)
        cudaGraphDestroy(graph);
      MFMGB_CHECK(rc);
      MFMGB_CUDA(ctx, ce);
      ctx->launches = launches_before;
      for (int rep = 0; rep < 3; ++rep) // the event nodes keep the timestamps of the last replay
        MFMGB_CUDA(ctx, cudaGraphLaunch(exec, ctx->stream));
    }
    else
    {
      rc = apply_level(ctx, H, b, x, 0);
      ctx->prof_on = false;
      MFMGB_CHECK(rc);
    }
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (exec)
      cudaGraphExecDestroy(exec);
    std::string joined;
    int n = 0;
    for (int k = 1; k < ctx->prof_n && n < max_marks; ++k)
    {
      float t = 0.f;
      MFMGB_CUDA(ctx, cudaEventElapsedTime(&t, ctx->prof_ev[(size_t)k - 1], ctx->prof_ev[(size_t)k]));
      ms[n++] = (double)t;
      if (!joined.empty())
        joined += ";";
      joined += ctx->prof_names[(size_t)k];
    }
    *n_marks = n;
    snprintf(names, (size_t)names_len, "%s", joined.c_str());
    return mfmgb_comm_check(ctx);
  }

  MFMGB_API int mfmgb_vcycle_host(mfmgb_ctx *ctx, mfmgb_hierarchy *H, const double *b_host, double *x_host)
  {
    NvtxRange nvtx_range("mfmgb: V-cycle apply (host vectors)");
    MFMGB_REQUIRE(ctx, ctx && H && H->finalized && b_host && x_host, "mfmgb_vcycle_host: bad arguments");
    const int64_t n = H->lev[0].n;
    MFMGB_CHECK(ensure_host_staging(ctx, H, n));
    const size_t bytes = sizeof(double) * (size_t)n;
    MFMGB_CUDA(ctx, cudaMemcpyAsync(H->b_dev, b_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (!H->is_preconditioner) // solver mode: x is an input as well
      MFMGB_CUDA(ctx, cudaMemcpyAsync(H->x_dev, x_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    MFMGB_CHECK(run_vcycle(ctx, H, H->b_dev, H->x_dev));
    MFMGB_CUDA(ctx, cudaMemcpyAsync(x_host, H->x_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return mfmgb_comm_check(ctx);
  }

  MFMGB_API int mfmgb_vcycle_host_batch(mfmgb_ctx *ctx, mfmgb_hierarchy *H, int n_rhs, const double *const *b_host,
                                        double *const *x_host)
  {
    NvtxRange nvtx_range("mfmgb: V-cycle apply (host vector batch)");
    MFMGB_REQUIRE(ctx, ctx && H && H->finalized && n_rhs >= 0 && (n_rhs == 0 || (b_host && x_host)),
                  "mfmgb_vcycle_host_batch: bad arguments");
    if (!H->is_preconditioner) // solver mode reads x as well: the plain entry point, one right-hand side at a time
    {
      for (int j = 0; j < n_rhs; ++j)
        MFMGB_CHECK(mfmgb_vcycle_host(ctx, H, b_host[j], x_host[j]));
      return MFMGB_OK;
    }
    constexpr int NB = mfmgb_hierarchy::kBatchBuffers;
    const int64_t n = H->lev[0].n;
    const int64_t n_vec = mfmgb_hierarchy_vector_size(H, 0);
    const size_t bytes = sizeof(double) * (size_t)n;
    if (!H->s_in)
    {
      MFMGB_CUDA(ctx, cudaStreamCreateWithFlags(&H->s_in, cudaStreamNonBlocking));
      MFMGB_CUDA(ctx, cudaStreamCreateWithFlags(&H->s_out, cudaStreamNonBlocking));
      for (int k = 0; k < NB; ++k)
      {
        MFMGB_CHECK(mfmgb_vec_alloc(ctx, n_vec, &H->bb_dev[k]));
        MFMGB_CHECK(mfmgb_vec_alloc(ctx, n_vec, &H->bx_dev[k]));
        MFMGB_CUDA(ctx, cudaEventCreateWithFlags(&H->ev_h2d[k], cudaEventDisableTiming));
        MFMGB_CUDA(ctx, cudaEventCreateWithFlags(&H->ev_comp[k], cudaEventDisableTiming));
        MFMGB_CUDA(ctx, cudaEventCreateWithFlags(&H->ev_d2h[k], cudaEventDisableTiming));
      }
      MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    // Three stages in flight: H2D of right-hand side j+1 (copy stream in) || V-cycle j (compute stream) || D2H of
    // result j-1 (copy stream out): PCIe runs full duplex, and the cycle hides behind the copies.  Every right-hand
    // side still makes its own trip host -> device -> host.
    for (int j = 0; j < n_rhs; ++j)
    {
      const int k = j % NB;
      if (j >= NB) // buffer k is free for new input once cycle j - NB has consumed it
        MFMGB_CUDA(ctx, cudaStreamWaitEvent(H->s_in, H->ev_comp[k], 0));
      MFMGB_CUDA(ctx, cudaMemcpyAsync(H->bb_dev[k], b_host[j], bytes, cudaMemcpyHostToDevice, H->s_in));
      MFMGB_CUDA(ctx, cudaEventRecord(H->ev_h2d[k], H->s_in));
      MFMGB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, H->ev_h2d[k], 0));
      if (j >= NB) // ... and its output buffer once result j - NB has left
        MFMGB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, H->ev_d2h[k], 0));
      MFMGB_CHECK(run_vcycle(ctx, H, H->bb_dev[k], H->bx_dev[k]));
      MFMGB_CUDA(ctx, cudaEventRecord(H->ev_comp[k], ctx->stream));
      MFMGB_CUDA(ctx, cudaStreamWaitEvent(H->s_out, H->ev_comp[k], 0));
      MFMGB_CUDA(ctx, cudaMemcpyAsync(x_host[j], H->bx_dev[k], bytes, cudaMemcpyDeviceToHost, H->s_out));
      MFMGB_CUDA(ctx, cudaEventRecord(H->ev_d2h[k], H->s_out));
    }
    MFMGB_CUDA(ctx, cudaStreamSynchronize(H->s_out));
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return mfmgb_comm_check(ctx);
  }

  MFMGB_API int mfmgb_pcg(mfmgb_ctx *ctx, mfmgb_hierarchy *H, const mfmgb_csr *A, const double *b, double *x, double tol,
                          int max_it, int *iterations, double *res_hist_host)
  {
    NvtxRange nvtx_range("mfmgb: PCG solve");
    MFMGB_REQUIRE(ctx, ctx && b && x && iterations && max_it >= 0, "mfmgb_pcg: bad arguments");
    MFMGB_REQUIRE(ctx, H || A, "mfmgb_pcg: need a hierarchy or a matrix");
    MFMGB_REQUIRE(ctx, !H || H->finalized, "mfmgb_pcg: hierarchy not finalized");
    mfmgb_level op;
    if (A)
    {
      op.A = A;
      op.n = A->n_rows;
      if (H && H->lev[0].halo && H->lev[0].A == A)
        op = H->lev[0]; // row-partitioned operator: reuse the level's halo plan
    }
    else
      op = H->lev[0];
    const int64_t n = op.n;
    MFMGB_REQUIRE(ctx, !H || H->lev[0].n == n, "mfmgb_pcg: hierarchy and matrix sizes differ");
    mfmgb_hierarchy local;
    mfmgb_hierarchy *W = H ? H : &local; // work vectors live in the hierarchy when there is one
    MFMGB_CHECK(ensure_pcg_work(ctx, W, n));
    double *g = W->g, *h = W->h, *d = W->d, *scal = W->scal, *sh = W->scal_host;
    const int nb = reduce_blocks(ctx, n);
    const int grid_ew = (int)std::min<int64_t>(ceil_div(n, kBlock), (int64_t)ctx->num_sms * 16);
    cudaStream_t st = ctx->stream;
    int rc = MFMGB_OK;
    auto cleanup = [&]() {
      if (!H)
      {
        cudaStreamSynchronize(st);
        cudaFree(local.g);
        cudaFree(local.h);
        cudaFree(local.d);
        cudaFree(local.scal);
        cudaFreeHost(local.scal_host);
      }
    };
#define PCG_TRY(call)                                                                               \
  do                                                                                                \
  {                                                                                                 \
    rc = (call);                                                                                    \
    if (rc != MFMGB_OK)                                                                             \
    {                                                                                               \
      cleanup();                                                                                    \
      return rc;                                                                                    \
    }                                                                                               \
  } while (0)
#define PCG_CUDA(call)                                                                              \
  do                                                                                                \
  {                                                                                                 \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
    {                                                                                               \
      cleanup();                                                                                    \
      return fail(ctx, MFMGB_ERR_CUDA, "mfmgb_pcg: %s: %s", #call, cudaGetErrorString(e__));        \
    }                                                                                               \
  } while (0)

    // g = A x - b ; res0 = |g|
    EpiArgs e;
    e.y = g;
    e.b = b;
    PCG_TRY(level_apply_A(ctx, op, x, Epi::Resid, e));
    PCG_TRY(vec_dot_async(ctx, g, g, n, scal + 3));
    PCG_TRY(allreduce_sum(ctx, scal + 3, 1));
    PCG_CUDA(cudaMemcpyAsync(sh, scal + 3, sizeof(double), cudaMemcpyDeviceToHost, st));
    PCG_CUDA(cudaStreamSynchronize(st));
    double res = std::sqrt(sh[0]);
    if (res_hist_host)
      res_hist_host[0] = res;
    int it = 0;
    bool converged = res <= tol;
    if (!converged)
    {
      // h = M^-1 g ; d = -h ; gh = g.h
      if (H)
        PCG_TRY(run_vcycle(ctx, H, g, h));
      else
        PCG_CUDA(cudaMemcpyAsync(h, g, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
      negate_kernel<<<grid_ew, kBlock, 0, st>>>(n, h, d);
      ctx->launches++;
      PCG_TRY(vec_dot_async(ctx, g, h, n, scal + 0));
      PCG_TRY(allreduce_sum(ctx, scal + 0, 1));
      while (!converged && it < max_it)
      {
        ++it;
        e = EpiArgs();
        e.y = h;
        PCG_TRY(level_apply_A(ctx, op, d, Epi::Spmv, e)); // h = A d
        PCG_TRY(vec_dot_async(ctx, d, h, n, scal + 2));   // dh = d.h
        PCG_TRY(allreduce_sum(ctx, scal + 2, 1));
        pcg_update_kernel<<<nb, kBlock, 0, st>>>(n, scal, h, d, g, x, ctx->red_partials);
        ctx->launches++;
        PCG_TRY(reduce_finalize(ctx, ctx->red_partials, nb, ctx->red_capacity, 1, scal + 3));
        PCG_TRY(allreduce_sum(ctx, scal + 3, 1));
        PCG_CUDA(cudaMemcpyAsync(sh, scal + 3, sizeof(double), cudaMemcpyDeviceToHost, st));
        PCG_CUDA(cudaStreamSynchronize(st));
        PCG_TRY(mfmgb_comm_check(ctx));
        res = std::sqrt(sh[0]);
        if (res_hist_host)
          res_hist_host[it] = res;
        if (res <= tol)
        {
          converged = true;
          break;
        }
        if (H)
          PCG_TRY(run_vcycle(ctx, H, g, h)); // h = M^-1 g
        else
          PCG_CUDA(cudaMemcpyAsync(h, g, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
        PCG_TRY(vec_dot_async(ctx, g, h, n, scal + 1)); // gh_new
        PCG_TRY(allreduce_sum(ctx, scal + 1, 1));
        pcg_direction_kernel<<<grid_ew, kBlock, 0, st>>>(n, scal, h, d);
        ctx->launches++;
        rotate_scalar_kernel<<<1, 1, 0, st>>>(scal);
        ctx->launches++;
      }
    }
    PCG_CUDA(cudaStreamSynchronize(st));
    PCG_CUDA(cudaGetLastError());
    cleanup();
#undef PCG_TRY
#undef PCG_CUDA
    *iterations = it;
    if (!converged)
      return fail(ctx, MFMGB_ERR_NOT_CONVERGED, "mfmgb_pcg: no convergence after %d iterations (residual %g > %g)", it,
                  res, tol);
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_pcg_host(mfmgb_ctx *ctx, mfmgb_hierarchy *H, const mfmgb_csr *A, const double *b_host,
                               double *x_host, double tol, int max_it, int *iterations, double *res_hist_host)
  {
    MFMGB_REQUIRE(ctx, ctx && (H || A) && b_host && x_host, "mfmgb_pcg_host: bad arguments");
    const int64_t n = A ? A->n_rows : H->lev[0].n;
    // a row-partitioned level gathers from [owned | ghost]: the device vectors carry the ghost tail (mfmgb_pcg
    // receives the neighbours' entries there); only the n owned entries travel to and from the host
    int64_t n_vec = n;
    if (H && !H->lev.empty() && H->lev[0].halo)
      n_vec = std::max(n_vec, mfmgb_hierarchy_vector_size(H, 0));
    if (A)
      n_vec = std::max(n_vec, A->n_cols);
    double *b = nullptr, *x = nullptr;
    MFMGB_CHECK(mfmgb_vec_alloc(ctx, n_vec, &b));
    MFMGB_CHECK(mfmgb_vec_alloc(ctx, n_vec, &x));
    MFMGB_CHECK(mfmgb_vec_upload(ctx, b, b_host, n));
    MFMGB_CHECK(mfmgb_vec_upload(ctx, x, x_host, n));
    const int rc = mfmgb_pcg(ctx, H, A, b, x, tol, max_it, iterations, res_hist_host);
    if (rc == MFMGB_OK || rc == MFMGB_ERR_NOT_CONVERGED)
      mfmgb_vec_download(ctx, x, x_host, n);
    mfmgb_vec_free(ctx, b);
    mfmgb_vec_free(ctx, x);
    return rc;
  }
}
