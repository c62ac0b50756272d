// mf_laplace.cu -- matrix-free Q_p Laplace/diffusion operator on a uniform Cartesian grid (p = 1, 2; 2D/3D).
//
// Fills the CudaMatrixFreeOperator slot (source/cuda/cuda_matrix_free_operator.cu:32-37,60-70 forwards to a user
// hook for which the reference ships no device kernel).  The operator is the one defined on the host by
// tests/laplace_matrix_free.hpp:121-156 inside deal.II's MatrixFreeOperators::Base::vmult semantics:
//   per cell: gather u (constrained entries read as 0), gradients at the (p+1)^d Gauss points, scale by the
//   coefficient table entry (cell, q) and JxW, integrate against grad(phi_a), add into y; constrained rows: y_i = x_i.
//
// Kernel design (shared-memory-staged cell kernel, deterministic, no atomics):
//   * a CTA owns a column of nodes (TX-1) x (TY-1) cells wide and sweeps it in z, one cell layer at a time;
//     one thread per cell of the layer (a warp = one x-row of 32 cells);
//   * the p+1 node planes a layer touches live in shared memory (x values, accumulators, constraint flags) in a
//     ring indexed by plane % (p+1); cells one layer/row/column outside the owned region are recomputed (halo
//     cells, ~10 % extra) so that every owned node receives all of its contributions inside one CTA;
//   * the cell operator is evaluated by sum factorisation with compile-time sizes (all in registers);
//   * contributions are added to the shared accumulators in 4 colour phases (x/y parity of the cell), a fixed order
//     => bit-reproducible;
//   * when a node plane is complete the epilogue (y = A x | r = A x - b | Jacobi update) is applied on the way out,
//     so the matrix-free level uses the same fused V-cycle stages as the CSR levels;
//   * the coefficient table is stored SoA on the device ([q][cell]) so each warp-level load is contiguous.
// Algorithmic bytes per apply: 16 n + 8 n_cells (p+1)^d + n.
#include <cmath>
#include <cstdlib>
#include <vector>

#include "mf.cuh"

using namespace mfmgb;

namespace
{
// 1D tables per degree (index P-1): shape values / derivatives at the Gauss points of the unit interval, weights
__constant__ double c_Sall[2][9]; // [P-1][q*(P+1) + a]
__constant__ double c_Dall[2][9];
__constant__ double c_Wall[2][3];

template <int EPI>
__device__ __forceinline__ void mf_epilogue(const EpiArgs &e, int64_t row, double s)
{
  if (EPI == (int)Epi::Spmv)
    e.y[row] = s;
  else if (EPI == (int)Epi::Resid)
    e.y[row] = __dsub_rn(s, e.b[row]);
  else
  {
    const double r = __dsub_rn(s, e.b[row]);
    double t = __dmul_rn(e.dinv[row], r);
    if (e.omega != 1.)
      t = __dmul_rn(e.omega, t);
    e.y[row] = __dsub_rn(e.xin[row], t);
  }
}

// out = K_cell(c) u by sum factorisation.  u, out: [az][ay][ax] (x fastest); c: [qz][qy][qx].
// k[d] = (prod_{e != d} h_e) / h_d folds the Jacobian: grad_phys = grad_ref / h, JxW = w_q prod h.
template <int DIM, int P>
__device__ __forceinline__ void cell_apply(const double *__restrict__ u, const double *__restrict__ c,
                                           double *__restrict__ out, double kx, double ky, double kz)
{
  constexpr int N = P + 1;
  constexpr int NZ = DIM == 3 ? N : 1;
  const double *c_S = c_Sall[P - 1], *c_D = c_Dall[P - 1], *c_W = c_Wall[P - 1];
  // stage 1: contract x -> values (tv) and x-derivatives (tdx) at qx
  double tv[NZ][N][N], tdx[NZ][N][N];
#pragma unroll
  for (int az = 0; az < NZ; ++az)
#pragma unroll
    for (int ay = 0; ay < N; ++ay)
#pragma unroll
      for (int qx = 0; qx < N; ++qx)
      {
        double v = 0., d = 0.;
#pragma unroll
        for (int ax = 0; ax < N; ++ax)
        {
          const double uu = u[(az * N + ay) * N + ax];
          v = fma(c_S[qx * N + ax], uu, v);
          d = fma(c_D[qx * N + ax], uu, d);
        }
        tv[az][ay][qx] = v;
        tdx[az][ay][qx] = d;
      }
  // stage 2: contract y
  double gx2[NZ][N][N], gy2[NZ][N][N], vv2[NZ][N][N];
#pragma unroll
  for (int az = 0; az < NZ; ++az)
#pragma unroll
    for (int qy = 0; qy < N; ++qy)
#pragma unroll
      for (int qx = 0; qx < N; ++qx)
      {
        double a = 0., b = 0., v = 0.;
#pragma unroll
        for (int ay = 0; ay < N; ++ay)
        {
          a = fma(c_S[qy * N + ay], tdx[az][ay][qx], a);
          b = fma(c_D[qy * N + ay], tv[az][ay][qx], b);
          v = fma(c_S[qy * N + ay], tv[az][ay][qx], v);
        }
        gx2[az][qy][qx] = a;
        gy2[az][qy][qx] = b;
        vv2[az][qy][qx] = v;
      }
  // stage 3: contract z, multiply by coefficient * weights, giving the fluxes at the quadrature points
  double fx[NZ][N][N], fy[NZ][N][N], fz[NZ][N][N];
#pragma unroll
  for (int qz = 0; qz < NZ; ++qz)
#pragma unroll
    for (int qy = 0; qy < N; ++qy)
#pragma unroll
      for (int qx = 0; qx < N; ++qx)
      {
        double a = 0., b = 0., g = 0.;
        if (DIM == 3)
        {
#pragma unroll
          for (int az = 0; az < NZ; ++az)
          {
            a = fma(c_S[qz * N + az], gx2[az][qy][qx], a);
            b = fma(c_S[qz * N + az], gy2[az][qy][qx], b);
            g = fma(c_D[qz * N + az], vv2[az][qy][qx], g);
          }
        }
        else
        {
          a = gx2[0][qy][qx];
          b = gy2[0][qy][qx];
        }
        double w = c[(qz * N + qy) * N + qx] * c_W[qx] * c_W[qy];
        if (DIM == 3)
          w *= c_W[qz];
        fx[qz][qy][qx] = w * kx * a;
        fy[qz][qy][qx] = w * ky * b;
        fz[qz][qy][qx] = DIM == 3 ? w * kz * g : 0.;
      }
  // integration = transpose of the above
  // contract z (test functions in z)
  double hx[NZ][N][N], hy[NZ][N][N], hz[NZ][N][N];
#pragma unroll
  for (int az = 0; az < NZ; ++az)
#pragma unroll
    for (int qy = 0; qy < N; ++qy)
#pragma unroll
      for (int qx = 0; qx < N; ++qx)
      {
        if (DIM == 3)
        {
          double a = 0., b = 0., g = 0.;
#pragma unroll
          for (int qz = 0; qz < NZ; ++qz)
          {
            a = fma(c_S[qz * N + az], fx[qz][qy][qx], a);
            b = fma(c_S[qz * N + az], fy[qz][qy][qx], b);
            g = fma(c_D[qz * N + az], fz[qz][qy][qx], g);
          }
          hx[az][qy][qx] = a;
          hy[az][qy][qx] = b;
          hz[az][qy][qx] = g;
        }
        else
        {
          hx[0][qy][qx] = fx[0][qy][qx];
          hy[0][qy][qx] = fy[0][qy][qx];
          hz[0][qy][qx] = 0.;
        }
      }
  // contract y
  double px[NZ][N][N], pv[NZ][N][N];
#pragma unroll
  for (int az = 0; az < NZ; ++az)
#pragma unroll
    for (int ay = 0; ay < N; ++ay)
#pragma unroll
      for (int qx = 0; qx < N; ++qx)
      {
        double a = 0., v = 0.;
#pragma unroll
        for (int qy = 0; qy < N; ++qy)
        {
          a = fma(c_S[qy * N + ay], hx[az][qy][qx], a);
          v = fma(c_D[qy * N + ay], hy[az][qy][qx], v);
          v = fma(c_S[qy * N + ay], hz[az][qy][qx], v);
        }
        px[az][ay][qx] = a; // to be contracted with D in x
        pv[az][ay][qx] = v; // to be contracted with S in x
      }
  // contract x
#pragma unroll
  for (int az = 0; az < NZ; ++az)
#pragma unroll
    for (int ay = 0; ay < N; ++ay)
#pragma unroll
      for (int ax = 0; ax < N; ++ax)
      {
        double r = 0.;
#pragma unroll
        for (int qx = 0; qx < N; ++qx)
        {
          r = fma(c_D[qx * N + ax], px[az][ay][qx], r);
          r = fma(c_S[qx * N + ax], pv[az][ay][qx], r);
        }
        out[(az * N + ay) * N + ax] = r;
      }
}

template <int SLOTS>
__device__ __forceinline__ int slot_of(int64_t G)
{
  int s = (int)(G % SLOTS);
  return s < 0 ? s + SLOTS : s;
}

struct MfParams
{
  int64_t cells[3];
  int64_t nodes[3];
  double kx, ky, kz;
  int64_t n_cells;
  const double *coef;   // [nq][n_cells]
  const uint8_t *constr; // [n]
  int tz;               // owned cell layers per CTA (3D)
};

// TXC x TYC cells per layer per CTA including the one-cell halo on the low side: owned cells are
// [X0, X0 + TXC - 1) x [Y0, Y0 + TYC - 1); thread (tx, ty) handles cell (X0 - 1 + tx, Y0 - 1 + ty).
template <int DIM, int P, int TXC, int TYC, int EPI>
__global__ void __launch_bounds__(TXC *TYC) mf_apply_kernel(MfParams prm, const double *__restrict__ x, EpiArgs e)
{
  constexpr int N = P + 1;
  constexpr int NZ = DIM == 3 ? N : 1;
  constexpr int NDOF = N * N * NZ;
  constexpr int SLOTS = DIM == 3 ? N : 1;
  constexpr int NXS = TXC * P + 1, NYS = TYC * P + 1;
  constexpr int PLANE = NXS * NYS;
  constexpr int NTHREADS = TXC * TYC;
  extern __shared__ double smem[];
  double *xs = smem;                         // [SLOTS][PLANE] raw x
  double *os = smem + SLOTS * PLANE;         // [SLOTS][PLANE] accumulators
  uint8_t *fs = reinterpret_cast<uint8_t *>(smem + 2 * SLOTS * PLANE); // [SLOTS][PLANE] constraint flags

  const int tid = threadIdx.x;
  const int tx = tid % TXC, ty = tid / TXC;
  const int64_t X0 = (int64_t)blockIdx.x * (TXC - 1), Y0 = (int64_t)blockIdx.y * (TYC - 1);
  const int64_t Z0 = DIM == 3 ? (int64_t)blockIdx.z * prm.tz : 0;
  const int64_t cxg = X0 - 1 + tx, cyg = Y0 - 1 + ty; // this thread's cell (may be outside the grid)
  const bool cell_xy_ok = cxg >= 0 && cxg < prm.cells[0] && cyg >= 0 && cyg < prm.cells[1];
  const int64_t gx0 = (X0 - 1) * P, gy0 = (Y0 - 1) * P; // global node index of tile-local node (0, 0)
  // owned node ranges
  const int64_t own_x0 = X0 * P, own_y0 = Y0 * P;
  const int64_t own_x1 = (X0 + TXC - 1 >= prm.cells[0]) ? prm.nodes[0] : (X0 + TXC - 1) * P;
  const int64_t own_y1 = (Y0 + TYC - 1 >= prm.cells[1]) ? prm.nodes[1] : (Y0 + TYC - 1) * P;
  const int64_t zc_end = DIM == 3 ? min(Z0 + (int64_t)prm.tz, prm.cells[2]) : 1; // one past the last owned layer
  const int64_t own_z0 = DIM == 3 ? Z0 * P : 0;
  const int64_t own_z1 = DIM == 3 ? (zc_end >= prm.cells[2] ? prm.nodes[2] : zc_end * P) : 1;
  const int color = (tx & 1) + 2 * (ty & 1);

  // load node plane G (global z index of the plane) into its ring slot, zero the accumulator
  auto load_plane = [&](int64_t G) {
    const int slot = DIM == 3 ? slot_of<SLOTS>(G) : 0;
    const bool z_ok = DIM == 3 ? (G >= 0 && G < prm.nodes[2]) : true;
    for (int i = tid; i < PLANE; i += NTHREADS)
    {
      const int lx = i % NXS, ly = i / NXS;
      const int64_t gx = gx0 + lx, gy = gy0 + ly;
      double v = 0.;
      uint8_t f = 1;
      if (z_ok && gx >= 0 && gx < prm.nodes[0] && gy >= 0 && gy < prm.nodes[1])
      {
        const int64_t g = gx + prm.nodes[0] * (gy + prm.nodes[1] * (DIM == 3 ? G : 0));
        v = x[g];
        f = prm.constr[g];
      }
      xs[slot * PLANE + i] = v;
      os[slot * PLANE + i] = 0.;
      fs[slot * PLANE + i] = f;
    }
  };
  // write the finished plane G (owned nodes only) through the epilogue
  auto store_plane = [&](int64_t G) {
    if (DIM == 3 && (G < own_z0 || G >= own_z1))
      return;
    const int slot = DIM == 3 ? slot_of<SLOTS>(G) : 0;
    for (int i = tid; i < PLANE; i += NTHREADS)
    {
      const int lx = i % NXS, ly = i / NXS;
      const int64_t gx = gx0 + lx, gy = gy0 + ly;
      if (gx >= own_x0 && gx < own_x1 && gy >= own_y0 && gy < own_y1)
      {
        const int64_t g = gx + prm.nodes[0] * (gy + prm.nodes[1] * (DIM == 3 ? G : 0));
        const double s = fs[slot * PLANE + i] ? xs[slot * PLANE + i] : os[slot * PLANE + i];
        mf_epilogue<EPI>(e, g, s);
      }
    }
  };

  // first cell layer: Z0 - 1 (halo layer below the owned region; only its top node plane is owned)
  const int64_t layer_begin = DIM == 3 ? Z0 - 1 : 0;
  const int64_t layer_end = DIM == 3 ? zc_end : 1;
  if (DIM == 3)
  {
    for (int s = 0; s < N; ++s)
      load_plane(layer_begin * P + s);
  }
  else
    load_plane(0);
  __syncthreads();

  for (int64_t lz = layer_begin; lz < layer_end; ++lz)
  {
    const bool cell_ok = cell_xy_ok && (DIM == 2 || (lz >= 0 && lz < prm.cells[2]));
    double out[NDOF];
    if (cell_ok)
    {
      double u[NDOF], c[NDOF];
      const int64_t cell = cxg + prm.cells[0] * (cyg + prm.cells[1] * (DIM == 3 ? lz : 0));
#pragma unroll
      for (int q = 0; q < NDOF; ++q)
        c[q] = ld_stream_f64(prm.coef + (int64_t)q * prm.n_cells + cell);
#pragma unroll
      for (int az = 0; az < NZ; ++az)
      {
        const int slot = DIM == 3 ? slot_of<SLOTS>(lz * P + az) : 0;
#pragma unroll
        for (int ay = 0; ay < N; ++ay)
#pragma unroll
          for (int ax = 0; ax < N; ++ax)
          {
            const int i = slot * PLANE + (ty * P + ay) * NXS + tx * P + ax;
            u[(az * N + ay) * N + ax] = fs[i] ? 0. : xs[i];
          }
      }
      cell_apply<DIM, P>(u, c, out, prm.kx, prm.ky, prm.kz);
    }
    // colour-ordered accumulation: cells of one colour share no node inside a layer
#pragma unroll 1
    for (int col = 0; col < 4; ++col)
    {
      if (cell_ok && col == color)
      {
#pragma unroll
        for (int az = 0; az < NZ; ++az)
        {
          const int slot = DIM == 3 ? slot_of<SLOTS>(lz * P + az) : 0;
#pragma unroll
          for (int ay = 0; ay < N; ++ay)
#pragma unroll
            for (int ax = 0; ax < N; ++ax)
              os[slot * PLANE + (ty * P + ay) * NXS + tx * P + ax] += out[(az * N + ay) * N + ax];
        }
      }
      __syncthreads();
    }
    if (DIM == 3)
    {
      // planes lz*P .. lz*P + P-1 are complete; retire them and bring in the next P planes
      for (int s = 0; s < P; ++s)
        store_plane(lz * P + s);
      __syncthreads();
      if (lz + 1 < layer_end)
      {
        for (int s = 0; s < P; ++s)
          load_plane((lz + 1) * P + 1 + s);
        __syncthreads();
      }
      else
        store_plane((lz + 1) * P); // top plane of the last layer (owned only at the global top)
    }
    else
      store_plane(0);
  }
}

// diagonal: thread per node, gather over the adjacent cells (setup; compute_diagonal,
// tests/laplace_matrix_free.hpp:75-98,158-199: constrained entries := 1)
__global__ void __launch_bounds__(256) mf_diag_kernel(int dim, int p, MfParams prm, const double *__restrict__ gdiag,
                                                      int64_t n, double *__restrict__ diag)
{
  const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (g >= n)
    return;
  if (prm.constr[g])
  {
    diag[g] = 1.;
    return;
  }
  const int n1 = p + 1;
  const int ndof = dim == 3 ? n1 * n1 * n1 : n1 * n1;
  int64_t idx[3] = {g % prm.nodes[0], (g / prm.nodes[0]) % prm.nodes[1], g / (prm.nodes[0] * prm.nodes[1])};
  int64_t c0[3], c1[3];
  for (int d = 0; d < 3; ++d)
  {
    if (d >= dim)
    {
      c0[d] = c1[d] = 0;
      continue;
    }
    if (idx[d] % p == 0)
    {
      const int64_t v = idx[d] / p;
      c0[d] = v > 0 ? v - 1 : 0;
      c1[d] = v < prm.cells[d] ? v : prm.cells[d] - 1;
    }
    else
      c0[d] = c1[d] = idx[d] / p;
  }
  double s = 0.;
  for (int64_t cz = c0[2]; cz <= c1[2]; ++cz)
    for (int64_t cy = c0[1]; cy <= c1[1]; ++cy)
      for (int64_t cx = c0[0]; cx <= c1[0]; ++cx)
      {
        const int64_t cell = cx + prm.cells[0] * (cy + prm.cells[1] * cz);
        const int ax = (int)(idx[0] - cx * p), ay = dim > 1 ? (int)(idx[1] - cy * p) : 0,
                  az = dim > 2 ? (int)(idx[2] - cz * p) : 0;
        const int a = ax + n1 * (ay + n1 * az);
        for (int q = 0; q < ndof; ++q)
          s = fma(prm.coef[(int64_t)q * prm.n_cells + cell], gdiag[q * ndof + a], s);
      }
  diag[g] = s;
}

void gauss_unit(int nq, double *pts, double *wts)
{
  if (nq == 2)
  {
    const double a = 0.5 / std::sqrt(3.);
    pts[0] = 0.5 - a;
    pts[1] = 0.5 + a;
    wts[0] = wts[1] = 0.5;
  }
  else
  {
    const double a = 0.5 * std::sqrt(0.6);
    pts[0] = 0.5 - a;
    pts[1] = 0.5;
    pts[2] = 0.5 + a;
    wts[0] = wts[2] = 5. / 18.;
    wts[1] = 8. / 18.;
  }
}

MfParams make_params(const mfmgb_mf *M)
{
  MfParams p;
  for (int d = 0; d < 3; ++d)
  {
    p.cells[d] = M->cells[d];
    p.nodes[d] = M->nodes[d];
  }
  const double *h = M->h;
  if (M->dim == 3)
  {
    p.kx = h[1] * h[2] / h[0];
    p.ky = h[0] * h[2] / h[1];
    p.kz = h[0] * h[1] / h[2];
  }
  else
  {
    p.kx = h[1] / h[0];
    p.ky = h[0] / h[1];
    p.kz = 0.;
  }
  p.n_cells = M->n_cells;
  p.coef = M->coef;
  p.constr = M->constr;
  p.tz = 32;
  return p;
}

template <int DIM, int P, int TXC, int TYC, int EPI>
int launch_mf(mfmgb_ctx *ctx, const mfmgb_mf *M, const double *x, const EpiArgs &e)
{
  constexpr int N = P + 1;
  constexpr int SLOTS = DIM == 3 ? N : 1;
  constexpr int PLANE = (TXC * P + 1) * (TYC * P + 1);
  const size_t smem = (size_t)SLOTS * PLANE * (2 * sizeof(double) + 1) + 16;
  auto kern = mf_apply_kernel<DIM, P, TXC, TYC, EPI>;
  static unsigned long long configured = 0; // one bit per device: the attribute is a per-device property
  if (!((configured >> (ctx->device & 63)) & 1ull))
  {
    MFMGB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured |= 1ull << (ctx->device & 63);
  }
  MfParams prm = make_params(M);
  dim3 grid((unsigned)ceil_div(M->cells[0], TXC - 1), (unsigned)ceil_div(M->cells[1], TYC - 1),
            DIM == 3 ? (unsigned)ceil_div(M->cells[2], prm.tz) : 1u);
  kern<<<grid, TXC * TYC, smem, ctx->stream>>>(prm, x, e);
  MFMGB_LAUNCHED(ctx);
  return MFMGB_OK;
}

template <int DIM, int P, int TXC, int TYC>
int dispatch_mf_epi(mfmgb_ctx *ctx, const mfmgb_mf *M, const double *x, Epi epi, const EpiArgs &e)
{
  switch (epi)
  {
  case Epi::Spmv:
    return launch_mf<DIM, P, TXC, TYC, (int)Epi::Spmv>(ctx, M, x, e);
  case Epi::Resid:
    return launch_mf<DIM, P, TXC, TYC, (int)Epi::Resid>(ctx, M, x, e);
  case Epi::Jacobi:
    return launch_mf<DIM, P, TXC, TYC, (int)Epi::Jacobi>(ctx, M, x, e);
  default:
    return fail(ctx, MFMGB_ERR_INVALID, "mf_apply: unsupported epilogue");
  }
}
} // namespace

#include "mf_q1.cuh"

namespace mfmgb
{
int mf_apply(mfmgb_ctx *ctx, const mfmgb_mf *M, const double *x, Epi epi, const EpiArgs &args)
{
  if (M->dim == 3 && M->degree == 1 && !M->force_generic)
    return mf_q1_apply(ctx, M, x, epi, args, -1, -1); // node-owner z-sweep (mf_q1.cuh)
  if (M->own0 != 0 || M->own1 != M->nodes[M->dim - 1])
    return fail(ctx, MFMGB_ERR_NOT_IMPLEMENTED, "mf_apply: slab layouts are implemented for 3D Q1 only");
  if (M->dim == 3 && M->degree == 1)
    return dispatch_mf_epi<3, 1, 32, 16>(ctx, M, x, epi, args);
  if (M->dim == 3 && M->degree == 2)
    return dispatch_mf_epi<3, 2, 32, 8>(ctx, M, x, epi, args);
  if (M->dim == 2 && M->degree == 1)
    return dispatch_mf_epi<2, 1, 32, 16>(ctx, M, x, epi, args);
  if (M->dim == 2 && M->degree == 2)
    return dispatch_mf_epi<2, 2, 32, 8>(ctx, M, x, epi, args);
  return fail(ctx, MFMGB_ERR_NOT_IMPLEMENTED, "mf_apply: dim %d degree %d not implemented", M->dim, M->degree);
}
// partitioned level: middle z chunks first (they read owned planes only), the two end chunks after the halo has landed
int mf_apply_chunks(mfmgb_ctx *ctx, const mfmgb_mf *M, const double *x, Epi epi, const EpiArgs &args, int zc0, int zc1)
{
  if (!(M->dim == 3 && M->degree == 1 && !M->force_generic))
    return fail(ctx, MFMGB_ERR_NOT_IMPLEMENTED, "mf_apply_chunks: 3D Q1 only");
  return mf_q1_apply(ctx, M, x, epi, args, zc0, zc1);
}
int mf_num_chunks(const mfmgb_mf *M)
{
  return M->dim == 3 && M->degree == 1 && !M->force_generic ? mf_q1_num_chunks(M) : 1;
}
} // namespace mfmgb

extern "C"
{
  MFMGB_API int mfmgb_mf_laplace_create(mfmgb_ctx *ctx, int dim, int degree, const int64_t *cells, const double *h,
                                        const double *coef, const uint8_t *constrained, mfmgb_mf **out)
  {
    MFMGB_REQUIRE(ctx, ctx && cells, "mfmgb_mf_laplace_create: bad arguments");
    const int64_t top = (dim == 2 || dim == 3) ? cells[dim - 1] * degree + 1 : 0;
    return mfmgb_mf_laplace_create_slab(ctx, dim, degree, cells, h, coef, constrained, 0, top, out);
  }

  MFMGB_API int mfmgb_mf_laplace_create_slab(mfmgb_ctx *ctx, int dim, int degree, const int64_t *cells, const double *h,
                                             const double *coef, const uint8_t *constrained, int64_t own_plane_begin,
                                             int64_t own_plane_end, mfmgb_mf **out)
  {
    MFMGB_REQUIRE(ctx, ctx && cells && h && coef && constrained && out, "mfmgb_mf_laplace_create: bad arguments");
    if (!((dim == 2 || dim == 3) && (degree == 1 || degree == 2)))
      return fail(ctx, MFMGB_ERR_NOT_IMPLEMENTED, "mfmgb_mf_laplace_create: dim %d degree %d not implemented (2D/3D, Q1/Q2)",
                  dim, degree);
    mfmgb_mf *M = new mfmgb_mf();
    M->dim = dim;
    M->degree = degree;
    int64_t n_local = 1;
    M->n_cells = 1;
    for (int d = 0; d < 3; ++d)
    {
      M->cells[d] = d < dim ? cells[d] : 1;
      M->h[d] = d < dim ? h[d] : 1.;
      M->nodes[d] = d < dim ? cells[d] * degree + 1 : 1;
      n_local *= M->nodes[d];
      M->n_cells *= M->cells[d];
      MFMGB_REQUIRE(ctx, M->cells[d] >= 1 && M->h[d] > 0., "mfmgb_mf_laplace_create: bad grid");
    }
    M->n_local = n_local;
    M->own0 = own_plane_begin;
    M->own1 = own_plane_end;
    MFMGB_REQUIRE(ctx, own_plane_begin >= 0 && own_plane_begin < own_plane_end && own_plane_end <= M->nodes[dim - 1],
                  "mfmgb_mf_laplace_create_slab: owned planes out of range");
    const bool slab = own_plane_begin != 0 || own_plane_end != M->nodes[dim - 1];
    if (slab && !(dim == 3 && degree == 1))
    {
      delete M;
      return fail(ctx, MFMGB_ERR_NOT_IMPLEMENTED, "mfmgb_mf_laplace_create_slab: slab layouts are implemented for 3D Q1");
    }
    M->n = (own_plane_end - own_plane_begin) * (n_local / M->nodes[dim - 1]); // rows = owned nodes
    const int n1 = degree + 1;
    M->nq = dim == 3 ? n1 * n1 * n1 : n1 * n1;
    // 3D Q1 fast path: a table whose entries are equal within every cell (any piecewise-constant material) is
    // stored once per cell and applied through the reference cell matrix
    const char *gen = getenv("MFMGB_MF_GENERIC");
    M->force_generic = gen && gen[0] == '1';
    bool cell_constant = dim == 3 && degree == 1 && !M->force_generic;
    for (int64_t c = 0; c < M->n_cells && cell_constant; ++c)
      for (int q = 1; q < M->nq; ++q)
        if (coef[(size_t)c * M->nq + q] != coef[(size_t)c * M->nq])
        {
          cell_constant = false;
          break;
        }
    M->q1_cell_constant = cell_constant;
    if (cell_constant)
    {
      // one coefficient for the whole grid, and all 8 cells exist around every owned unconstrained node (vector layout
      // [owned planes | ghost planes]): the factorised-stencil z-sweep applies (MFMGB_MF_STENCIL=0 keeps the cell kernel)
      const char *st = getenv("MFMGB_MF_STENCIL");
      bool stencil = !(st && st[0] == '0') && M->nodes[0] >= 3 && M->nodes[1] >= 3;
      for (int64_t c = 1; c < M->n_cells && stencil; ++c)
        stencil = coef[(size_t)c * M->nq] == coef[0];
      const int64_t nx = M->nodes[0], ny = M->nodes[1], pl = nx * ny;
      for (int64_t r = 0; r < M->n && stencil; ++r)
      {
        const int64_t g = own_plane_begin + r / pl, j = (r % pl) / nx, i = r % nx;
        if ((i == 0 || i == nx - 1 || j == 0 || j == ny - 1 || g == 0 || g == M->nodes[2] - 1) && !constrained[r])
          stencil = false;
      }
      M->q1_stencil = stencil;
      M->q1_const_coef = coef[0];
      if (stencil)
      {
        // are the flags of the WHOLE local box (vector layout [owned | ghost below | ghost above]) the box's x / y faces
        // plus, possibly, its first / last plane?  Then the kernel computes them.
        const int64_t nz = M->nodes[2];
        auto vec_off = [&](int64_t g) {
          if (g >= own_plane_begin && g < own_plane_end)
            return (g - own_plane_begin) * pl;
          if (g < own_plane_begin)
            return M->n + g * pl;
          return M->n + (own_plane_begin + (g - own_plane_end)) * pl;
        };
        auto interior_flag = [&](int64_t g) { return constrained[vec_off(g) + (ny / 2) * nx + nx / 2] != 0; };
        M->q1_bottom_bc = interior_flag(0);
        M->q1_top_bc = interior_flag(nz - 1);
        bool arith = true;
        for (int64_t g = 0; g < nz && arith; ++g)
        {
          const bool face = (g == 0 && M->q1_bottom_bc) || (g == nz - 1 && M->q1_top_bc);
          const uint8_t *f = constrained + vec_off(g);
          for (int64_t j = 0; j < ny && arith; ++j)
            for (int64_t i = 0; i < nx; ++i)
              if ((f[j * nx + i] != 0) != (face || i == 0 || i == nx - 1 || j == 0 || j == ny - 1))
              {
                arith = false;
                break;
              }
        }
        const char *af = getenv("MFMGB_MF_ARITH_FLAGS");
        M->q1_arith_flags = arith && !(af && af[0] == '0');
      }
      std::vector<double> cc((size_t)M->n_cells);
      for (int64_t c = 0; c < M->n_cells; ++c)
        cc[(size_t)c] = coef[(size_t)c * M->nq];
      MFMGB_CUDA(ctx, cudaMalloc(&M->coef_cell, sizeof(double) * cc.size()));
      MFMGB_CUDA(ctx, cudaMemcpy(M->coef_cell, cc.data(), sizeof(double) * cc.size(), cudaMemcpyHostToDevice));
    }
    else
    {
      // transpose the (cell, q) table of tests/laplace_matrix_free.hpp:100-119 to SoA [q][cell]
      std::vector<double> soa((size_t)M->nq * (size_t)M->n_cells);
      for (int64_t c = 0; c < M->n_cells; ++c)
        for (int q = 0; q < M->nq; ++q)
          soa[(size_t)q * M->n_cells + c] = coef[(size_t)c * M->nq + q];
      MFMGB_CUDA(ctx, cudaMalloc(&M->coef, sizeof(double) * soa.size()));
      MFMGB_CUDA(ctx, cudaMemcpy(M->coef, soa.data(), sizeof(double) * soa.size(), cudaMemcpyHostToDevice));
    }
    MFMGB_CUDA(ctx, cudaMalloc(&M->constr, (size_t)M->n_local + 16));
    MFMGB_CUDA(ctx, cudaMemcpy(M->constr, constrained, (size_t)M->n_local, cudaMemcpyHostToDevice));
    // 1D tables: Lagrange basis on equidistant nodes at the Gauss points of the unit interval, for p = 1 and 2
    double Sall[2][9] = {{0}}, Dall[2][9] = {{0}}, Wall[2][3] = {{0}};
    for (int deg = 1; deg <= 2; ++deg)
    {
      const int m1 = deg + 1;
      double pts[3];
      gauss_unit(m1, pts, Wall[deg - 1]);
      for (int q = 0; q < m1; ++q)
        for (int a = 0; a < m1; ++a)
        {
          const double xa = (double)a / deg;
          double v = 1., dv = 0.;
          for (int c = 0; c < m1; ++c)
            if (c != a)
              v *= (pts[q] - (double)c / deg) / (xa - (double)c / deg);
          for (int e = 0; e < m1; ++e)
            if (e != a)
            {
              double t = 1. / (xa - (double)e / deg);
              for (int c = 0; c < m1; ++c)
                if (c != a && c != e)
                  t *= (pts[q] - (double)c / deg) / (xa - (double)c / deg);
              dv += t;
            }
          Sall[deg - 1][q * m1 + a] = v;
          Dall[deg - 1][q * m1 + a] = dv;
        }
    }
    MFMGB_CUDA(ctx, cudaMemcpyToSymbol(c_Sall, Sall, sizeof(Sall)));
    MFMGB_CUDA(ctx, cudaMemcpyToSymbol(c_Dall, Dall, sizeof(Dall)));
    MFMGB_CUDA(ctx, cudaMemcpyToSymbol(c_Wall, Wall, sizeof(Wall)));
    for (int i = 0; i < 9; ++i)
    {
      M->S[i] = Sall[degree - 1][i];
      M->D[i] = Dall[degree - 1][i];
    }
    for (int i = 0; i < 3; ++i)
      M->W[i] = Wall[degree - 1][i];
    if (dim == 3 && degree == 1)
    {
      // K_ref[a][b] = sum_q (grad phi_a . grad phi_b)(x_q) JxW_q on a cell of size h
      for (int a = 0; a < 8; ++a)
        for (int b = 0; b < 8; ++b)
        {
          double s = 0.;
          for (int q = 0; q < 8; ++q)
          {
            const int qx = q & 1, qy = (q >> 1) & 1, qz = q >> 2;
            const double jxw = M->W[qx] * M->W[qy] * M->W[qz] * h[0] * h[1] * h[2];
            double g[2][3];
            const int ab[2] = {a, b};
            for (int t = 0; t < 2; ++t)
            {
              const int ax = ab[t] & 1, ay = (ab[t] >> 1) & 1, az = ab[t] >> 2;
              const double sx = M->S[qx * 2 + ax], sy = M->S[qy * 2 + ay], sz = M->S[qz * 2 + az];
              g[t][0] = M->D[qx * 2 + ax] / h[0] * sy * sz;
              g[t][1] = sx * M->D[qy * 2 + ay] / h[1] * sz;
              g[t][2] = sx * sy * M->D[qz * 2 + az] / h[2];
            }
            s += (g[0][0] * g[1][0] + g[0][1] * g[1][1] + g[0][2] * g[1][2]) * jxw;
          }
          M->Kref[a * 8 + b] = s;
        }
      if (!M->force_generic)
        MFMGB_CHECK(mf_q1_prepare(ctx, M));
    }
    *out = M;
    return MFMGB_OK;
  }

  MFMGB_API int mfmgb_mf_destroy(mfmgb_ctx *ctx, mfmgb_mf *M)
  {
    if (!M)
      return MFMGB_OK;
    MFMGB_REQUIRE(ctx, ctx, "ctx is NULL");
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(M->coef);
    cudaFree(M->coef_cell);
    cudaFree(M->brick_flags);
    cudaFree(M->constr);
    delete M;
    return MFMGB_OK;
  }

  MFMGB_API int64_t mfmgb_mf_size(const mfmgb_mf *M) { return M ? M->n : 0; }
  MFMGB_API int64_t mfmgb_mf_vector_size(const mfmgb_mf *M) { return M ? M->n_local : 0; }
  MFMGB_API int mfmgb_mf_kernel(const mfmgb_mf *M)
  {
    return !M ? -1
              : (M->dim == 3 && M->degree == 1 && !M->force_generic ? (M->q1_stencil ? 3 : (M->q1_cell_constant ? 1 : 2))
                                                                    : 0);
  }

  MFMGB_API int mfmgb_mf_apply(mfmgb_ctx *ctx, const mfmgb_mf *M, const double *x, double *y)
  {
    MFMGB_REQUIRE(ctx, ctx && M && x && y && x != y, "mfmgb_mf_apply: bad arguments");
    EpiArgs e;
    e.y = y;
    return mf_apply(ctx, M, x, Epi::Spmv, e);
  }

  MFMGB_API int mfmgb_mf_diagonal(mfmgb_ctx *ctx, const mfmgb_mf *M, double *diag_dev)
  {
    MFMGB_REQUIRE(ctx, ctx && M && diag_dev, "mfmgb_mf_diagonal: bad arguments");
    // gdiag[q][a] = (grad phi_a . grad phi_a)(x_q) JxW_q
    const int n1 = M->degree + 1, dim = M->dim;
    const int ndof = M->nq;
    std::vector<double> gd((size_t)ndof * ndof, 0.);
    const double *h = M->h;
    for (int qz = 0; qz < (dim == 3 ? n1 : 1); ++qz)
      for (int qy = 0; qy < n1; ++qy)
        for (int qx = 0; qx < n1; ++qx)
        {
          const int q = (qz * n1 + qy) * n1 + qx;
          double jxw = M->W[qx] * M->W[qy] * h[0] * h[1];
          if (dim == 3)
            jxw *= M->W[qz] * h[2];
          for (int az = 0; az < (dim == 3 ? n1 : 1); ++az)
            for (int ay = 0; ay < n1; ++ay)
              for (int ax = 0; ax < n1; ++ax)
              {
                const int a = (az * n1 + ay) * n1 + ax;
                const double sx = M->S[qx * n1 + ax], sy = M->S[qy * n1 + ay], sz = dim == 3 ? M->S[qz * n1 + az] : 1.;
                const double dx = M->D[qx * n1 + ax] / h[0], dy = M->D[qy * n1 + ay] / h[1],
                             dz = dim == 3 ? M->D[qz * n1 + az] / h[2] : 0.;
                const double g0 = dx * sy * sz, g1 = sx * dy * sz, g2 = sx * sy * dz;
                gd[(size_t)q * ndof + a] = (g0 * g0 + g1 * g1 + g2 * g2) * jxw;
              }
        }
    double *gd_dev = nullptr;
    MFMGB_CUDA(ctx, cudaMalloc(&gd_dev, sizeof(double) * gd.size()));
    MFMGB_CUDA(ctx, cudaMemcpyAsync(gd_dev, gd.data(), sizeof(double) * gd.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (dim == 3 && M->degree == 1 && !M->force_generic)
      MFMGB_CHECK(mf_q1_diagonal(ctx, M, gd_dev, diag_dev));
    else
    {
      MfParams prm = make_params(M);
      mf_diag_kernel<<<(unsigned)ceil_div(M->n, 256), 256, 0, ctx->stream>>>(dim, M->degree, prm, gd_dev, M->n, diag_dev);
      MFMGB_LAUNCHED(ctx);
    }
    MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(gd_dev);
    return MFMGB_OK;
  }
}
