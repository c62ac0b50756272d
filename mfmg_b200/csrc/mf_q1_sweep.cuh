// mf_q1_sweep.cuh -- matrix-free 3D Q1 Laplace operator with ONE coefficient for the whole grid: persistent z-sweep
// of the factorised 27-point stencil (the cfg4 fine level with the reference's "constant" material).
//
// Same operator as mf_q1.cuh / mf_laplace.cu (tests/laplace_matrix_free.hpp:121-156 inside deal.II's MatrixFree vmult
// semantics: constrained entries read as 0, constrained rows act as identity).  On a uniform grid with a constant
// coefficient c the sum over the 8 cells around an interior node collapses to
//     A = c [ ax Dx (x) My (x) Mz + ay Mx (x) Dy (x) Mz + az Mx (x) My (x) Dz ],   M = [1 4 1],  D = [-1 2 -1],
//     ax = hy hz / (36 hx), ...   (the node stencils of the unscaled 1D cell matrices [[2,1],[1,2]], [[1,-1],[-1,1]])
// and is evaluated direction by direction -- 17 FP64 operations per node instead of 71 per cell:
//     x:  m = Mx u, d = Dx u                       neighbours in x are lanes of the same warp (shuffles)
//     y:  P = My m,  Q = c ax My d + c ay Dy m     rows j-1, j+1 come from shared memory (one barrier per plane)
//     z:  y_k = (Q_{k-1} + 4 Q_k + Q_{k+1}) + c az (2 P_k - P_{k-1} - P_{k+1})     planes k-1, k live in registers
// A CTA owns a 30 x 14 node tile (32 x 16 threads with a one-node halo ring) and sweeps a contiguous range of z planes.
// x travels global -> shared by 8-byte cp.async into a per-thread ring SW_RING planes deep (rows of 2^k + 1 doubles are
// neither 16-byte aligned nor 16-byte strided, so TMA bulk / tensor copies cannot serve this layout); flags and the
// epilogue operands are requested one / two planes ahead through registers.  There is no per-CTA brick prologue and no
// redundancy in z.  HBM-bound: 8 (x) + 1 (flag) + 8 (y) bytes per node, + 16 for the fused Jacobi sweep.
// Valid when every owned unconstrained node is interior to the local box (all 8 cells around it exist) -- checked at
// creation; otherwise the per-cell kernel of mf_q1.cuh serves the operator.  Fixed evaluation order => bit-reproducible.
// (included by mf_laplace.cu after mf_q1.cuh: Q1Params, plane_offset)

namespace
{
constexpr int SW_TX = 32, SW_TY = 16;           // threads per CTA
constexpr int SW_UX = SW_TX - 2, SW_UY = SW_TY - 2; // nodes a CTA emits per plane
constexpr int SW_RING = 6;                       // x planes in flight per thread (cp.async ring in shared memory)

template <int EPI, int MINB>
__global__ void __launch_bounds__(SW_TX *SW_TY, MINB)
    mf_q1_stencil_kernel(const Q1Params p, const double *__restrict__ x, const EpiArgs e, const int64_t g_begin,
                         const int64_t g_end, const int seg_planes, const double cax, const double cay, const double caz)
{
  __shared__ double sm_m[2][SW_TY][SW_TX], sm_d[2][SW_TY][SW_TX];
  // x of the next SW_RING planes: slot [t % SW_RING][thread] is written by this thread's own cp.async and read by this
  // thread only (x-neighbours travel by shuffle), so the ring needs no barrier -- it is an asynchronous register file
  __shared__ double xr[SW_RING][SW_TX * SW_TY];
  const int tid = threadIdx.x, tx = tid % SW_TX, ty = tid / SW_TX;
  const int64_t gi = (int64_t)blockIdx.x * SW_UX - 1 + tx, gj = (int64_t)blockIdx.y * SW_UY - 1 + ty;
  const bool node_ok = gi >= 0 && gi < p.nx && gj >= 0 && gj < p.ny;
  const bool emit_xy = node_ok && tx >= 1 && tx <= SW_UX && ty >= 1 && ty <= SW_UY;
  const int64_t P0 = g_begin + (int64_t)blockIdx.z * seg_planes;
  const int64_t P1 = P0 + seg_planes < g_end ? P0 + seg_planes : g_end;
  if (P0 >= P1)
    return;
  const int64_t pl = p.nx * p.ny;
  const int64_t node_xy = node_ok ? gj * p.nx + gi : 0;
  const int tyd = ty > 0 ? ty - 1 : 0, tyu = ty < SW_TY - 1 ? ty + 1 : SW_TY - 1;
  const int n_steps = (int)(P1 - P0) + 2; // step t handles plane g = P0 - 1 + t; plane g - 1 is emitted at t >= 2

  // asynchronous copy of this thread's node of the plane of step t into its ring slot (zero-fill outside the box);
  // exactly one commit group per step, so "all but the newest SW_RING - 1 groups done" == the plane of step t landed
  auto request_x = [&](int t) {
    if (t < n_steps)
    {
      const int64_t g = P0 - 1 + t;
      const bool ok = node_ok && g >= 0 && g < p.nz;
      cp_async_f64(&xr[t % SW_RING][tid], x + (ok ? plane_offset(p, g) + node_xy : 0), ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto load_flag = [&](int64_t g) -> unsigned { // nodes outside the box read as constrained zeros
    return node_ok && g >= 0 && g < p.nz ? (unsigned)p.constr[plane_offset(p, g) + node_xy] : 1u;
  };
#pragma unroll
  for (int t = 0; t < SW_RING - 1; ++t)
    request_x(t);
  unsigned fa = load_flag(P0 - 1), fb = load_flag(P0); // flags of planes g and g + 1 of the current step
  // epilogue operands: "b" slot = plane g of the current step, "a" slot = plane g - 1 (the one that is emitted).  The
  // operands of plane P0 enter the pipeline as the "next" of step 0, like every later plane's do one step ahead.
  double ba = 0., bb = 0., da = 0., db = 0.;
  double pend_b = 0., pend_d = 0.;
  if (emit_xy && EPI != (int)Epi::Spmv)
  {
    const int64_t row = (P0 - p.own0) * pl + node_xy;
    pend_b = e.b[row];
    if (EPI == (int)Epi::Jacobi)
      pend_d = e.dinv[row];
  }
  double u_prev = 0.;
  unsigned f_prev = 1u;
  double Pm = 0., Pc = 0., Qm = 0., Qc = 0.;
  for (int t = 0; t < n_steps; ++t)
  {
    const int64_t g = P0 - 1 + t;
    // ---- requests for later steps: x of step t + SW_RING - 1, flag of plane g + 2, epilogue operands of plane g + 1
    request_x(t + SW_RING - 1);
    const unsigned fn = g + 2 <= P1 ? load_flag(g + 2) : 1u;
    double bn = 0., dn = 0.;
    if (t == 0)
    {
      bn = pend_b;
      dn = pend_d;
    }
    else if (emit_xy && EPI != (int)Epi::Spmv && g + 1 < P1)
    {
      const int64_t row = (g + 1 - p.own0) * pl + node_xy;
      bn = e.b[row];
      if (EPI == (int)Epi::Jacobi)
        dn = e.dinv[row];
    }
    // ---- x stage of plane g ----
    asm volatile("cp.async.wait_group %0;" ::"n"(SW_RING - 1) : "memory");
    const double ua = xr[t % SW_RING][tid];
    const double uz = fa ? 0. : ua;
    const double ul = __shfl_up_sync(0xffffffffu, uz, 1), ur = __shfl_down_sync(0xffffffffu, uz, 1);
    const double lr = ul + ur;
    const double m = fma(4., uz, lr), d = fma(2., uz, -lr);
    const int buf = t & 1;
    sm_m[buf][ty][tx] = m;
    sm_d[buf][ty][tx] = d;
    __syncthreads();
    // ---- y stage: in-plane operators of plane g ----
    const double mo = sm_m[buf][tyd][tx] + sm_m[buf][tyu][tx], dO = sm_d[buf][tyd][tx] + sm_d[buf][tyu][tx];
    const double Pn = fma(4., m, mo);
    const double Qn = fma(cax, fma(4., d, dO), cay * fma(2., m, -mo));
    // ---- z stage: plane g - 1 is complete ----
    if (emit_xy && t >= 2)
    {
      const int64_t row = (g - 1 - p.own0) * pl + node_xy;
      const double stencil = (Qm + fma(4., Qc, Qn)) + caz * fma(2., Pc, -(Pm + Pn));
      const double s = f_prev ? u_prev : stencil; // constrained rows act as identity on the raw value
      if (EPI == (int)Epi::Spmv)
        e.y[row] = s;
      else if (EPI == (int)Epi::Resid)
        e.y[row] = __dsub_rn(s, ba);
      else
      {
        const double r = __dsub_rn(s, ba);
        double tt = __dmul_rn(da, r);
        if (e.omega != 1.)
          tt = __dmul_rn(e.omega, tt);
        e.y[row] = __dsub_rn(e.xin == x ? u_prev : e.xin[row], tt);
      }
    }
    // ---- rotate the pipeline ----
    u_prev = ua;
    f_prev = fa;
    fa = fb;
    fb = fn;
    ba = bb;
    da = db;
    bb = bn;
    db = dn;
    Pm = Pc;
    Pc = Pn;
    Qm = Qc;
    Qc = Qn;
  }
}

template <int EPI>
int launch_q1_stencil(mfmgb_ctx *ctx, const mfmgb_mf *M, const double *x, const EpiArgs &e, int64_t g0, int64_t g1)
{
  if (g1 <= g0)
    return MFMGB_OK;
  const Q1Params p = make_q1_params(M);
  const int64_t tiles = ceil_div(p.nx, SW_UX) * ceil_div(p.ny, SW_UY);
  // z segments: enough CTAs for about three per SM, each long enough that its two-plane lead-in is noise
  static const int env_seg = [] {
    const char *v = getenv("MFMGB_MF_SEGMENTS");
    return v && *v ? atoi(v) : 0;
  }();
  // resident CTAs per SM: 3 (40 registers, a few spilled loop invariants) or 2 (64 registers); MFMGB_MF_MINB selects
  static const int env_minb = [] {
    const char *v = getenv("MFMGB_MF_MINB");
    return v && *v ? atoi(v) : 3;
  }();
  int64_t seg = env_seg > 0 ? env_seg : std::max<int64_t>(1, ((int64_t)ctx->num_sms * 3) / tiles);
  seg = std::min<int64_t>(seg, std::max<int64_t>(1, (g1 - g0) / 8));
  const int seg_planes = (int)ceil_div(g1 - g0, seg);
  seg = ceil_div(g1 - g0, seg_planes);
  const double c = M->q1_const_coef;
  dim3 grid((unsigned)ceil_div(p.nx, SW_UX), (unsigned)ceil_div(p.ny, SW_UY), (unsigned)seg);
  if (env_minb == 2)
    mf_q1_stencil_kernel<EPI, 2><<<grid, SW_TX * SW_TY, 0, ctx->stream>>>(p, x, e, g0, g1, seg_planes, c * p.ax, c * p.ay,
                                                                          c * p.az);
  else
    mf_q1_stencil_kernel<EPI, 3><<<grid, SW_TX * SW_TY, 0, ctx->stream>>>(p, x, e, g0, g1, seg_planes, c * p.ax, c * p.ay,
                                                                          c * p.az);
  MFMGB_LAUNCHED(ctx);
  return MFMGB_OK;
}

int dispatch_q1_stencil(mfmgb_ctx *ctx, const mfmgb_mf *M, const double *x, Epi epi, const EpiArgs &e, int64_t g0,
                        int64_t g1)
{
  switch (epi)
  {
  case Epi::Spmv:
    return launch_q1_stencil<(int)Epi::Spmv>(ctx, M, x, e, g0, g1);
  case Epi::Resid:
    return launch_q1_stencil<(int)Epi::Resid>(ctx, M, x, e, g0, g1);
  case Epi::Jacobi:
    return launch_q1_stencil<(int)Epi::Jacobi>(ctx, M, x, e, g0, g1);
  default:
    return fail(ctx, MFMGB_ERR_INVALID, "mf_apply: unsupported epilogue");
  }
}
} // namespace
