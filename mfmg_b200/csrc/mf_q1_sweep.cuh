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
constexpr int SW_RING = 4;                       // x planes in flight per thread (cp.async ring in shared memory)

struct StencilArgs
{
  int64_t nx, ny, nz, pl;     // nodes of the local box, nodes per plane
  int64_t own0, own1, n_owned; // owned planes [own0, own1) (slab layout: [owned | ghost below | ghost above])
  int64_t g_begin, g_end;     // planes this launch emits
  int seg_planes;
  const uint8_t *constr;      // vector layout (read when the flags are not arithmetic)
  int bottom_bc, top_bc;      // arithmetic flags: plane 0 / nz-1 of the local box is a Dirichlet face
  double cax, cay, caz;
};

__device__ __forceinline__ int64_t sw_plane_offset(const StencilArgs &a, int64_t g)
{
  if (g >= a.own0 && g < a.own1)
    return (g - a.own0) * a.pl;
  if (g < a.own0)
    return a.n_owned + g * a.pl;
  return a.n_owned + (a.own0 + (g - a.own1)) * a.pl;
}

// (ncu of the first version, profiles/r02_ncu_full_mf_stencil_v1_raw.csv: issue slots 77 % busy, FP64 pipe and barrier
// stalls, long-scoreboard stalls ~0 -- 236 instructions per warp and plane, most of them 64-bit index arithmetic.  This
// version keeps running pointers: the planes of a segment are contiguous except possibly its first and its last one.)
template <int EPI, bool ARITH, int MINB>
__global__ void __launch_bounds__(SW_TX *SW_TY, MINB)
    mf_q1_stencil_kernel(const StencilArgs a, const double *__restrict__ x, const EpiArgs e)
{
  constexpr int NT = SW_TX * SW_TY;
  extern __shared__ __align__(16) double sw_smem[];
  double(*sm_m)[SW_TY][SW_TX] = reinterpret_cast<double(*)[SW_TY][SW_TX]>(sw_smem);               // [2][TY][TX]
  double(*sm_d)[SW_TY][SW_TX] = reinterpret_cast<double(*)[SW_TY][SW_TX]>(sw_smem + 2 * NT);      // [2][TY][TX]
  // rings SW_RING planes deep: slot [t % SW_RING][thread] is written by this thread's own cp.async and read by this
  // thread only (x-neighbours travel by shuffle), so they need no barrier -- asynchronous register files for x and
  // for the epilogue operands b and D^-1
  double(*xr)[NT] = reinterpret_cast<double(*)[NT]>(sw_smem + 4 * NT);                            // [RING][NT]
  double(*br)[NT] = reinterpret_cast<double(*)[NT]>(sw_smem + (4 + SW_RING) * NT);                // [RING][NT]
  double(*dr)[NT] = reinterpret_cast<double(*)[NT]>(sw_smem + (4 + 2 * SW_RING) * NT);            // [RING][NT]
  const int tid = threadIdx.x, tx = tid % SW_TX, ty = tid / SW_TX;
  const int gi = (int)blockIdx.x * SW_UX - 1 + tx, gj = (int)blockIdx.y * SW_UY - 1 + ty;
  const bool node_ok = gi >= 0 && gi < (int)a.nx && gj >= 0 && gj < (int)a.ny;
  const bool emit_xy = node_ok && tx >= 1 && tx <= SW_UX && ty >= 1 && ty <= SW_UY;
  const int64_t P0 = a.g_begin + (int64_t)blockIdx.z * a.seg_planes;
  const int64_t P1 = P0 + a.seg_planes < a.g_end ? P0 + a.seg_planes : a.g_end;
  if (P0 >= P1)
    return;
  const int64_t pl = a.pl;
  const int64_t node_xy = node_ok ? (int64_t)gj * a.nx + gi : 0;
  const int tyd = ty > 0 ? ty - 1 : 0, tyu = ty < SW_TY - 1 ? ty + 1 : SW_TY - 1;
  const int n_steps = (int)(P1 - P0) + 2; // step t handles plane g = P0 - 1 + t; plane g - 1 is emitted at t >= 2
  // planes P0 .. P1-1 are owned and contiguous; only the first (P0-1) and the last (P1) plane of the segment can be a
  // ghost plane or lie outside the box
  const bool ok_first = node_ok && P0 - 1 >= 0, ok_last = node_ok && P1 < a.nz;
  const int64_t off_first = ok_first ? sw_plane_offset(a, P0 - 1) + node_xy : 0;
  const int64_t off_last = ok_last ? sw_plane_offset(a, P1) + node_xy : 0;
  const int64_t off_mid = (P0 - a.own0) * pl + node_xy; // plane P0 (step 1)
  const bool edge_xy = !node_ok || gi == 0 || gi == (int)a.nx - 1 || gj == 0 || gj == (int)a.ny - 1;

  // asynchronous copy of this thread's node of the plane of step t into its ring slot (zero-fill outside the box);
  // exactly one commit group per step, so "all but the newest SW_RING - 1 groups done" == the plane of step t landed
  const double *xq = x + off_mid - pl; // running pointer: plane of step t at xq + t pl for 1 <= t <= n_steps - 2
  const int64_t row0 = off_mid;        // row of this thread's node on plane P0 (owned planes: vector offset == row)
  // group t = x of step t + the epilogue operands of the plane that step t emits (plane P0 + t - 2)
  auto request_x = [&](int t) {
    if (t < n_steps)
    {
      const bool first = t == 0, last = t == n_steps - 1;
      const bool ok = first ? ok_first : (last ? ok_last : node_ok);
      const double *src = first ? x + off_first : (last ? x + off_last : xq + (int64_t)t * pl);
      cp_async_f64(&xr[t & (SW_RING - 1)][tid], ok ? src : x, ok);
      if (EPI != (int)Epi::Spmv && emit_xy && t >= 2)
      {
        const int64_t row = row0 + (int64_t)(t - 2) * pl;
        cp_async_f64(&br[t & (SW_RING - 1)][tid], e.b + row, true);
        if (EPI == (int)Epi::Jacobi)
          cp_async_f64(&dr[t & (SW_RING - 1)][tid], e.dinv + row, true);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto flag_of = [&](int t) -> unsigned { // constraint flag of this thread's node on the plane of step t
    const int64_t g = P0 - 1 + t;
    if (ARITH)
      return (edge_xy || (g <= 0 && a.bottom_bc) || (g >= a.nz - 1 && a.top_bc) || g < 0 || g >= a.nz) ? 1u : 0u;
    if (t >= n_steps)
      return 1u;
    const bool first = t == 0, last = t == n_steps - 1;
    const bool ok = first ? ok_first : (last ? ok_last : node_ok);
    if (!ok)
      return 1u; // nodes outside the box read as constrained zeros
    return a.constr[first ? off_first : (last ? off_last : off_mid + (int64_t)(t - 1) * pl)];
  };
  static_assert((SW_RING & (SW_RING - 1)) == 0, "the ring depth is a power of two");
#pragma unroll
  for (int t = 0; t < SW_RING - 1; ++t)
    request_x(t);
  unsigned fa = flag_of(0), fb = flag_of(1); // flags of the planes of steps t and t + 1
  double u_prev = 0.;
  unsigned f_prev = 1u;
  double Pm = 0., Pc = 0., Qm = 0., Qc = 0.;
  const int64_t row_emit = row0 - 2 * pl; // row emitted at step t is row_emit + t pl  (plane P0 at t = 2)
  for (int t = 0; t < n_steps; ++t)
  {
    // ---- requests for later steps: x and epilogue operands of step t + SW_RING - 1, flag of step t + 2
    request_x(t + SW_RING - 1);
    const unsigned fn = flag_of(t + 2);
    // ---- x stage of the plane of step t ----
    asm volatile("cp.async.wait_group %0;" ::"n"(SW_RING - 1) : "memory");
    const double ua = xr[t & (SW_RING - 1)][tid];
    const double uz = fa ? 0. : ua;
    const double ul = __shfl_up_sync(0xffffffffu, uz, 1), ur = __shfl_down_sync(0xffffffffu, uz, 1);
    const double lr = ul + ur;
    const double m = fma(4., uz, lr), d = fma(2., uz, -lr);
    const int buf = t & 1;
    sm_m[buf][ty][tx] = m;
    sm_d[buf][ty][tx] = d;
    __syncthreads();
    // ---- y stage: in-plane operators ----
    const double mo = sm_m[buf][tyd][tx] + sm_m[buf][tyu][tx], dO = sm_d[buf][tyd][tx] + sm_d[buf][tyu][tx];
    const double Pn = fma(4., m, mo);
    const double Qn = fma(a.cax, fma(4., d, dO), a.cay * fma(2., m, -mo));
    // ---- z stage: the plane of step t - 1 is complete ----
    if (emit_xy && t >= 2)
    {
      const int64_t row = row_emit + (int64_t)t * pl;
      const double stencil = (Qm + fma(4., Qc, Qn)) + a.caz * fma(2., Pc, -(Pm + Pn));
      const double s = f_prev ? u_prev : stencil; // constrained rows act as identity on the raw value
      if (EPI == (int)Epi::Spmv)
        e.y[row] = s;
      else if (EPI == (int)Epi::Resid)
        e.y[row] = __dsub_rn(s, br[t & (SW_RING - 1)][tid]);
      else
      {
        const double r = __dsub_rn(s, br[t & (SW_RING - 1)][tid]);
        double tt = __dmul_rn(dr[t & (SW_RING - 1)][tid], r);
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
  // z segments (MFMGB_MF_SEGMENTS overrides): enough CTAs to fill the resident slots about twice -- the kernel is
  // issue-bound, a second wave costs nothing and evens out the SMs -- each long enough that its lead-in is noise
  static const int env_seg = [] {
    const char *v = getenv("MFMGB_MF_SEGMENTS");
    return v && *v ? atoi(v) : 0;
  }();
  // resident CTAs per SM: 2 (64 registers, no spills: measured faster) or 3 (40 registers); MFMGB_MF_MINB selects
  static const int env_minb = [] {
    const char *v = getenv("MFMGB_MF_MINB");
    return v && *v ? atoi(v) : 2;
  }();
  int64_t seg = env_seg > 0 ? env_seg : std::max<int64_t>(1, ((int64_t)ctx->num_sms * 7) / tiles);
  seg = std::min<int64_t>(seg, std::max<int64_t>(1, (g1 - g0) / 8));
  const int seg_planes = (int)ceil_div(g1 - g0, seg);
  seg = ceil_div(g1 - g0, seg_planes);
  const double c = M->q1_const_coef;
  StencilArgs a;
  a.nx = p.nx;
  a.ny = p.ny;
  a.nz = p.nz;
  a.pl = p.nx * p.ny;
  a.own0 = p.own0;
  a.own1 = p.own1;
  a.n_owned = p.n_owned;
  a.g_begin = g0;
  a.g_end = g1;
  a.seg_planes = seg_planes;
  a.constr = p.constr;
  a.bottom_bc = M->q1_bottom_bc ? 1 : 0;
  a.top_bc = M->q1_top_bc ? 1 : 0;
  a.cax = c * p.ax;
  a.cay = c * p.ay;
  a.caz = c * p.az;
  dim3 grid((unsigned)ceil_div(p.nx, SW_UX), (unsigned)ceil_div(p.ny, SW_UY), (unsigned)seg);
  const int threads = SW_TX * SW_TY;
  // shared memory: the two stage buffers, the x ring, and the b / D^-1 rings of the fused epilogues
  const int n_rings = EPI == (int)Epi::Spmv ? 1 : (EPI == (int)Epi::Resid ? 2 : 3);
  const size_t smem = sizeof(double) * (size_t)threads * (size_t)(4 + n_rings * SW_RING);
  auto launch = [&](auto kernel) {
    static unsigned long long configured = 0; // one bit per device (per instantiation: the lambda is)
    if (!((configured >> (ctx->device & 63)) & 1ull))
    {
      MFMGB_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(80 * 1024)));
      configured |= 1ull << (ctx->device & 63);
    }
    kernel<<<grid, threads, smem, ctx->stream>>>(a, x, e);
    MFMGB_LAUNCHED(ctx);
    return (int)MFMGB_OK;
  };
  if (M->q1_arith_flags)
    return env_minb == 3 ? launch(mf_q1_stencil_kernel<EPI, true, 3>) : launch(mf_q1_stencil_kernel<EPI, true, 2>);
  return env_minb == 3 ? launch(mf_q1_stencil_kernel<EPI, false, 3>) : launch(mf_q1_stencil_kernel<EPI, false, 2>);
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
