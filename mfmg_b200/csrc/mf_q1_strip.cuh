// mf_q1_strip.cuh -- third form of the constant-coefficient Q1 stencil sweep (mf_q1_sweep.cuh): a thread owns a STRIP of
// R y-adjacent nodes of one x column.
//
// ncu of the node-pair form (profiles/r02_ncu_full_mf_stencil_pairs_v1_raw.csv): the issue slots are 32 % busy, but the L1 /
// shared-memory data pipe is at 76 % -- 36 shared-memory wavefronts per warp and plane (ring write + read 12, y-stage
// store 8 + load 16) plus ~20 for the 16-byte-strided global accesses of the pairs, for 64 nodes.  Here
//   * the lanes of a warp are 32 x-adjacent nodes: every global access (cp.async ring, store) is one contiguous 256-byte
//     row piece, x-neighbours are warp shuffles;
//   * a warp owns R = 4 consecutive rows of the tile, so the y-neighbours of the inner rows are the thread's own
//     registers; only the first and the last row of a strip go through shared memory (one 128-bit store each, two
//     128-bit loads per thread and plane instead of four per node pair);
//   * the rows below / above the tile are computed (x-stage only) by two extra warps.
// => 8 + 8 shared wavefronts of y-stage per 128 nodes instead of 24 per 64, ~0.4 data-pipe wavefronts per node in all
// instead of ~0.9.  Same arithmetic, same evaluation order as the other two forms (bit-identical results).
// A CTA is NS strip warps + 2 halo warps = 256 threads = a 32 x 24 node tile, 30 x 24 emitted.
// (included by mf_q1_sweep.cuh: StencilArgs, sw_plane_offset, cp_async_f64)

constexpr int S3_R = 4;                   // rows per thread
constexpr int S3_NS = 6;                  // strip warps per CTA
constexpr int S3_NW = S3_NS + 2;          // + the halo row below and the halo row above
constexpr int S3_NT = S3_NW * 32;
constexpr int S3_UX = 30, S3_UY = S3_NS * S3_R; // nodes a CTA emits per plane

// STRIP: the warp owns R rows of the tile (all stages); !STRIP: it owns ONE halo row (x-stage only).  Both roles run the
// same number of steps, i.e. the same number of CTA barriers.
template <int EPI, bool ARITH, int RING, bool XIN_IS_X, bool STRIP>
__device__ __forceinline__ void s3_sweep(const StencilArgs &a, const double *__restrict__ x, const EpiArgs &e)
{
  constexpr int NT = S3_NT, R = S3_R, NS = S3_NS;
  constexpr int NR = STRIP ? R : 1; // rows this thread loads
  constexpr int U = (RING % 2 == 0) ? RING : 2 * RING;
  extern __shared__ __align__(16) double sw_smem[];
  // y-stage exchange, double-buffered: HI[buf][s][lane] = (m, d) of the LAST row of strip s - 1 (s = 0: the halo row
  // below the tile), LO[buf][s][lane] = (m, d) of the FIRST row of strip s (s = NS: the halo row above the tile)
  double2(*HI)[NS + 1][32] = reinterpret_cast<double2(*)[NS + 1][32]>(sw_smem);
  double2(*LO)[NS + 1][32] = reinterpret_cast<double2(*)[NS + 1][32]>(sw_smem + 2 * 2 * (NS + 1) * 32);
  double *ring0 = sw_smem + 2 * 2 * 2 * (NS + 1) * 32;
  double(*xr)[R][NT] = reinterpret_cast<double(*)[R][NT]>(ring0);                          // [RING][R][NT]
  double(*br)[R][NT] = reinterpret_cast<double(*)[R][NT]>(ring0 + (size_t)RING * R * NT);
  double(*dr)[R][NT] = reinterpret_cast<double(*)[R][NT]>(ring0 + (size_t)2 * RING * R * NT);
  const int tid = threadIdx.x, lx = tid & 31, w = tid >> 5;
  constexpr bool is_strip = STRIP;
  const int gi = (int)blockIdx.x * S3_UX - 1 + lx;
  const int gj0 = (int)blockIdx.y * S3_UY;
  // rows of this thread: a strip of R rows, or the single halo row below / above the tile
  const int gj_first = is_strip ? gj0 + w * R : (w == NS ? gj0 - 1 : gj0 + S3_UY);
  constexpr int nr = NR;
  const bool col_ok = gi >= 0 && gi < (int)a.nx;
  const int gic = gi < 0 ? 0 : (gi >= (int)a.nx ? (int)a.nx - 1 : gi);
  bool ok[R], emit[R], edge[R];
  int o[R]; // offset of row r relative to the thread's first (clamped) node: addresses stay inside the vector
  const int gjc0 = gj_first < 0 ? 0 : (gj_first >= (int)a.ny ? (int)a.ny - 1 : gj_first);
#pragma unroll
  for (int r = 0; r < R; ++r)
  {
    const int gj = gj_first + r;
    const bool row_ok = r < nr && gj >= 0 && gj < (int)a.ny;
    const int gjc = gj < 0 ? 0 : (gj >= (int)a.ny ? (int)a.ny - 1 : gj);
    ok[r] = row_ok && col_ok;
    emit[r] = is_strip && ok[r] && lx >= 1 && lx <= S3_UX;
    edge[r] = !ok[r] || gi == 0 || gi == (int)a.nx - 1 || gj == 0 || gj == (int)a.ny - 1;
    o[r] = (gjc - gjc0) * (int)a.nx;
  }
  const int64_t P0 = a.g_begin + (int64_t)blockIdx.z * a.seg_planes;
  const int64_t P1 = P0 + a.seg_planes < a.g_end ? P0 + a.seg_planes : a.g_end;
  if (P0 >= P1)
    return;
  const int64_t pl = a.pl;
  const int64_t node_xy = (int64_t)gjc0 * a.nx + gic;
  const int n_steps = (int)(P1 - P0) + 2; // step t handles plane g = P0 - 1 + t; plane g - 1 is emitted at t >= 2
  const bool pl_first = P0 - 1 >= 0, pl_last = P1 < a.nz;
  const int64_t off_first = pl_first ? sw_plane_offset(a, P0 - 1) + node_xy : node_xy;
  const int64_t off_last = pl_last ? sw_plane_offset(a, P1) + node_xy : node_xy;
  const int64_t off_mid = (P0 - a.own0) * pl + node_xy; // plane P0 (step 1); owned planes: vector offset == row
  const int64_t lo_lim = a.bottom_bc ? 0 : -1, hi_lim = a.top_bc ? a.nz - 1 : a.nz;
  const int t_lo = (int)(lo_lim - (P0 - 1) + 1 > 0 ? lo_lim - (P0 - 1) + 1 : 0);
  const int64_t t_hi64 = hi_lim - (P0 - 1);
  const int t_hi = (int)(t_hi64 < n_steps ? t_hi64 : n_steps);
  unsigned f_edge = 0u;
#pragma unroll
  for (int r = 0; r < R; ++r)
    f_edge |= edge[r] ? (1u << r) : 0u;
  constexpr unsigned F_ALL = (1u << R) - 1u;

  // one running offset: roff = (row of the thread's first node) on the plane emitted at the current step
  const int64_t row0 = off_mid;
  int64_t roff = row0 - 2 * pl;
  const int64_t d_x = (int64_t)RING * pl, d_b = (int64_t)(RING - 1) * pl;

  // LEAN (decided per warp, see below): every node of the warp's rows is an unconstrained node inside the box, so there
  // is nothing to predicate, select or mask
  auto request = [&](auto steady, auto lean_t, int tr, int slot) {
    constexpr bool STEADY = decltype(steady)::value, LEAN = decltype(lean_t)::value;
    if (STEADY || tr < n_steps)
    {
      const bool first = !STEADY && tr == 0, last = !STEADY && tr == n_steps - 1;
      const bool pok = first ? pl_first : (last ? pl_last : true);
      const double *src = first ? x + off_first : (last ? x + off_last : x + roff + d_x);
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (r < nr)
          cp_async_f64(&xr[slot][r][tid], src + o[r], LEAN ? true : (pok && ok[r]));
      if (EPI != (int)Epi::Spmv && is_strip && (STEADY || tr >= 2))
      {
#pragma unroll
        for (int r = 0; r < R; ++r)
        {
          cp_async_f64(&br[slot][r][tid], e.b + roff + d_b + o[r], LEAN ? true : emit[r]);
          if (EPI == (int)Epi::Jacobi)
            cp_async_f64(&dr[slot][r][tid], e.dinv + roff + d_b + o[r], LEAN ? true : emit[r]);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto flags_of = [&](int t) -> unsigned { // bit r: node r of the plane of step t reads as a constrained zero
    if (ARITH)
      return (t < t_lo || t >= t_hi) ? F_ALL : f_edge;
    if (t >= n_steps)
      return F_ALL;
    const bool first = t == 0, last = t == n_steps - 1;
    const bool pok = first ? pl_first : (last ? pl_last : true);
    const int64_t off = first ? off_first : (last ? off_last : off_mid + (int64_t)(t - 1) * pl);
    unsigned f = 0u;
#pragma unroll
    for (int r = 0; r < R; ++r)
      f |= ((pok && ok[r]) ? (a.constr[off + o[r]] ? 1u : 0u) : 1u) << r;
    return f;
  };

  roff -= (int64_t)(RING - 1) * pl;
#pragma unroll
  for (int tr = 0; tr < RING - 1; ++tr)
  {
    request(std::false_type{}, std::false_type{}, tr, tr % RING);
    roff += pl;
  }
  unsigned fa = flags_of(0), fb = ARITH ? 0u : flags_of(1), f_prev = F_ALL;
  double up[R], Pm[R], Pc[R], Qm[R], Qc[R];
#pragma unroll
  for (int r = 0; r < R; ++r)
    up[r] = Pm[r] = Pc[r] = Qm[r] = Qc[r] = 0.;

  const bool col_emit = lx >= 1 && lx <= S3_UX;
  auto step = [&](auto steady, auto lean_t, const int t, const int slot, const int buf) {
    constexpr bool STEADY = decltype(steady)::value, LEAN = decltype(lean_t)::value;
    request(steady, lean_t, t + RING - 1, (slot + RING - 1) % RING);
    unsigned fn = 0u;
    if (!ARITH)
      fn = flags_of(t + 2);
    const unsigned f_cur = (ARITH && STEADY) ? f_edge : fa, f_old = (ARITH && STEADY) ? f_edge : f_prev;
    // ---- x stage
    asm volatile("cp.async.wait_group %0;" ::"n"(RING - 1) : "memory");
    double ua[R], m[R], d[R];
#pragma unroll
    for (int r = 0; r < R; ++r)
    {
      if (r < nr)
      {
        ua[r] = xr[slot][r][tid];
        const double u = LEAN ? ua[r] : (((f_cur >> r) & 1u) ? 0. : ua[r]);
        const double uL = __shfl_up_sync(0xffffffffu, u, 1), uR = __shfl_down_sync(0xffffffffu, u, 1);
        const double lr = uL + uR;
        m[r] = fma(4., u, lr);
        d[r] = fma(2., u, -lr);
      }
      else
        ua[r] = m[r] = d[r] = 0.;
    }
    if (is_strip)
    {
      LO[buf][w][lx] = make_double2(m[0], d[0]);
      HI[buf][w + 1][lx] = make_double2(m[R - 1], d[R - 1]);
    }
    else if (w == NS)
      HI[buf][0][lx] = make_double2(m[0], d[0]);
    else
      LO[buf][NS][lx] = make_double2(m[0], d[0]);
    __syncthreads();
    if (is_strip)
    {
      // ---- y stage: inner rows from registers, the rows next to the strip from the neighbouring warps
      const double2 below = HI[buf][w][lx], above = LO[buf][w + 1][lx];
      double Pn[R], Qn[R];
#pragma unroll
      for (int r = 0; r < R; ++r)
      {
        const double m_dn = r > 0 ? m[r - 1] : below.x, m_up = r < R - 1 ? m[r + 1] : above.x;
        const double d_dn = r > 0 ? d[r - 1] : below.y, d_up = r < R - 1 ? d[r + 1] : above.y;
        const double mo = m_dn + m_up, dO = d_dn + d_up;
        Pn[r] = fma(4., m[r], mo);
        Qn[r] = fma(a.cax, fma(4., d[r], dO), a.cay * fma(2., m[r], -mo));
      }
      // ---- z stage: the plane of step t - 1 is complete
      if (STEADY || t >= 2)
      {
#pragma unroll
        for (int r = 0; r < R; ++r)
        {
          const double st = (Qm[r] + fma(4., Qc[r], Qn[r])) + a.caz * fma(2., Pc[r], -(Pm[r] + Pn[r]));
          const double s = LEAN ? st : (((f_old >> r) & 1u) ? up[r] : st); // constrained rows: identity on the raw value
          double out;
          if (EPI == (int)Epi::Spmv)
            out = s;
          else
          {
            const double res = __dsub_rn(s, br[slot][r][tid]);
            if (EPI == (int)Epi::Resid)
              out = res;
            else
            {
              const double tt = __dmul_rn(e.omega, __dmul_rn(dr[slot][r][tid], res)); // omega == 1: exact
              const double xi = XIN_IS_X ? up[r] : ((LEAN ? col_emit : emit[r]) ? e.xin[roff + o[r]] : 0.);
              out = __dsub_rn(xi, tt);
            }
          }
          if (LEAN ? col_emit : emit[r])
            e.y[roff + o[r]] = out;
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r)
      {
        Pm[r] = Pc[r];
        Pc[r] = Pn[r];
        Qm[r] = Qc[r];
        Qc[r] = Qn[r];
        up[r] = ua[r];
      }
    }
    roff += pl;
    if (ARITH)
    {
      if (!STEADY)
      {
        f_prev = fa;
        fa = flags_of(t + 1);
      }
    }
    else
    {
      f_prev = fa;
      fa = fb;
      fb = fn;
    }
  };

  int t = 0;
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (u < n_steps)
    {
      step(std::false_type{}, std::false_type{}, u, u % RING, u & 1);
      t = u + 1;
    }
  if (t == U)
  {
    // warp-uniform: no lane of this warp holds a constrained node or a node outside the box (true for ~3/4 of the
    // warps of a 257^2 plane: all but the tiles on the x faces and the strips on the y faces)
    const bool lean = ARITH && is_strip && __all_sync(0xffffffffu, f_edge == 0u);
    if (lean)
      for (; t + U + RING - 1 <= n_steps - 1; t += U)
      {
#pragma unroll
        for (int u = 0; u < U; ++u)
          step(std::true_type{}, std::true_type{}, t + u, u % RING, u & 1);
      }
    else
      for (; t + U + RING - 1 <= n_steps - 1; t += U)
      {
#pragma unroll
        for (int u = 0; u < U; ++u)
          step(std::true_type{}, std::false_type{}, t + u, u % RING, u & 1);
      }
    if (ARITH)
    {
      f_prev = flags_of(t - 1);
      fa = flags_of(t);
    }
  }
  for (int u = 0; t < n_steps; ++u, ++t)
    step(std::false_type{}, std::false_type{}, t, u % RING, u & 1);
}

template <int EPI, bool ARITH, int RING, bool XIN_IS_X>
__global__ void __launch_bounds__(S3_NT, 2)
    mf_q1_stencil3_kernel(const StencilArgs a, const double *__restrict__ x, const EpiArgs e)
{
  if ((threadIdx.x >> 5) < S3_NS) // warp-uniform
    s3_sweep<EPI, ARITH, RING, XIN_IS_X, true>(a, x, e);
  else
    s3_sweep<EPI, ARITH, RING, XIN_IS_X, false>(a, x, e);
}

template <int EPI>
int launch_q1_stencil3(mfmgb_ctx *ctx, const mfmgb_mf *M, const double *x, const EpiArgs &e, int64_t g0, int64_t g1)
{
  if (g1 <= g0)
    return MFMGB_OK;
  const Q1Params p = make_q1_params(M);
  const int64_t tiles = ceil_div(p.nx, S3_UX) * ceil_div(p.ny, S3_UY);
  static const int env_seg = [] {
    const char *v = getenv("MFMGB_MF_SEGMENTS");
    return v && *v ? atoi(v) : 0;
  }();
  // two resident CTAs per SM; about four waves of CTAs (measured at 257^2 x 257: 6 segments 0.080 ms, 8: 0.072, 12:
  // 0.072 -- many short sweeps balance the SMs better than the two lead-in planes of a segment cost)
  int64_t seg = env_seg > 0 ? env_seg : std::max<int64_t>(1, ((int64_t)ctx->num_sms * 8 + tiles / 2) / tiles);
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
  dim3 grid((unsigned)ceil_div(p.nx, S3_UX), (unsigned)ceil_div(p.ny, S3_UY), (unsigned)seg);
  // ring depth = planes in flight per thread.  (Deeper rings -- 8 / 6 / 4 -- were measured SLOWER: 0.083 vs 0.072 ms per
  // apply, fused forms +30 %: ncu shows barrier and dependency waits, not memory latency, as what is left)
  constexpr int RING = EPI == (int)Epi::Jacobi ? 3 : 4;
  constexpr int n_rings = EPI == (int)Epi::Spmv ? 1 : (EPI == (int)Epi::Resid ? 2 : 3);
  const size_t smem = sizeof(double) * ((size_t)2 * 2 * 2 * (S3_NS + 1) * 32 + (size_t)n_rings * RING * S3_R * S3_NT);
  auto launch = [&](auto kernel) {
    static unsigned long long configured = 0; // one bit per device (per instantiation: the lambda is)
    if (!((configured >> (ctx->device & 63)) & 1ull))
    {
      MFMGB_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(112 * 1024)));
      configured |= 1ull << (ctx->device & 63);
    }
    kernel<<<grid, S3_NT, smem, ctx->stream>>>(a, x, e);
    MFMGB_LAUNCHED(ctx);
    return (int)MFMGB_OK;
  };
  const bool xin_is_x = EPI != (int)Epi::Jacobi || e.xin == x;
  if (M->q1_arith_flags)
    return xin_is_x ? launch(mf_q1_stencil3_kernel<EPI, true, RING, true>)
                    : launch(mf_q1_stencil3_kernel<EPI, true, RING, EPI != (int)Epi::Jacobi>);
  return xin_is_x ? launch(mf_q1_stencil3_kernel<EPI, false, RING, true>)
                  : launch(mf_q1_stencil3_kernel<EPI, false, RING, EPI != (int)Epi::Jacobi>);
}
