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

// ---------------------------------------------------------------------------------------------------------------
// Second form of the sweep (default): TWO x-adjacent nodes per thread.  ncu of the one-node form at HEAD
// (profiles/r02_ncu_full_mf_stencil_head_raw.csv): still issue-bound -- 105 instructions per warp and plane for 32 nodes,
// ~20 of them FP64; the rest is per-thread bookkeeping (ring slots, pointers, predicates, register rotation, 4 SHFL, 2 STS,
// 4 LDS, barrier).  With a node pair per thread that bookkeeping is paid once per TWO nodes, the x-neighbours of the pair
// cost the same 4 SHFL, and the y-stage goes through shared memory in 128-bit pieces.  The plane loop is unrolled by
// lcm(RING, 2) so that ring slots and stage buffers are compile-time constants and the register pipeline needs no moves.
// A CTA is 16 x 32 threads = a 32 x 32 node tile with a one-node halo ring, 30 x 30 emitted (12 % halo instead of 18 %).
// ---------------------------------------------------------------------------------------------------------------
constexpr int S2_LX = 16, S2_TY = 24;                   // threads per CTA: lanes in x (a node pair each) x rows
constexpr int S2_UX = 2 * S2_LX - 2, S2_UY = S2_TY - 2; // nodes a CTA emits per plane
constexpr int S2_NT = S2_LX * S2_TY;

__device__ __forceinline__ void cp_async_f64x2(double2 *dst_smem, const double *src0, bool ok0, bool ok1, int fix0, int fix1)
{
  // two 8-byte copies (rows of 2^k + 1 doubles are not 16-byte aligned); !ok zero-fills (source size 0).  The address
  // of a node outside the box is replaced by its partner's (fix0 / fix1), so every address handed to the copy unit
  // lies inside the vector whatever the allocation around it looks like
  cp_async_f64(&dst_smem->x, src0 + fix0, ok0);
  cp_async_f64(&dst_smem->y, src0 + 1 - fix1, ok1);
}

template <int EPI, bool ARITH, int RING, bool XIN_IS_X>
__global__ void __launch_bounds__(S2_NT, 2)
    mf_q1_stencil2_kernel(const StencilArgs a, const double *__restrict__ x, const EpiArgs e)
{
  constexpr int NT = S2_NT;
  constexpr int U = (RING % 2 == 0) ? RING : 2 * RING; // unroll: ring slot and stage buffer are constants in the body
  extern __shared__ __align__(16) double sw_smem[];
  double2 *sm2 = reinterpret_cast<double2 *>(sw_smem);
  double2(*sm_m)[S2_TY][S2_LX] = reinterpret_cast<double2(*)[S2_TY][S2_LX]>(sm2);          // [2][TY][LX]
  double2(*sm_d)[S2_TY][S2_LX] = reinterpret_cast<double2(*)[S2_TY][S2_LX]>(sm2 + 2 * NT); // [2][TY][LX]
  double2(*xr)[NT] = reinterpret_cast<double2(*)[NT]>(sm2 + 4 * NT);                        // [RING][NT]
  double2(*br)[NT] = reinterpret_cast<double2(*)[NT]>(sm2 + (4 + RING) * NT);               // [RING][NT]
  double2(*dr)[NT] = reinterpret_cast<double2(*)[NT]>(sm2 + (4 + 2 * RING) * NT);           // [RING][NT]
  const int tid = threadIdx.x, lx = tid % S2_LX, ty = tid / S2_LX;
  const int gi0 = (int)blockIdx.x * S2_UX - 1 + 2 * lx, gi1 = gi0 + 1, gj = (int)blockIdx.y * S2_UY - 1 + ty;
  const bool row_ok = gj >= 0 && gj < (int)a.ny;
  const bool ok0 = row_ok && gi0 >= 0 && gi0 < (int)a.nx, ok1 = row_ok && gi1 >= 0 && gi1 < (int)a.nx;
  const bool row_emit_ok = ty >= 1 && ty <= S2_UY;
  const bool emit0 = ok0 && row_emit_ok && lx >= 1, emit1 = ok1 && row_emit_ok && lx <= S2_LX - 2;
  const int64_t P0 = a.g_begin + (int64_t)blockIdx.z * a.seg_planes;
  const int64_t P1 = P0 + a.seg_planes < a.g_end ? P0 + a.seg_planes : a.g_end;
  if (P0 >= P1)
    return;
  const int64_t pl = a.pl;
  // offset of node 0 in its plane; a thread with no node inside the box works on column 0 (all its copies zero-fill,
  // nothing is emitted), one with a single node inside redirects the other node's addresses to it
  const int64_t node_xy = (ok0 || ok1) ? (int64_t)gj * a.nx + gi0 : 0;
  const int fix0 = (!ok0 && ok1) ? 1 : 0, fix1 = (ok0 && !ok1) ? 1 : 0;
  const int tyd = ty > 0 ? ty - 1 : 0, tyu = ty < S2_TY - 1 ? ty + 1 : S2_TY - 1;
  const int n_steps = (int)(P1 - P0) + 2; // step t handles plane g = P0 - 1 + t; plane g - 1 is emitted at t >= 2
  const bool pl_first = P0 - 1 >= 0, pl_last = P1 < a.nz;
  const int64_t off_first = pl_first ? sw_plane_offset(a, P0 - 1) + node_xy : 0;
  const int64_t off_last = pl_last ? sw_plane_offset(a, P1) + node_xy : 0;
  const int64_t off_mid = (P0 - a.own0) * pl + node_xy; // plane P0 (step 1); owned planes: vector offset == row
  const bool edge0 = !ok0 || gi0 == 0 || gi0 == (int)a.nx - 1 || gj == 0 || gj == (int)a.ny - 1;
  const bool edge1 = !ok1 || gi1 == 0 || gi1 == (int)a.nx - 1 || gj == 0 || gj == (int)a.ny - 1;
  // arithmetic flags: the steps whose plane is a Dirichlet face or lies outside the box are the first t_lo and those
  // from t_hi on
  const int64_t lo_lim = a.bottom_bc ? 0 : -1, hi_lim = a.top_bc ? a.nz - 1 : a.nz;
  const int t_lo = (int)(lo_lim - (P0 - 1) + 1 > 0 ? lo_lim - (P0 - 1) + 1 : 0);
  const int64_t t_hi64 = hi_lim - (P0 - 1);
  const int t_hi = (int)(t_hi64 < n_steps ? t_hi64 : n_steps);

  // ONE running offset per thread: roff = row emitted at the current step (plane P0 + t - 2); everything else is a base
  // pointer from the parameter bank plus roff plus a multiple of the plane size (five running pointers cost ten registers
  // and pushed the fused forms into spills)
  const int64_t row0 = off_mid; // row of node 0 on plane P0
  int64_t roff = row0 - 2 * pl;
  const int64_t d_x = (int64_t)RING * pl, d_b = (int64_t)(RING - 1) * pl; // request of step t + RING - 1 relative to roff
  constexpr bool xin_is_x = XIN_IS_X; // the smoother updates x in place of its own input (out-of-place otherwise)

  // request of step tr (x of its plane + the epilogue operands of the plane step tr emits) into ring slot `slot`.
  // STEADY: tr is neither the first nor the last step of the segment (no special planes, no range check)
  auto request = [&](auto steady, int tr, int slot) {
    constexpr bool STEADY = decltype(steady)::value;
    if (STEADY || tr < n_steps)
    {
      const bool first = !STEADY && tr == 0, last = !STEADY && tr == n_steps - 1;
      const bool pok = first ? pl_first : (last ? pl_last : true);
      // plane of step tr: vector offset off_mid + (tr - 1) pl = roff + RING pl when issued at step t = tr - (RING - 1)
      const double *src = first ? x + off_first : (last ? x + off_last : x + roff + d_x);
      cp_async_f64x2(&xr[slot][tid], src, pok && ok0, pok && ok1, fix0, fix1);
      if (EPI != (int)Epi::Spmv && (STEADY || tr >= 2))
      {
        cp_async_f64x2(&br[slot][tid], e.b + roff + d_b, emit0, emit1, fix0, fix1);
        if (EPI == (int)Epi::Jacobi)
          cp_async_f64x2(&dr[slot][tid], e.dinv + roff + d_b, emit0, emit1, fix0, fix1);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const unsigned f_edge = (edge0 ? 1u : 0u) | (edge1 ? 2u : 0u);
  auto flags_of = [&](int t) -> unsigned { // bit 0 / 1: node 0 / 1 of the plane of step t reads as a constrained zero
    if (ARITH)
      return (t < t_lo || t >= t_hi) ? 3u : f_edge;
    if (t >= n_steps)
      return 3u;
    const bool first = t == 0, last = t == n_steps - 1;
    const bool pok = first ? pl_first : (last ? pl_last : true);
    const int64_t off = first ? off_first : (last ? off_last : off_mid + (int64_t)(t - 1) * pl);
    const unsigned f0 = (pok && ok0) ? a.constr[off] : 1u, f1 = (pok && ok1) ? a.constr[off + 1] : 1u;
    return (f0 ? 1u : 0u) | (f1 ? 2u : 0u);
  };

  // prologue: requests of steps 0 .. RING-2 (roff is positioned for step t = tr - (RING - 1), so step it through)
  roff -= (int64_t)(RING - 1) * pl;
#pragma unroll
  for (int tr = 0; tr < RING - 1; ++tr)
  {
    request(std::false_type{}, tr, tr % RING);
    roff += pl;
  }
  unsigned fa = flags_of(0), fb = ARITH ? 0u : flags_of(1);
  double up0 = 0., up1 = 0.; // raw values of the previous plane (identity rows)
  unsigned f_prev = 3u;
  double Pm0 = 0., Pc0 = 0., Qm0 = 0., Qc0 = 0., Pm1 = 0., Pc1 = 0., Qm1 = 0., Qc1 = 0.;

  // one plane.  STEADY (the bulk of a segment): t >= 2, the planes of steps t - 1 and t carry no z flag (ARITH), and the
  // request it issues is an ordinary plane
  auto step = [&](auto steady, const int t, const int slot, const int buf) {
    constexpr bool STEADY = decltype(steady)::value;
    // ---- requests for later steps
    request(steady, t + RING - 1, (slot + RING - 1) % RING);
    unsigned fn = 0u;
    if (!ARITH)
      fn = flags_of(t + 2);
    const unsigned f_cur = (ARITH && STEADY) ? f_edge : fa, f_old = (ARITH && STEADY) ? f_edge : f_prev;
    // ---- x stage of the plane of step t
    asm volatile("cp.async.wait_group %0;" ::"n"(RING - 1) : "memory");
    const double2 ua = xr[slot][tid];
    const double u0 = (f_cur & 1u) ? 0. : ua.x, u1 = (f_cur & 2u) ? 0. : ua.y;
    const double uL = __shfl_up_sync(0xffffffffu, u1, 1, S2_LX), uR = __shfl_down_sync(0xffffffffu, u0, 1, S2_LX);
    const double lr0 = uL + u1, lr1 = u0 + uR;
    const double m0 = fma(4., u0, lr0), d0 = fma(2., u0, -lr0);
    const double m1 = fma(4., u1, lr1), d1 = fma(2., u1, -lr1);
    sm_m[buf][ty][lx] = make_double2(m0, m1);
    sm_d[buf][ty][lx] = make_double2(d0, d1);
    __syncthreads();
    // ---- y stage
    const double2 ma = sm_m[buf][tyd][lx], mb = sm_m[buf][tyu][lx], da = sm_d[buf][tyd][lx], db = sm_d[buf][tyu][lx];
    const double mo0 = ma.x + mb.x, mo1 = ma.y + mb.y, do0 = da.x + db.x, do1 = da.y + db.y;
    const double Pn0 = fma(4., m0, mo0), Pn1 = fma(4., m1, mo1);
    const double Qn0 = fma(a.cax, fma(4., d0, do0), a.cay * fma(2., m0, -mo0));
    const double Qn1 = fma(a.cax, fma(4., d1, do1), a.cay * fma(2., m1, -mo1));
    // ---- z stage: the plane of step t - 1 is complete
    if (STEADY || t >= 2)
    {
      const double st0 = (Qm0 + fma(4., Qc0, Qn0)) + a.caz * fma(2., Pc0, -(Pm0 + Pn0));
      const double st1 = (Qm1 + fma(4., Qc1, Qn1)) + a.caz * fma(2., Pc1, -(Pm1 + Pn1));
      const double s0 = (f_old & 1u) ? up0 : st0, s1 = (f_old & 2u) ? up1 : st1; // constrained rows: identity
      double o0, o1;
      if (EPI == (int)Epi::Spmv)
      {
        o0 = s0;
        o1 = s1;
      }
      else
      {
        const double2 bv = br[slot][tid];
        const double r0 = __dsub_rn(s0, bv.x), r1 = __dsub_rn(s1, bv.y);
        if (EPI == (int)Epi::Resid)
        {
          o0 = r0;
          o1 = r1;
        }
        else
        {
          const double2 dv = dr[slot][tid];
          // (omega == 1: the product is exact, so the unconditional multiply gives the bits of the form without it)
          const double t0 = __dmul_rn(e.omega, __dmul_rn(dv.x, r0)), t1 = __dmul_rn(e.omega, __dmul_rn(dv.y, r1));
          const double xi0 = xin_is_x ? up0 : (emit0 ? e.xin[roff] : 0.), xi1 = xin_is_x ? up1 : (emit1 ? e.xin[roff + 1] : 0.);
          o0 = __dsub_rn(xi0, t0);
          o1 = __dsub_rn(xi1, t1);
        }
      }
      if (emit0)
        e.y[roff] = o0;
      if (emit1)
        e.y[roff + 1] = o1;
    }
    roff += pl;
    // ---- rotate the pipeline
    up0 = ua.x;
    up1 = ua.y;
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
    Pm0 = Pc0;
    Pc0 = Pn0;
    Qm0 = Qc0;
    Qc0 = Qn0;
    Pm1 = Pc1;
    Pc1 = Pn1;
    Qm1 = Qc1;
    Qc1 = Qn1;
  };

  // lead-in: the first U steps in general form (plane flags at the bottom of the segment, nothing emitted before t = 2)
  int t = 0;
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (u < n_steps)
    {
      step(std::false_type{}, u, u % RING, u & 1);
      t = u + 1;
    }
  // bulk: t is a multiple of U, every step of the block and every request it issues is an ordinary plane
  if (t == U)
  {
    for (; t + U + RING - 1 <= n_steps - 1; t += U)
    {
#pragma unroll
      for (int u = 0; u < U; ++u)
        step(std::true_type{}, t + u, u % RING, u & 1);
    }
    if (ARITH)
    {
      f_prev = flags_of(t - 1);
      fa = flags_of(t);
    }
  }
  // tail in general form (t is a multiple of U here or the segment is already done)
  for (int u = 0; t < n_steps; ++u, ++t)
    step(std::false_type{}, t, u % RING, u & 1);
}

template <int EPI>
int launch_q1_stencil2(mfmgb_ctx *ctx, const mfmgb_mf *M, const double *x, const EpiArgs &e, int64_t g0, int64_t g1)
{
  if (g1 <= g0)
    return MFMGB_OK;
  const Q1Params p = make_q1_params(M);
  const int64_t tiles = ceil_div(p.nx, S2_UX) * ceil_div(p.ny, S2_UY);
  static const int env_seg = [] {
    const char *v = getenv("MFMGB_MF_SEGMENTS");
    return v && *v ? atoi(v) : 0;
  }();
  // two resident CTAs per SM; about two waves of CTAs, each segment long enough that its two lead-in planes are noise
  int64_t seg = env_seg > 0 ? env_seg : std::max<int64_t>(1, ((int64_t)ctx->num_sms * 4 + tiles / 2) / tiles);
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
  dim3 grid((unsigned)ceil_div(p.nx, S2_UX), (unsigned)ceil_div(p.ny, S2_UY), (unsigned)seg);
  constexpr int RING = EPI == (int)Epi::Jacobi ? 3 : 4; // 104 KB / 96 KB / 64 KB of shared memory: two CTAs per SM
  constexpr int n_rings = EPI == (int)Epi::Spmv ? 1 : (EPI == (int)Epi::Resid ? 2 : 3);
  const size_t smem = sizeof(double2) * (size_t)S2_NT * (size_t)(4 + n_rings * RING);
  auto launch = [&](auto kernel) {
    static unsigned long long configured = 0; // one bit per device (per instantiation: the lambda is)
    if (!((configured >> (ctx->device & 63)) & 1ull))
    {
      MFMGB_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(112 * 1024)));
      configured |= 1ull << (ctx->device & 63);
    }
    kernel<<<grid, S2_NT, smem, ctx->stream>>>(a, x, e);
    MFMGB_LAUNCHED(ctx);
    return (int)MFMGB_OK;
  };
  const bool xin_is_x = EPI != (int)Epi::Jacobi || e.xin == x;
  if (M->q1_arith_flags)
    return xin_is_x ? launch(mf_q1_stencil2_kernel<EPI, true, RING, true>)
                    : launch(mf_q1_stencil2_kernel<EPI, true, RING, EPI != (int)Epi::Jacobi>);
  return xin_is_x ? launch(mf_q1_stencil2_kernel<EPI, false, RING, true>)
                  : launch(mf_q1_stencil2_kernel<EPI, false, RING, EPI != (int)Epi::Jacobi>);
}

#include "mf_q1_strip.cuh"

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
  // MFMGB_MF_STENCIL_FORM=1 / 2: the one-node-per-thread / node-pair forms (measurement aids); default 3: strips
  static const int form = [] {
    const char *v = getenv("MFMGB_MF_STENCIL_FORM");
    return v && *v ? atoi(v) : 3;
  }();
  if (form == 3)
    switch (epi)
    {
    case Epi::Spmv:
      return launch_q1_stencil3<(int)Epi::Spmv>(ctx, M, x, e, g0, g1);
    case Epi::Resid:
      return launch_q1_stencil3<(int)Epi::Resid>(ctx, M, x, e, g0, g1);
    case Epi::Jacobi:
      return launch_q1_stencil3<(int)Epi::Jacobi>(ctx, M, x, e, g0, g1);
    default:
      return fail(ctx, MFMGB_ERR_INVALID, "mf_apply: unsupported epilogue");
    }
  if (form == 2)
    switch (epi)
    {
    case Epi::Spmv:
      return launch_q1_stencil2<(int)Epi::Spmv>(ctx, M, x, e, g0, g1);
    case Epi::Resid:
      return launch_q1_stencil2<(int)Epi::Resid>(ctx, M, x, e, g0, g1);
    case Epi::Jacobi:
      return launch_q1_stencil2<(int)Epi::Jacobi>(ctx, M, x, e, g0, g1);
    default:
      return fail(ctx, MFMGB_ERR_INVALID, "mf_apply: unsupported epilogue");
    }
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
