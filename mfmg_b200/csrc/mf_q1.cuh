// mf_q1.cuh -- matrix-free 3D Q1 Laplace/diffusion operator: node-owner z-sweep (the cfg4 fine-level operator).
//
// Same operator as mf_laplace.cu (tests/laplace_matrix_free.hpp:121-156 inside deal.II's MatrixFree vmult semantics:
// constrained entries read as 0, constrained rows act as identity), restructured so that a layer of cells costs ONE
// block barrier and no shared-memory read-modify-write:
//   * thread (tx, ty) of a 32 x 8 tile owns the node column (X0-1+tx, Y0-1+ty) and the cell whose low corner is that
//     node; a CTA owns TZ node planes of its tile and sweeps upwards through the TZ+1 cell layers that touch them.
//   * ONE load phase per CTA: the brick (x of the TZ+2 planes incl. the high-side halo column / row, the per-cell
//     coefficients of the TZ+1 layers) goes global -> shared with 8-byte cp.async, every request in flight at once;
//     constraint flags are batched through registers, constrained entries are zeroed in place.  Bricks without any
//     constrained node (brick_flags, computed once at creation) skip every flag access.
//   * per cell layer L the thread reads the 4 upper nodal values of its cell (the lower 4 are last layer's upper ones),
//     evaluates the cell operator, publishes its 8 local results to shared memory,            __syncthreads()
//     and as NODE owner adds the 4 lower-plane results of the 4 cells around it to the carry of the previous layer --
//     that is plane L, finished: fused epilogue (operands requested before the cell work), one coalesced store -- and
//     keeps the 4 upper-plane results as the new carry.  Fixed order => bit-reproducible, no atomics, no colouring.
//   * coefficient modes: per-cell constant (detected at creation: the (cell, q) table has equal entries per cell,
//     e.g. every cell-wise constant material) -> c (ax D(x)M(x)M + ay M(x)D(x)M + az M(x)M(x)D) u in 71 FP64
//     operations with literal constants; general per-quadrature-point table -> sum factorisation in ~160 (the
//     derivative in direction d does not depend on q_d, so only partial sums of the table enter).
//   * vectors may be the [owned | ghost-below | ghost-above] slab layout of the row-partitioned hierarchy; z chunks can
//     be launched separately (only the first / last chunk read a ghost plane), so the middle ones overlap the exchange.
// History of what did NOT work is in profiles/r01_summary.md section 5 (register pipeline: one exposed DRAM round trip
// per layer; load->store loops: not unrolled by the compiler; reference matrix in the parameter bank: R2UR + IMAD.MOV
// per DFMA).
// Algorithmic bytes per apply: 16 n + 8 n_cells (+ 56 n_cells in the per-q mode) + n (flags).
// (included by mf_laplace.cu after cell_apply<DIM, P>: one translation unit, shared constant tables)

namespace
{
struct Q1Params
{
  int64_t cx, cy, cz; // cells of the local box
  int64_t nx, ny, nz; // nodes = cells + 1
  int64_t own0, own1; // owned node planes [own0, own1) (local plane numbering); the rest are ghost planes
  int64_t n_owned;
  int64_t n_cells;
  const double *coef; // per-cell mode: [n_cells]; per-q mode: [8][n_cells]
  const uint8_t *constr; // vector layout
  const uint8_t *brick_flags; // [grid.z][grid.y][grid.x]: 1 = the CTA's brick contains a constrained node (or NULL)
  int tz;                // owned node planes per CTA
  int zc_begin;          // first z chunk of this launch (blockIdx.z + zc_begin = chunk index)
  double K[64];          // reference cell matrix (per-cell mode), row-major [a][b], Jacobian folded in
  double kx, ky, kz;     // per-q mode: (prod_{e != d} h_e) / h_d
  double ax, ay, az;     // per-cell mode: k_d / 36 (unscaled 1D mass / stiffness matrices)
};

__device__ __forceinline__ int64_t plane_offset(const Q1Params &p, int64_t g)
{
  const int64_t pl = p.nx * p.ny;
  if (g >= p.own0 && g < p.own1)
    return (g - p.own0) * pl;
  if (g < p.own0)
    return p.n_owned + g * pl;
  return p.n_owned + (p.own0 + (g - p.own1)) * pl;
}

// ---- Q1 cell operators with every table entry as a literal: no constant-bank traffic in the inner loop.
// (The first version applied the 8x8 reference matrix from the kernel-parameter bank: ncu showed one R2UR and one
// IMAD.MOV per DFMA, FP64 instructions were 15 % of the issue slots.)
// Local DoF a = x + 2 y + 4 z.  1D Q1 tables on the unit interval: derivative (-1, +1) at both Gauss points, values
// S = (SA, SB) / (SB, SA), weights 1/2; 1D mass matrix (1/6) [[2,1],[1,2]], stiffness [[1,-1],[-1,1]].

// per-cell coefficient: out = c K_ref u with K_ref = ax Dx(x)My(x)Mz + ay Mx(x)Dy(x)Mz + az Mx(x)My(x)Dz (unscaled
// 1D matrices, ax = hy hz / (36 hx) etc. folded into ca[] together with c); 2-point Gauss is exact for these
// integrands, so this is the quadrature result up to rounding.
__device__ __forceinline__ void cell_q1_const(const double *__restrict__ u, const double ca[3], double *__restrict__ out)
{
  double mx[8], dx[4]; // x stage: M~ and D~ along x on the 4 (z, y) lines
#pragma unroll
  for (int l = 0; l < 4; ++l)
  {
    const double u0 = u[2 * l], u1 = u[2 * l + 1];
    mx[2 * l] = fma(2., u0, u1);
    mx[2 * l + 1] = fma(2., u1, u0);
    dx[l] = u1 - u0;
  }
  // y stage.  index helpers: mx[(z*2 + y)*2 + x'], dx[z*2 + y]
  double mdx[4], dy[4], my[8]; // mdx[z*2 + y'], dy[z*2 + x'], my[(z*2 + y')*2 + x']
#pragma unroll
  for (int z = 0; z < 2; ++z)
  {
    mdx[z * 2] = fma(2., dx[z * 2], dx[z * 2 + 1]);
    mdx[z * 2 + 1] = fma(2., dx[z * 2 + 1], dx[z * 2]);
#pragma unroll
    for (int xx = 0; xx < 2; ++xx)
    {
      const double a = mx[(z * 2) * 2 + xx], b = mx[(z * 2 + 1) * 2 + xx];
      dy[z * 2 + xx] = b - a;
      my[(z * 2) * 2 + xx] = fma(2., a, b);
      my[(z * 2 + 1) * 2 + xx] = fma(2., b, a);
    }
  }
  // z stage
  double tx[4], ty[4], dz[4]; // tx[z'*2 + y'], ty[z'*2 + x'], dz[y'*2 + x']
#pragma unroll
  for (int k = 0; k < 2; ++k)
  {
    tx[k] = ca[0] * fma(2., mdx[k], mdx[2 + k]);
    tx[2 + k] = ca[0] * fma(2., mdx[2 + k], mdx[k]);
    ty[k] = ca[1] * fma(2., dy[k], dy[2 + k]);
    ty[2 + k] = ca[1] * fma(2., dy[2 + k], dy[k]);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    dz[k] = ca[2] * (my[4 + k] - my[k]);
#pragma unroll
  for (int z = 0; z < 2; ++z)
#pragma unroll
    for (int y = 0; y < 2; ++y)
#pragma unroll
      for (int xx = 0; xx < 2; ++xx)
      {
        const double a = tx[z * 2 + y], b = ty[z * 2 + xx], c = dz[y * 2 + xx];
        out[(z * 2 + y) * 2 + xx] = ((xx ? a : -a) + (y ? b : -b)) + (z ? c : -c);
      }
}

// per-quadrature-point coefficient c[(qz*2 + qy)*2 + qx] (sum factorisation; the x-derivative does not depend on qx,
// so only the sums of c over the free quadrature index enter)
__device__ __forceinline__ void cell_q1_perq(const double *__restrict__ u, const double *__restrict__ c,
                                             double *__restrict__ out, double kx, double ky, double kz)
{
  constexpr double SA = 0.78867513459481288225, SB = 0.21132486540518711775; // (1 +- 1/sqrt 3) / 2
  // S-combination of a pair: (SA a + SB b, SB a + SA b)
#define MFMGB_S2(a, b, r0, r1)                                                                       \
  {                                                                                                  \
    r0 = fma(SA, (a), SB * (b));                                                                     \
    r1 = fma(SB, (a), SA * (b));                                                                     \
  }
  double gx[4], gy[4], gz[4]; // gx[qz*2 + qy], gy[qz*2 + qx], gz[qy*2 + qx]
  {
    double d[4], t[4]; // d[z*2 + y] = du/dx on the (z, y) line; t[z*2 + qy]
#pragma unroll
    for (int l = 0; l < 4; ++l)
      d[l] = u[2 * l + 1] - u[2 * l];
    MFMGB_S2(d[0], d[1], t[0], t[1]);
    MFMGB_S2(d[2], d[3], t[2], t[3]);
    MFMGB_S2(t[0], t[2], gx[0], gx[2]);
    MFMGB_S2(t[1], t[3], gx[1], gx[3]);
  }
  {
    double d[4], t[4]; // d[z*2 + x] = du/dy; t[z*2 + qx]
#pragma unroll
    for (int z = 0; z < 2; ++z)
#pragma unroll
      for (int xx = 0; xx < 2; ++xx)
        d[z * 2 + xx] = u[(z * 2 + 1) * 2 + xx] - u[(z * 2) * 2 + xx];
    MFMGB_S2(d[0], d[1], t[0], t[1]);
    MFMGB_S2(d[2], d[3], t[2], t[3]);
    MFMGB_S2(t[0], t[2], gy[0], gy[2]);
    MFMGB_S2(t[1], t[3], gy[1], gy[3]);
  }
  {
    double d[4], t[4]; // d[y*2 + x] = du/dz; t[y*2 + qx]
#pragma unroll
    for (int k = 0; k < 4; ++k)
      d[k] = u[4 + k] - u[k];
    MFMGB_S2(d[0], d[1], t[0], t[1]);
    MFMGB_S2(d[2], d[3], t[2], t[3]);
    MFMGB_S2(t[0], t[2], gz[0], gz[2]);
    MFMGB_S2(t[1], t[3], gz[1], gz[3]);
  }
  // fluxes summed over the quadrature index the derivative does not depend on (weights 1/8)
  const double wx = 0.125 * kx, wy = 0.125 * ky, wz = 0.125 * kz;
  double fx[4], fy[4], fz[4];
#pragma unroll
  for (int qz = 0; qz < 2; ++qz)
#pragma unroll
    for (int q = 0; q < 2; ++q)
    {
      fx[qz * 2 + q] = wx * gx[qz * 2 + q] * (c[(qz * 2 + q) * 2] + c[(qz * 2 + q) * 2 + 1]);   // q = qy, sum over qx
      fy[qz * 2 + q] = wy * gy[qz * 2 + q] * (c[(qz * 2) * 2 + q] + c[(qz * 2 + 1) * 2 + q]);   // q = qx, sum over qy
    }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    fz[k] = wz * gz[k] * (c[k] + c[4 + k]); // k = qy*2 + qx, sum over qz
  // transposed S-combinations back to the nodes
  double X[4], Y[4], Z[4], t[4]; // X[z*2 + y], Y[z*2 + x], Z[y*2 + x]
  MFMGB_S2(fx[0], fx[1], t[0], t[1]); // over qy for qz = 0
  MFMGB_S2(fx[2], fx[3], t[2], t[3]);
  MFMGB_S2(t[0], t[2], X[0], X[2]);
  MFMGB_S2(t[1], t[3], X[1], X[3]);
  MFMGB_S2(fy[0], fy[1], t[0], t[1]);
  MFMGB_S2(fy[2], fy[3], t[2], t[3]);
  MFMGB_S2(t[0], t[2], Y[0], Y[2]);
  MFMGB_S2(t[1], t[3], Y[1], Y[3]);
  MFMGB_S2(fz[0], fz[1], t[0], t[1]);
  MFMGB_S2(fz[2], fz[3], t[2], t[3]);
  MFMGB_S2(t[0], t[2], Z[0], Z[2]);
  MFMGB_S2(t[1], t[3], Z[1], Z[3]);
#undef MFMGB_S2
#pragma unroll
  for (int z = 0; z < 2; ++z)
#pragma unroll
    for (int y = 0; y < 2; ++y)
#pragma unroll
      for (int xx = 0; xx < 2; ++xx)
      {
        const double a = X[z * 2 + y], b = Y[z * 2 + xx], g = Z[y * 2 + xx];
        out[(z * 2 + y) * 2 + xx] = ((xx ? a : -a) + (y ? b : -b)) + (z ? g : -g);
      }
}

constexpr int TX = 32;

// shared-memory bytes of one CTA: raw x and flags of tz+2 planes, per-cell coefficients of tz+1 layers (per-cell mode),
// two buffers of 8 local results per cell
template <int TY, bool PERQ>
__host__ __device__ constexpr size_t q1_smem_bytes(int tz)
{
  const size_t XS = (TX + 1) * (TY + 1), NT = TX * TY;
  const size_t xr = (size_t)(tz + 2) * XS * 8, cs = PERQ ? 0 : (size_t)(tz + 1) * NT * 8, os = 2 * 8 * NT * 8;
  const size_t fs = ((size_t)(tz + 2) * XS + 15) / 16 * 16;
  return xr + cs + os + fs;
}

__device__ __forceinline__ void cp_async_f64(double *dst_smem, const double *src, bool valid)
{
  // 8-byte asynchronous copy global -> shared (no register, no scoreboard); !valid zero-fills
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  const int bytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

template <int TY, int TZ, int EPI, bool PERQ, int MINB>
__global__ void __launch_bounds__(TX *TY, MINB) mf_q1_kernel(const Q1Params p, const double *__restrict__ x,
                                                             const EpiArgs e)
{
  constexpr int NT = TX * TY;
  constexpr int XS = (TX + 1) * (TY + 1); // one x plane incl. the high-side halo column / row
  extern __shared__ __align__(16) unsigned char q1_smem[];
  constexpr int tz = TZ;
  double *xr = reinterpret_cast<double *>(q1_smem);                 // [tz+2][XS] raw x of planes P0-1 .. P0+tz
  double *cs = xr + (size_t)(tz + 2) * XS;                           // [tz+1][NT] per-cell coefficients (per-cell mode)
  double *os = cs + (PERQ ? 0 : (size_t)(tz + 1) * NT);              // [2][8][NT] local results of the cells
  uint8_t *fs = reinterpret_cast<uint8_t *>(os + 2 * 8 * NT);       // [tz+2][XS] constraint flags
  const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
  const int64_t gi0 = (int64_t)blockIdx.x * (TX - 1) - 1, gj0 = (int64_t)blockIdx.y * (TY - 1) - 1;
  const int64_t gi = gi0 + tx, gj = gj0 + ty;
  const bool node_ok = gi >= 0 && gi < p.nx && gj >= 0 && gj < p.ny;
  const bool cell_xy = gi >= 0 && gi < p.cx && gj >= 0 && gj < p.cy;
  const bool owner = tx >= 1 && ty >= 1 && node_ok;
  const int64_t node_xy = gj * p.nx + gi;
  const int64_t chunk = (int64_t)blockIdx.z + p.zc_begin;
  const int64_t P0 = p.own0 + chunk * tz;
  const int64_t P1 = P0 + tz < p.own1 ? P0 + tz : p.own1;
  const int64_t L0 = P0 - 1; // first cell layer: contributes only the carry of plane P0
  const int sx = ty * (TX + 1) + tx;

  // ---- one load phase per CTA: the whole brick of x / flags / coefficients, every request in flight at once.
  // x and the per-cell coefficients go global -> shared with 8-byte cp.async (no register dependency); the 1-byte
  // flags are batched through registers.  (A layer-by-layer register pipeline, and then a load->store loop that the
  // compiler cannot unroll, each exposed one DRAM round trip per plane: 0.7 ms at 256^3 instead of ~0.1.)
  // Element r of a plane (r < XS = 297) is loaded by thread r, the last XS - NT of them by the first threads too;
  // everything that does not depend on the plane is computed once.
  const int n_planes = (int)(P1 - L0) + 1; // planes L0 .. P1
  constexpr int NPL = TZ + 2;
  const int r1 = NT + tid;                 // second element of this thread (only tid < XS - NT)
  const bool has1 = r1 < XS;
  const int64_t li0 = gi0 + tid % (TX + 1), lj0 = gj0 + tid / (TX + 1);
  const int64_t li1 = gi0 + r1 % (TX + 1), lj1 = gj0 + r1 / (TX + 1);
  const bool ok0 = li0 >= 0 && li0 < p.nx && lj0 >= 0 && lj0 < p.ny;
  const bool ok1 = has1 && li1 >= 0 && li1 < p.nx && lj1 >= 0 && lj1 < p.ny;
  const int64_t xy0 = ok0 ? lj0 * p.nx + li0 : 0, xy1 = ok1 ? lj1 * p.nx + li1 : 0;
  // most bricks contain no constrained node at all (Dirichlet nodes sit on the boundary): they skip every flag access
  const bool any_c =
      p.brick_flags == nullptr || p.brick_flags[((size_t)chunk * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] != 0;
  uint8_t fl0[NPL], fl1[NPL];
  const double *xp0 = x + xy0, *xp1 = x + xy1;
  const uint8_t *fp0 = p.constr + xy0, *fp1 = p.constr + xy1;
#pragma unroll
  for (int s = 0; s < NPL; ++s)
  {
    const int64_t g = L0 + s;
    const bool pl_ok = s < n_planes && g >= 0 && g < p.nz;
    const int64_t base = pl_ok ? plane_offset(p, g) : 0;
    if (s < n_planes)
      cp_async_f64(xr + s * XS + tid, xp0 + base, pl_ok && ok0);
    fl0[s] = any_c ? (pl_ok && ok0 ? fp0[base] : (uint8_t)1) : (uint8_t)0;
  }
  if (has1) // the XS - NT = 41 extra elements of a plane: only the first two warps take this branch
  {
#pragma unroll
    for (int s = 0; s < NPL; ++s)
    {
      const int64_t g = L0 + s;
      const bool pl_ok = s < n_planes && g >= 0 && g < p.nz;
      const int64_t base = pl_ok ? plane_offset(p, g) : 0;
      if (s < n_planes)
        cp_async_f64(xr + s * XS + r1, xp1 + base, pl_ok && ok1);
      fl1[s] = any_c ? (pl_ok && ok1 ? fp1[base] : (uint8_t)1) : (uint8_t)0;
    }
  }
  if (!PERQ)
  {
    const int64_t cell_xy_off = cell_xy ? gi + p.cx * gj : 0, cell_pl = p.cx * p.cy;
#pragma unroll
    for (int l = 0; l < TZ + 1; ++l)
    {
      const int64_t L = L0 + l;
      const bool ok = l < n_planes - 1 && cell_xy && L >= 0 && L < p.cz;
      cp_async_f64(cs + l * NT + tid, p.coef + (ok ? cell_xy_off + cell_pl * L : 0), ok);
    }
  }
  if (any_c)
  {
#pragma unroll
    for (int s = 0; s < NPL; ++s)
      if (s < n_planes)
        fs[s * XS + tid] = fl0[s];
    if (has1)
    {
#pragma unroll
      for (int s = 0; s < NPL; ++s)
        if (s < n_planes)
          fs[s * XS + r1] = fl1[s];
    }
  }
  auto load_coef_q = [&](int64_t L, double *c) {
    const bool ok = cell_xy && L >= 0 && L < p.cz;
    const int64_t cell = ok ? gi + p.cx * (gj + p.cy * L) : 0;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      c[q] = ok ? ld_stream_f64(p.coef + (int64_t)q * p.n_cells + cell) : 0.;
  };
  double coef[PERQ ? 8 : 1], coef_next[PERQ ? 8 : 1];
  if (PERQ)
    load_coef_q(L0, coef);
  asm volatile("cp.async.wait_all;" ::: "memory");
  // constrained entries read as 0 (the raw value is only needed by the identity row: re-read there).  Each thread
  // fixes the elements it copied itself, which are visible to it after its own wait.
  if (any_c)
  {
#pragma unroll
    for (int s = 0; s < NPL; ++s)
      if (s < n_planes && fl0[s])
        xr[s * XS + tid] = 0.;
  }
  if (any_c && has1)
  {
#pragma unroll
    for (int s = 0; s < NPL; ++s)
      if (s < n_planes && fl1[s])
        xr[s * XS + r1] = 0.;
  }
  __syncthreads();

  auto get4 = [&](int s, double *v) { // u values (constrained -> 0) of this thread's cell corners on brick plane s
    const double *xp = xr + s * XS + sx;
    v[0] = xp[0];
    v[1] = xp[1];
    v[2] = xp[TX + 1];
    v[3] = xp[TX + 2];
  };
  double xl[4], xu[4];
  get4(0, xl);
  double carry = 0.;
  for (int l = 0; l < n_planes - 1; ++l)
  {
    // layer L = L0 + l uses brick planes l (lower) and l + 1 (upper); a layer above the box (L == cz, reached when
    // this CTA owns the top plane) has no cells and only retires the carry
    const int64_t L = L0 + l;
    const int buf = l & 1;
    const bool cell_ok = cell_xy && L >= 0 && L < p.cz;
    if (PERQ)
      load_coef_q(L + 1, coef_next);
    // epilogue operands of plane L: requested before the cell work so that their latency overlaps it
    const bool emit = owner && L >= P0 && L < P1;
    const int64_t row = (L - p.own0) * (p.nx * p.ny) + node_xy; // owned planes: vector offset == row
    double eb = 0., ed = 0.;
    if (emit && EPI != (int)Epi::Spmv)
    {
      eb = e.b[row];
      if (EPI == (int)Epi::Jacobi)
        ed = e.dinv[row];
    }
    double out[8];
    if (cell_ok)
    {
      get4(l + 1, xu);
      double u[8] = {xl[0], xl[1], xl[2], xl[3], xu[0], xu[1], xu[2], xu[3]};
      if (PERQ)
        cell_q1_perq(u, coef, out, p.kx, p.ky, p.kz);
      else
      {
        const double c = cs[l * NT + tid];
        const double ca[3] = {c * p.ax, c * p.ay, c * p.az};
        cell_q1_const(u, ca, out);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        xl[k] = xu[k];
    }
    else
    {
#pragma unroll
      for (int a = 0; a < 8; ++a)
        out[a] = 0.;
      get4(l + 1, xl);
    }
#pragma unroll
    for (int a = 0; a < 8; ++a)
      os[(buf * 8 + a) * NT + tid] = out[a];
    __syncthreads();
    // node owner: plane L is complete
    if (owner)
    {
      const double *o = os + buf * 8 * NT + tid;
      const double lo = (o[0] + o[NT - 1]) + (o[2 * NT - TX] + o[3 * NT - TX - 1]);
      const double up = (o[4 * NT] + o[5 * NT - 1]) + (o[6 * NT - TX] + o[7 * NT - TX - 1]);
      const double total = carry + lo;
      carry = up;
      if (emit)
      {
        const bool constrained = any_c && fs[l * XS + sx] != 0;
        const double xraw = constrained ? x[row] : xr[l * XS + sx];
        const double s = constrained ? xraw : total;
        if (EPI == (int)Epi::Spmv)
          e.y[row] = s;
        else if (EPI == (int)Epi::Resid)
          e.y[row] = __dsub_rn(s, eb);
        else
        {
          const double r = __dsub_rn(s, eb);
          double t = __dmul_rn(ed, r);
          if (e.omega != 1.)
            t = __dmul_rn(e.omega, t);
          e.y[row] = __dsub_rn(e.xin == x ? xraw : e.xin[row], t);
        }
      }
    }
    if (PERQ)
    {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        coef[q] = coef_next[q];
    }
  }
}

// one CTA per brick of the apply grid: does it contain a constrained node?  (setup, once)
template <int TY>
__global__ void __launch_bounds__(256) mf_q1_brick_flags_kernel(const Q1Params p, int tz, uint8_t *__restrict__ out)
{
  const int64_t gi0 = (int64_t)blockIdx.x * (TX - 1) - 1, gj0 = (int64_t)blockIdx.y * (TY - 1) - 1;
  const int64_t P0 = p.own0 + (int64_t)blockIdx.z * tz;
  const int64_t P1 = P0 + tz < p.own1 ? P0 + tz : p.own1;
  const int n_planes = (int)(P1 - (P0 - 1)) + 1;
  constexpr int XS = (TX + 1) * (TY + 1);
  int any = 0;
  for (int i = threadIdx.x; i < n_planes * XS; i += 256)
  {
    const int s = i / XS, r = i % XS;
    const int64_t g = P0 - 1 + s, li = gi0 + r % (TX + 1), lj = gj0 + r / (TX + 1);
    if (g >= 0 && g < p.nz && li >= 0 && li < p.nx && lj >= 0 && lj < p.ny)
      any |= p.constr[plane_offset(p, g) + lj * p.nx + li] != 0;
  }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0)
    out[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = any ? 1 : 0;
}

// diagonal of the operator on the owned nodes (constrained entries := 1), thread per node
template <bool PERQ>
__global__ void __launch_bounds__(256) mf_q1_diag_kernel(const Q1Params p, const double *__restrict__ gdiag,
                                                         double *__restrict__ diag)
{
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (r >= p.n_owned)
    return;
  if (p.constr[r])
  {
    diag[r] = 1.;
    return;
  }
  const int64_t pl = p.nx * p.ny;
  const int64_t gk = p.own0 + r / pl, gj = (r % pl) / p.nx, gi = r % p.nx;
  double s = 0.;
  for (int dz = 0; dz < 2; ++dz)
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx)
      {
        const int64_t ci = gi - dx, cj = gj - dy, ck = gk - dz;
        if (ci < 0 || ci >= p.cx || cj < 0 || cj >= p.cy || ck < 0 || ck >= p.cz)
          continue;
        const int64_t cell = ci + p.cx * (cj + p.cy * ck);
        const int a = dx + 2 * dy + 4 * dz;
        if (PERQ)
        {
          for (int q = 0; q < 8; ++q)
            s = fma(p.coef[(int64_t)q * p.n_cells + cell], gdiag[q * 8 + a], s);
        }
        else
          s = fma(p.coef[cell], p.K[a * 8 + a], s);
      }
  diag[r] = s;
}

Q1Params make_q1_params(const mfmgb_mf *M)
{
  Q1Params p;
  p.cx = M->cells[0];
  p.cy = M->cells[1];
  p.cz = M->cells[2];
  p.nx = M->nodes[0];
  p.ny = M->nodes[1];
  p.nz = M->nodes[2];
  p.own0 = M->own0;
  p.own1 = M->own1;
  p.n_owned = (M->own1 - M->own0) * M->nodes[0] * M->nodes[1];
  p.n_cells = M->n_cells;
  p.coef = M->q1_cell_constant ? M->coef_cell : M->coef;
  p.constr = M->constr;
  p.brick_flags = M->brick_flags;
  p.tz = M->q1_tz;
  p.zc_begin = 0;
  for (int i = 0; i < 64; ++i)
    p.K[i] = M->Kref[i];
  const double *h = M->h;
  p.kx = h[1] * h[2] / h[0];
  p.ky = h[0] * h[2] / h[1];
  p.kz = h[0] * h[1] / h[2];
  p.ax = p.kx / 36.;
  p.ay = p.ky / 36.;
  p.az = p.kz / 36.;
  return p;
}

template <int TY, int TZ, int EPI, bool PERQ, int MINB>
int launch_q1_cfg(mfmgb_ctx *ctx, const mfmgb_mf *M, const double *x, const EpiArgs &e, int zc0, int zc1)
{
  Q1Params p = make_q1_params(M);
  p.tz = TZ;
  const size_t smem = q1_smem_bytes<TY, PERQ>(TZ);
  static unsigned long long configured = 0; // one bit per device: the attribute is a per-device property
  if (!((configured >> (ctx->device & 63)) & 1ull))
  {
    MFMGB_CUDA(ctx, cudaFuncSetAttribute(mf_q1_kernel<TY, TZ, EPI, PERQ, MINB>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
    configured |= 1ull << (ctx->device & 63);
  }
  const int n_chunks = (int)ceil_div(p.own1 - p.own0, (int64_t)TZ);
  if (zc0 < 0) // all chunks
  {
    zc0 = 0;
    zc1 = n_chunks;
  }
  zc1 = zc1 < n_chunks ? zc1 : n_chunks;
  if (zc1 <= zc0)
    return MFMGB_OK;
  p.zc_begin = zc0;
  dim3 grid((unsigned)ceil_div(p.nx, TX - 1), (unsigned)ceil_div(p.ny, TY - 1), (unsigned)(zc1 - zc0));
  mf_q1_kernel<TY, TZ, EPI, PERQ, MINB><<<grid, TX * TY, smem, ctx->stream>>>(p, x, e);
  MFMGB_LAUNCHED(ctx);
  return MFMGB_OK;
}

template <int TY, int EPI, bool PERQ>
int launch_q1(mfmgb_ctx *ctx, const mfmgb_mf *M, const double *x, const EpiArgs &e, int zc0, int zc1)
{
  // owned planes per CTA / CTAs per SM, fixed at creation (mf_q1_prepare): per-q 8 / 2; per-cell 6 / 3 (three 68 KB
  // bricks resident at 80 registers) or 12 / 2 (MFMGB_MF_TZ=12: half the halo redundancy at 128 registers)
  if (PERQ)
    return launch_q1_cfg<TY, 8, EPI, PERQ, 2>(ctx, M, x, e, zc0, zc1);
  if (M->q1_tz == 12)
    return launch_q1_cfg<TY, 12, EPI, PERQ, 2>(ctx, M, x, e, zc0, zc1);
  return launch_q1_cfg<TY, 6, EPI, PERQ, 3>(ctx, M, x, e, zc0, zc1);
}

template <int TY, bool PERQ>
int dispatch_q1(mfmgb_ctx *ctx, const mfmgb_mf *M, const double *x, Epi epi, const EpiArgs &e, int zc0, int zc1)
{
  switch (epi)
  {
  case Epi::Spmv:
    return launch_q1<TY, (int)Epi::Spmv, PERQ>(ctx, M, x, e, zc0, zc1);
  case Epi::Resid:
    return launch_q1<TY, (int)Epi::Resid, PERQ>(ctx, M, x, e, zc0, zc1);
  case Epi::Jacobi:
    return launch_q1<TY, (int)Epi::Jacobi, PERQ>(ctx, M, x, e, zc0, zc1);
  default:
    return fail(ctx, MFMGB_ERR_INVALID, "mf_apply: unsupported epilogue");
  }
}
} // namespace

#include "mf_q1_sweep.cuh"

namespace mfmgb
{
int mf_q1_num_chunks(const mfmgb_mf *M);
// z chunks [zc0, zc1) of the owned planes (zc0 < 0: all).  Only the first chunk reads the ghost plane below and only
// the last one the ghost plane above, so a partitioned level runs the middle chunks while the halo is exchanged.
int mf_q1_apply(mfmgb_ctx *ctx, const mfmgb_mf *M, const double *x, Epi epi, const EpiArgs &args, int zc0, int zc1)
{
  if (M->q1_stencil)
  {
    // plane ranges of the sweep kernel: chunk 0 = the first owned plane (the only one that reads the ghost plane
    // below), chunk 2 = the last one (ghost plane above), chunk 1 = everything in between
    const int nc = mf_q1_num_chunks(M);
    if (zc0 < 0 || nc == 1)
      return zc0 <= 0 ? dispatch_q1_stencil(ctx, M, x, epi, args, M->own0, M->own1) : MFMGB_OK;
    const int64_t cut[4] = {M->own0, M->own0 + 1, M->own1 - 1, M->own1};
    zc1 = zc1 < nc ? zc1 : nc;
    if (zc1 <= zc0)
      return MFMGB_OK;
    return dispatch_q1_stencil(ctx, M, x, epi, args, cut[zc0], cut[zc1]);
  }
  if (M->q1_cell_constant)
    return dispatch_q1<8, false>(ctx, M, x, epi, args, zc0, zc1);
  return dispatch_q1<8, true>(ctx, M, x, epi, args, zc0, zc1);
}

int mf_q1_num_chunks(const mfmgb_mf *M)
{
  if (M->q1_stencil)
    return M->own1 - M->own0 >= 3 ? 3 : 1;
  return (int)ceil_div(M->own1 - M->own0, (int64_t)M->q1_tz);
}

// fixes the brick height and marks the bricks that contain constrained nodes (called once by the create functions)
int mf_q1_prepare(mfmgb_ctx *ctx, mfmgb_mf *M)
{
  const char *v = getenv("MFMGB_MF_TZ");
  const int env_tz = v && *v ? atoi(v) : 0;
  M->q1_tz = M->q1_cell_constant ? (env_tz == 12 ? 12 : 6) : 8;
  M->brick_flags = nullptr;
  Q1Params p = make_q1_params(M);
  dim3 grid((unsigned)ceil_div(p.nx, TX - 1), (unsigned)ceil_div(p.ny, 8 - 1),
            (unsigned)ceil_div(p.own1 - p.own0, (int64_t)M->q1_tz));
  uint8_t *flags = nullptr;
  MFMGB_CUDA(ctx, cudaMalloc(&flags, (size_t)grid.x * grid.y * grid.z));
  mf_q1_brick_flags_kernel<8><<<grid, 256, 0, ctx->stream>>>(p, M->q1_tz, flags);
  MFMGB_LAUNCHED(ctx);
  MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  M->brick_flags = flags;
  return MFMGB_OK;
}

int mf_q1_diagonal(mfmgb_ctx *ctx, const mfmgb_mf *M, const double *gdiag_dev, double *diag_dev)
{
  const Q1Params p = make_q1_params(M);
  const unsigned nb = (unsigned)ceil_div(p.n_owned, 256);
  if (M->q1_cell_constant)
    mf_q1_diag_kernel<false><<<nb, 256, 0, ctx->stream>>>(p, gdiag_dev, diag_dev);
  else
    mf_q1_diag_kernel<true><<<nb, 256, 0, ctx->stream>>>(p, gdiag_dev, diag_dev);
  MFMGB_LAUNCHED(ctx);
  return MFMGB_OK;
}

} // namespace mfmgb
