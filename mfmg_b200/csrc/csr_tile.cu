// csr_tile.cu -- tile-streamed CSR SpMV family: col/val/rowptr staged into shared memory by TMA bulk copies.
//
// Why: the plain vector-CSR kernel (csr.cu) moves exactly the algorithmic bytes but is latency bound -- every
// byte in flight needs a resident warp waiting on it (ncu, round 1: issue slots 23 % busy, 92 % occupancy, DRAM
// 60 %).  Here the matrix stream does not occupy warps at all:
//   * a tile is RPT = 256 / LPR consecutive rows; its col / val / rowptr ranges are contiguous in the CSR arrays,
//     so ONE elected producer thread moves them with three cp.async.bulk copies (1-D TMA, 16-byte aligned ranges,
//     L2 evict-first: the stream is read once and must not push x out of L2) into a ring of S stages, each
//     guarded by a full/empty mbarrier pair;
//   * 8 consumer warps wait on the full barrier, compute their rows out of shared memory with the same
//     LPR-lanes-per-row mapping and shuffle tree as the vector-CSR kernel (so the x gather of a warp-level load
//     stays within a few 128-byte lines), gather x through the read-only path, apply the fused epilogue and
//     release the stage; the epilogue operands (b, D^-1, x_old) are requested before the wait;
//   * CTAs are persistent (contiguous runs of tiles per CTA, so the j+-1 neighbour lines of x are re-used from L1).
// Bytes in flight per SM = CTAs/SM x S x tile bytes (~170 KB by default) instead of 64 warps x 2 loads x 12 B.
// Summation order depends only on (LPR, row length): deterministic, and graph replay == eager launch.
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "csr.cuh"
#include "peer.cuh"

using namespace mfmgb;

namespace mfmgb
{
namespace
{
constexpr int kConsumerWarps = 8;
constexpr int kTileThreads = (kConsumerWarps + 1) * 32;
constexpr int kUnroll = 8;      // gathers in flight per lane (rows served by <= 4 lanes)
constexpr int kUnrollLong = 16; // ... by >= 8 lanes (long rows: Q2 stencils have up to 125 entries; shared memory limits such
                                // kernels to 2 CTAs per SM, so registers are free and one gather round covers a row)
template <int LPR> constexpr int unroll_of() { return LPR >= 8 ? kUnrollLong : kUnroll; }
template <int LPR> constexpr int min_ctas_of() { return LPR >= 8 ? 2 : 4; }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  uint32_t done;
  do
  {
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                 "selp.u32 %0, 1, 0, p;\n"
                 "}"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  } while (!done);
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               :
               : "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}

template <typename OffT>
struct TileArgs
{
  int64_t n_rows;             // rows of the matrix
  int64_t row_begin, row_end; // rows this launch computes
  int64_t tile_begin, n_tiles;
  // optional second row range [row_begin2, row_end2) in the same launch (the two boundary blocks of a partitioned
  // level): tile slots [n_tiles1, n_tiles) map to tiles tile_begin2 + (slot - n_tiles1)
  int64_t row_begin2, row_end2, tile_begin2, n_tiles1;
  const OffT *rowptr;
  const int *col;
  const double *val;
  const double *x;
  int cap;    // elements per stage (multiple of 4)
  int stages; // ring depth
};

__host__ __device__ constexpr size_t round_up_sz(size_t a, size_t b) { return (a + b - 1) / b * b; }

template <int LPR, typename OffT>
__host__ __device__ constexpr size_t tile_stage_bytes(int cap)
{
  return round_up_sz((size_t)cap * 12 + (size_t)(kConsumerWarps * 32 / LPR + 4) * sizeof(OffT), 128);
}

// dot product of one row's staged entries [k0, ke) (stride LPR) with x; GHOST: columns >= n_owned are read from the
// NVLink mailbox of the current exchange instead of the vector's ghost tail
template <int LPR, bool GHOST, int UNR = unroll_of<LPR>()>
__device__ __forceinline__ double tile_row_sum(const double *__restrict__ sval, const int *__restrict__ scol,
                                               const double *__restrict__ x, int k, const int ke, const GhostArgs &g,
                                               const double *__restrict__ gbox)
{
  double s0 = 0., s1 = 0.;
  // UNR gathers of x in flight per lane (the gather latency under load is what the consumers wait on);
  // even multiples of LPR go to s0, odd ones to s1, ascending: the vector-CSR kernel's summation order
  for (; k < ke; k += UNR * LPR)
  {
    // No predicated memory operation in this block: lanes past the row end re-read the row's last entry and
    // get their product zeroed afterwards.  (With predicated loads ptxas paired every gather with its FMA --
    // one gather in flight per lane; unconditional, the UNR gathers issue back to back.)
    int kk[UNR], c[UNR];
    double v[UNR], xv[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u)
    {
      kk[u] = min(k + u * LPR, ke - 1);
      c[u] = scol[kk[u]];
    }
    if (!GHOST)
    {
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        xv[u] = __ldg(x + c[u]);
    }
    else
    {
      // pointer select first, then unconditional loads (L2: the mailbox lines were written by a peer during this kernel)
      const double *px[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u)
      {
        const long long gc = (long long)c[u] - g.n_owned;
        px[u] = x + c[u];
        if (gc >= 0)
        {
          int l = 0;
          while (l + 1 < g.n_links && gc >= g.links[l + 1].recv_off)
            ++l;
          px[u] = gbox + (size_t)g.links[l].rank * (size_t)g.box_cap + (size_t)(gc - g.links[l].recv_off);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        xv[u] = __ldcg(px[u]);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      v[u] = sval[kk[u]];
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (k + u * LPR >= ke)
        v[u] = 0., xv[u] = 0.;
#pragma unroll
    for (int u = 0; u < UNR; u += 2)
    {
      s0 = fma(v[u], xv[u], s0);
      s1 = fma(v[u + 1], xv[u + 1], s1);
    }
  }
  return s0 + s1;
}

template <int LPR, int EPI, typename OffT, bool GHOST>
__global__ void __launch_bounds__(kTileThreads, min_ctas_of<LPR>()) csr_tile_kernel(const TileArgs<OffT> a, const EpiArgs e, const GhostArgs g)
{
  constexpr int RPT = kConsumerWarps * 32 / LPR; // rows per tile
  constexpr int RP_ELEMS = RPT + 4;              // staged row offsets (a multiple of 4 => 16-byte multiple)
  extern __shared__ __align__(128) unsigned char smem[];
  const int S = a.stages;
  const size_t stage_bytes = tile_stage_bytes<LPR, OffT>(a.cap);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)S * stage_bytes);
  uint64_t *empty = full + S;
  if (threadIdx.x == 0)
  {
    for (int s = 0; s < S; ++s)
    {
      mbar_init(full + s, 1);
      mbar_init(empty + s, kConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (GHOST && (int)blockIdx.x < g.n_push_ctas)
  {
    // ---- fused exchange, push half: the first CTAs store this rank's boundary entries of x into the neighbours'
    // mailboxes over NVLink (every CTA a slice of every send list), the last of them to finish raises the flags.
    // Pushes never wait, and all CTAs of this single-wave grid are resident: the neighbour's kernel, which waits for
    // these flags at its boundary tiles, cannot be starved by anything this rank does.
    const unsigned long long ps = g.seq[0] + 1;
    const size_t pslot = ((size_t)(ps & 1ull) * (size_t)g.nranks + (size_t)g.rank) * (size_t)g.box_cap;
    for (int k = 0; k < g.n_links; ++k)
    {
      const GhostLink l = g.links[k];
      double *dst = reinterpret_cast<double *>(g.base[l.rank] + g.box_off) + pslot;
      const long long chunk = (l.send_cnt + g.n_push_ctas - 1) / g.n_push_ctas;
      const long long lo = (long long)blockIdx.x * chunk, hi = lo + chunk < l.send_cnt ? lo + chunk : l.send_cnt;
      for (long long i = lo + threadIdx.x; i < hi; i += kTileThreads)
        dst[i] = g.send_idx ? a.x[g.send_idx[l.send_off + i]] : a.x[l.send_first + i];
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
      __threadfence_system();
      const unsigned int prev = atomicAdd(&g.done[0], 1u);
      if (prev == (unsigned)g.n_push_ctas - 1u)
      {
        g.done[0] = 0;
        __threadfence_system();
        for (int k = 0; k < g.n_links; ++k)
          st_release_sys(reinterpret_cast<unsigned long long *>(g.base[g.links[k].rank] + g.flag_off) +
                             ((ps & 1ull) * (unsigned long long)g.nranks + (unsigned long long)g.rank),
                         ps);
        g.seq[0] = ps;
      }
    }
  }
  // this CTA's contiguous run of tile slots; slot -> tile (two row ranges may share one launch)
  // Plain mode: CTA b owns the contiguous slots [b tpc, (b+1) tpc) (the j+-1 neighbour lines of x are re-used from L1).
  // Ghost mode (all rows of a partitioned level in one launch): interior tiles first, in contiguous runs as above, then
  // the boundary tiles dealt out round-robin -- at most a few per CTA, at the END of its work.  (With the boundary
  // tiles in their natural place the first and last CTAs held nothing else: they started only when the neighbour's
  // flag arrived and then had a full share of tiles to do, which made the whole launch ~35 us longer.)
  int64_t t_begin, t_end, n_first = 0, bt_lo = 0, bt_hi = 0, a_int0 = 0, bnd0 = 0;
  if (GHOST)
  {
    const int64_t nt = a.n_tiles;                       // tiles 0 .. nt-1 of rows [0, n_rows)
    bt_lo = (g.blo + RPT - 1) / RPT;                    // tiles [0, bt_lo) touch rows below blo
    bt_hi = g.bhi / RPT;                                // tiles [bt_hi, nt) touch rows from bhi on
    if (bt_hi < bt_lo)
      bt_hi = bt_lo;
    const int64_t n_int = bt_hi - bt_lo, n_bnd = nt - n_int;
    const int64_t G = gridDim.x, P = g.n_push_ctas < (int)gridDim.x ? g.n_push_ctas : 0, pen = P > 0 ? g.push_penalty : 0;
    // interior tiles: contiguous runs; the P pushing CTAs get `pen` tiles fewer than the others
    const int64_t tpc = (n_int + P * pen + G - 1) / G, q = tpc > pen ? tpc - pen : 0;
    int64_t i0 = (int64_t)blockIdx.x < P ? (int64_t)blockIdx.x * q : P * q + ((int64_t)blockIdx.x - P) * tpc;
    int64_t i1 = i0 + ((int64_t)blockIdx.x < P ? q : tpc);
    i0 = i0 < n_int ? i0 : n_int;
    i1 = i1 < n_int ? i1 : n_int;
    n_first = i1 - i0;
    // boundary tiles: round-robin over the CTAs, starting behind the pushing ones
    bnd0 = ((int64_t)blockIdx.x - P + G) % G;
    const int64_t mine_bnd = n_bnd > bnd0 ? (n_bnd - 1 - bnd0) / G + 1 : 0;
    t_begin = 0;
    t_end = n_first + mine_bnd; // local slot count
    a_int0 = i0; // first interior tile of this CTA, relative to bt_lo
  }
  else
  {
    const int64_t tpc = (a.n_tiles + gridDim.x - 1) / gridDim.x;
    t_begin = (int64_t)blockIdx.x * tpc;
    t_end = t_begin + tpc < a.n_tiles ? t_begin + tpc : a.n_tiles;
  }
  auto tile_of = [&](int64_t slot) -> int64_t {
    if (GHOST)
    {
      if (slot < n_first)
        return bt_lo + a_int0 + slot;
      const int64_t j = bnd0 + (slot - n_first) * (int64_t)gridDim.x; // j-th boundary tile
      return j < bt_lo ? j : bt_hi + (j - bt_lo);
    }
    return slot < a.n_tiles1 ? a.tile_begin + slot : a.tile_begin2 + (slot - a.n_tiles1);
  };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == kConsumerWarps)
  {
    // ---- producer: one elected thread streams the tiles of this CTA through the ring ----
    if (lane != 0 || t_begin >= t_end)
      return;
    const uint64_t policy = policy_evict_first();
    auto row_off = [&](int64_t r) { return a.rowptr[r < a.n_rows ? r : a.n_rows]; };
    OffT k_lo = row_off(tile_of(t_begin) * RPT), k_hi = row_off((tile_of(t_begin) + 1) * RPT);
    int s = 0;
    uint32_t ph = 0;
    for (int64_t i = t_begin; i < t_end; ++i)
    {
      const int64_t t = tile_of(i);
      // offsets of the next tile, requested one tile ahead of their use
      const int64_t tn = i + 1 < t_end ? tile_of(i + 1) : t;
      const OffT n_lo = row_off(tn * RPT), n_hi = row_off((tn + 1) * RPT);
      mbar_wait(empty + s, ph ^ 1u);
      const OffT a0 = k_lo & ~(OffT)3, a1 = (k_hi + 3) & ~(OffT)3;
      const uint32_t span = (uint32_t)(a1 - a0);
      unsigned char *st = smem + (size_t)s * stage_bytes;
      mbar_expect_tx(full + s, span * 12u + (uint32_t)(RP_ELEMS * sizeof(OffT)));
      bulk_g2s(st + (size_t)a.cap * 12, a.rowptr + t * RPT, (uint32_t)(RP_ELEMS * sizeof(OffT)), full + s, policy);
      if (span)
      {
        bulk_g2s(st, a.val + a0, span * 8u, full + s, policy);
        bulk_g2s(st + (size_t)a.cap * 8, a.col + a0, span * 4u, full + s, policy);
      }
      k_lo = n_lo;
      k_hi = n_hi;
      if (++s == S)
      {
        s = 0;
        ph ^= 1u;
      }
    }
    return;
  }

  // ---- consumers: LPR lanes per row, one row per lane group and tile ----
  const int lr = warp * (32 / LPR) + lane / LPR;
  const int sub = lane % LPR;
  int s = 0;
  uint32_t ph = 0;
  // partitioned level served in one launch: exchange number and mailbox of this exchange (see GhostArgs)
  const unsigned long long xs = GHOST ? g.seq[1] + 1 : 0;
  const double *gbox = GHOST ? reinterpret_cast<const double *>(g.local + g.box_off) +
                                       (size_t)(xs & 1ull) * (size_t)g.nranks * (size_t)g.box_cap
                                 : nullptr;
  bool ghosts_ready = false;
  for (int64_t i = t_begin; i < t_end; ++i)
  {
    const int64_t t = tile_of(i);
    const int64_t row = t * RPT + lr;
    const bool active = (GHOST || i < a.n_tiles1) ? (row >= a.row_begin && row < a.row_end)
                                                  : (row >= a.row_begin2 && row < a.row_end2);
    const bool writer = active && sub == 0;
    // a tile whose rows may reference ghost columns: wait (once per warp) for every neighbour's flag of this exchange
    const bool tile_ghost = GHOST && (t * RPT < g.blo || (t + 1) * RPT > g.bhi);
    if (tile_ghost && !ghosts_ready)
    {
      if (lane < g.n_links)
        wait_flag(reinterpret_cast<const unsigned long long *>(g.local + g.flag_off) +
                      ((size_t)(xs & 1ull) * (size_t)g.nranks + (size_t)g.links[lane].rank),
                  xs, g.timeout_ns, g.err);
      __syncwarp();
      ghosts_ready = true;
    }
    // epilogue operands are requested before the wait so that their latency overlaps it
    double eb = 0., ed = 0., ex = 0.;
    if (writer)
    {
      if (EPI == (int)Epi::Resid || EPI == (int)Epi::Jacobi)
        eb = e.b[row];
      if (EPI == (int)Epi::Jacobi)
      {
        ed = e.dinv[row];
        ex = e.xin[row];
      }
      if (EPI == (int)Epi::Sub)
        ex = e.y[row];
    }
    mbar_wait(full + s, ph);
    const unsigned char *st = smem + (size_t)s * stage_bytes;
    const double *sval = reinterpret_cast<const double *>(st);
    const int *scol = reinterpret_cast<const int *>(st + (size_t)a.cap * 8);
    const OffT *srp = reinterpret_cast<const OffT *>(st + (size_t)a.cap * 12);
    double rs = 0.;
    if (active)
    {
      const OffT base = srp[0] & ~(OffT)3;
      const int k0 = (int)(srp[lr] - base) + sub, ke = (int)(srp[lr + 1] - base);
      // (two separate instantiations: the ghost form must not touch the register allocation and the load scheduling
      // of the form that serves 98 % of the tiles)
      rs = tile_ghost ? tile_row_sum<LPR, true>(sval, scol, a.x, k0, ke, g, gbox)
                      : tile_row_sum<LPR, false>(sval, scol, a.x, k0, ke, g, gbox);
    }
    __syncwarp();
    if (lane == 0)
      mbar_arrive(empty + s); // every lane of this warp has its col/val in registers: the stage may be refilled
    const double sum = subwarp_sum<LPR>(rs);
    if (writer)
    {
      if (EPI == (int)Epi::Spmv)
        e.y[row] = sum;
      else if (EPI == (int)Epi::Resid)
        e.y[row] = __dsub_rn(sum, eb);
      else if (EPI == (int)Epi::Jacobi)
      {
        const double r = __dsub_rn(sum, eb);
        double tt = __dmul_rn(ed, r);
        if (e.omega != 1.)
          tt = __dmul_rn(e.omega, tt);
        e.y[row] = __dsub_rn(ex, tt);
      }
      else
        e.y[row] = __dsub_rn(ex, sum);
    }
    if (++s == S)
    {
      s = 0;
      ph ^= 1u;
    }
  }
  // exchange bookkeeping: the last CTA of the grid advances the count of consumed exchanges (every consumer warp read
  // it before it could finish, so all of them saw the same exchange number); one atomic per CTA, behind a barrier of
  // the consumer warps (the producer warp has left)
  if (GHOST)
  {
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");
    if (threadIdx.x == 0)
    {
      const unsigned int prev = atomicAdd(&g.done[1], 1u);
      if (prev == gridDim.x - 1u)
      {
        g.done[1] = 0;
        g.seq[1] = xs;
      }
    }
  }
}

int env_int(const char *name, int dflt)
{
  const char *v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

template <int LPR, int EPI, typename OffT, bool GHOST>
int launch_tile_kernel(mfmgb_ctx *ctx, const mfmgb_csr *A, const TileArgs<OffT> &a, const EpiArgs &e, const GhostArgs &g,
                       size_t smem)
{
  // per instantiation: opt in to large dynamic shared memory once, and ask how many CTAs are really co-resident
  // (registers can allow fewer than the planned number; the grid must not spill into a second wave)
  // (cached per device: the attribute is a per-device property, and one process may hold contexts on several devices)
  constexpr int kMaxDevices = 64;
  struct PerDevice
  {
    bool attr_set = false;
    size_t occ_smem = 0;
    int occ_ctas = 0;
  };
  static PerDevice cache[kMaxDevices];
  static std::mutex cache_mutex;
  int occ_ctas = 0;
  {
    std::lock_guard<std::mutex> lock(cache_mutex);
    PerDevice &pd = cache[ctx->device % kMaxDevices];
    if (!pd.attr_set)
    {
      MFMGB_CUDA(ctx, cudaFuncSetAttribute(csr_tile_kernel<LPR, EPI, OffT, GHOST>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
      pd.attr_set = true;
    }
    if (pd.occ_smem != smem)
    {
      MFMGB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pd.occ_ctas, csr_tile_kernel<LPR, EPI, OffT, GHOST>,
                                                                    kTileThreads, smem));
      pd.occ_smem = smem;
    }
    occ_ctas = pd.occ_ctas;
  }
  if (occ_ctas < 1)
    return fail(ctx, MFMGB_ERR_CUDA, "csr_tile_kernel: %zu bytes of shared memory do not fit an SM", smem);
  const int64_t grid = std::min<int64_t>(a.n_tiles, (int64_t)ctx->num_sms * std::min(A->tile_ctas, occ_ctas));
  GhostArgs gg = g;
  gg.n_push_ctas = (int)std::min<int64_t>(gg.n_push_ctas, grid);
  csr_tile_kernel<LPR, EPI, OffT, GHOST><<<(unsigned)grid, kTileThreads, smem, ctx->stream>>>(a, e, gg);
  MFMGB_LAUNCHED(ctx);
  return MFMGB_OK;
}

template <int LPR, int EPI, typename OffT>
int launch_tile(mfmgb_ctx *ctx, const mfmgb_csr *A, const double *x, const EpiArgs &e, int64_t r0, int64_t r1,
                int64_t q0, int64_t q1, const GhostArgs &g)
{
  constexpr int RPT = kConsumerWarps * 32 / LPR;
  if (r1 <= r0) // (an empty first range: the second one takes its place)
  {
    r0 = q0;
    r1 = q1;
    q0 = q1 = 0;
  }
  if (r1 <= r0)
    return MFMGB_OK;
  TileArgs<OffT> a;
  a.n_rows = A->n_rows;
  a.row_begin = r0;
  a.row_end = r1;
  a.tile_begin = r0 / RPT;
  a.n_tiles1 = ceil_div(r1, RPT) - a.tile_begin;
  a.row_begin2 = q0;
  a.row_end2 = q1;
  a.tile_begin2 = q1 > q0 ? q0 / RPT : 0;
  a.n_tiles = a.n_tiles1 + (q1 > q0 ? ceil_div(q1, RPT) - a.tile_begin2 : 0);
  a.rowptr = (const OffT *)A->rowptr;
  a.col = A->col;
  a.val = A->val;
  a.x = x;
  a.cap = A->tile_cap[tile_cap_slot(LPR)];
  a.stages = A->tile_stages;
  const size_t smem = (size_t)a.stages * tile_stage_bytes<LPR, OffT>(a.cap) + (size_t)a.stages * 16;
  // per instantiation: opt in to large dynamic shared memory once, and ask how many CTAs are really co-resident
  // (registers can allow fewer than the planned number; the grid must not spill into a second wave)
  if (g.enabled)
    return launch_tile_kernel<LPR, EPI, OffT, true>(ctx, A, a, e, g, smem);
  return launch_tile_kernel<LPR, EPI, OffT, false>(ctx, A, a, e, g, smem);
}

template <int EPI, typename OffT>
int dispatch_lanes(mfmgb_ctx *ctx, const mfmgb_csr *A, const double *x, const EpiArgs &e, int64_t r0, int64_t r1,
                   int64_t q0, int64_t q1, const GhostArgs &g)
{
  switch (A->lanes)
  {
  case 1:
    return launch_tile<1, EPI, OffT>(ctx, A, x, e, r0, r1, q0, q1, g);
  case 2:
    return launch_tile<2, EPI, OffT>(ctx, A, x, e, r0, r1, q0, q1, g);
  case 4:
    return launch_tile<4, EPI, OffT>(ctx, A, x, e, r0, r1, q0, q1, g);
  case 8:
    return launch_tile<8, EPI, OffT>(ctx, A, x, e, r0, r1, q0, q1, g);
  case 16:
    return launch_tile<16, EPI, OffT>(ctx, A, x, e, r0, r1, q0, q1, g);
  default:
    return launch_tile<32, EPI, OffT>(ctx, A, x, e, r0, r1, q0, q1, g);
  }
}

template <typename OffT>
int dispatch_epi(mfmgb_ctx *ctx, const mfmgb_csr *A, const double *x, Epi epi, const EpiArgs &e, int64_t r0, int64_t r1,
                 int64_t q0, int64_t q1, const GhostArgs &g)
{
  switch (epi)
  {
  case Epi::Spmv:
    return dispatch_lanes<(int)Epi::Spmv, OffT>(ctx, A, x, e, r0, r1, q0, q1, g);
  case Epi::Resid:
    return dispatch_lanes<(int)Epi::Resid, OffT>(ctx, A, x, e, r0, r1, q0, q1, g);
  case Epi::Jacobi:
    return dispatch_lanes<(int)Epi::Jacobi, OffT>(ctx, A, x, e, r0, r1, q0, q1, g);
  default:
    return dispatch_lanes<(int)Epi::Sub, OffT>(ctx, A, x, e, r0, r1, q0, q1, g);
  }
}

// widest aligned nnz span of any tile of `rpt` rows (what one stage must hold)
template <typename OffT>
__global__ void __launch_bounds__(256) tile_span_kernel(int64_t n_rows, const OffT *__restrict__ rowptr, int rpt,
                                                        unsigned long long *__restrict__ out_max)
{
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t r0 = t * rpt;
  if (r0 >= n_rows)
    return;
  const int64_t r1 = r0 + rpt < n_rows ? r0 + rpt : n_rows;
  const int64_t span = (((int64_t)rowptr[r1] + 3) & ~(int64_t)3) - ((int64_t)rowptr[r0] & ~(int64_t)3);
  atomicMax(out_max, (unsigned long long)span);
}
} // namespace

int tile_cap_slot(int lanes)
{
  int slot = 0;
  while ((1 << slot) < lanes)
    ++slot;
  return slot;
}

// Decide whether the tile-streamed kernel serves this matrix with its current lanes-per-row, and with which ring.
void csr_plan_tile(mfmgb_csr *A)
{
  A->tile_ok = false;
  if (!A->aligned16 || A->n_rows == 0 || A->nnz == 0 || A->lanes > 32)
    return;
  if (A->tile_rows[tile_cap_slot(A->lanes)] <= 0)
    return;
  const int rpt = kConsumerWarps * 32 / A->lanes;
  const int64_t cap = A->tile_cap[tile_cap_slot(A->lanes)];
  if (cap <= 0 || cap > (1 << 20))
    return;
  const size_t stage = round_up_sz((size_t)cap * 12 + (size_t)(rpt + 4) * (A->off64 ? 8 : 4), 128);
  // Measured on B200 (profiles/r01_tile_sweep.md): what counts is the number of co-resident consumer warps (the x
  // gather is their critical path), a ring of 2 already keeps HBM busy.  Take the largest CTA count <= 4 whose
  // ring still has >= 2 stages (Q1 stencils: 4 CTAs x 2 stages x 21 KB; Q2: 2 CTAs x 2 stages x 49 KB).
  const int want_stages = env_int("MFMGB_TILE_STAGES", 3);
  const int want_ctas = std::max(1, std::min(env_int("MFMGB_TILE_CTAS", 4), 7));
  int stages = 0, ctas = want_ctas;
  for (; ctas >= 1; --ctas)
  {
    const size_t budget = (size_t)224 * 1024 / (size_t)ctas - 1024; // shared memory per CTA (1 KB reserved each)
    stages = (int)std::min<size_t>((size_t)want_stages, budget / (stage + 16));
    if (stages >= 2)
      break;
  }
  if (ctas < 1)
    return;
  A->tile_stages = stages;
  A->tile_ctas = ctas;
  A->tile_ok = true;
}

int csr_measure_tiles(mfmgb_ctx *ctx, mfmgb_csr *A)
{
  for (int s = 0; s < 6; ++s)
    A->tile_cap[s] = A->tile_rows[s] = 0;
  A->aligned16 = ((reinterpret_cast<uintptr_t>(A->val) | reinterpret_cast<uintptr_t>(A->col) |
                   reinterpret_cast<uintptr_t>(A->rowptr)) & 15) == 0;
  if (!A->aligned16 || A->n_rows == 0)
    return MFMGB_OK;
  unsigned long long *dmax = nullptr;
  MFMGB_CUDA(ctx, cudaMalloc(&dmax, sizeof(unsigned long long) * 6));
  MFMGB_CUDA(ctx, cudaMemsetAsync(dmax, 0, sizeof(unsigned long long) * 6, ctx->stream));
  for (int s = 0; s < 6; ++s)
  {
    const int rpt = kConsumerWarps * 32 / (1 << s);
    const int64_t n_tiles = ceil_div(A->n_rows, rpt);
    const unsigned nb = (unsigned)ceil_div(n_tiles, 256);
    if (A->off64)
      tile_span_kernel<int64_t><<<nb, 256, 0, ctx->stream>>>(A->n_rows, (const int64_t *)A->rowptr, rpt, dmax + s);
    else
      tile_span_kernel<int32_t><<<nb, 256, 0, ctx->stream>>>(A->n_rows, (const int32_t *)A->rowptr, rpt, dmax + s);
    MFMGB_LAUNCHED(ctx);
  }
  unsigned long long hmax[6];
  MFMGB_CUDA(ctx, cudaMemcpyAsync(hmax, dmax, sizeof(hmax), cudaMemcpyDeviceToHost, ctx->stream));
  MFMGB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(dmax);
  for (int s = 0; s < 6; ++s)
  {
    A->tile_cap[s] = (int64_t)hmax[s];
    A->tile_rows[s] = A->n_rows;
  }
  if (A->padded)
    return MFMGB_OK;
  // Arrays without slack (the reference's take-ownership constructor): a tile is servable when its staged row
  // offsets [t rpt, t rpt + rpt + 4) lie inside rowptr[0 .. n_rows] and its rounded-up val/col span ends inside nnz.
  for (int s = 0; s < 6; ++s)
  {
    const int rpt = kConsumerWarps * 32 / (1 << s);
    int64_t t = (A->n_rows + 1 - (rpt + 4)) / rpt + 1; // tiles [0, t) keep their row offsets in bounds
    if (A->n_rows + 1 < rpt + 4)
      t = 0;
    while (t > 0)
    {
      int64_t end = 0;
      if (A->off64)
        MFMGB_CUDA(ctx, cudaMemcpy(&end, (const int64_t *)A->rowptr + t * rpt, sizeof(int64_t), cudaMemcpyDeviceToHost));
      else
      {
        int32_t e32 = 0;
        MFMGB_CUDA(ctx, cudaMemcpy(&e32, (const int32_t *)A->rowptr + t * rpt, sizeof(int32_t), cudaMemcpyDeviceToHost));
        end = e32;
      }
      if (((end + 3) & ~(int64_t)3) <= A->nnz)
        break;
      --t;
    }
    A->tile_rows[s] = t * rpt;
  }
  return MFMGB_OK;
}

bool csr_can_fuse_ghost(const mfmgb_csr *A)
{
  return csr_uses_tile_kernel(A) && A->tile_rows[tile_cap_slot(A->lanes)] >= A->n_rows;
}

int csr_apply_tile(mfmgb_ctx *ctx, const mfmgb_csr *A, const double *x, Epi epi, const EpiArgs &args, int64_t row_begin,
                   int64_t row_end, int64_t row_begin2, int64_t row_end2, const GhostArgs *ghost)
{
  const GhostArgs no_ghost;
  const GhostArgs &g = ghost ? *ghost : no_ghost;
  // adopted arrays: rows past the last servable tile go to the direct-load kernel (same summation order)
  const int64_t lim = A->tile_rows[tile_cap_slot(A->lanes)];
  if (row_end > lim || row_end2 > lim)
  {
    const int64_t b1 = std::max(row_begin, lim), b2 = std::max(row_begin2, lim);
    if (row_end > b1)
      MFMGB_CHECK(csr_apply_vec(ctx, A, x, epi, args, b1, row_end));
    if (row_end2 > b2)
      MFMGB_CHECK(csr_apply_vec(ctx, A, x, epi, args, b2, row_end2));
    row_end = std::min(row_end, lim);
    row_end2 = std::min(row_end2, lim);
    if (row_end <= row_begin && row_end2 <= row_begin2)
      return MFMGB_OK;
  }
  if (A->off64)
    return dispatch_epi<int64_t>(ctx, A, x, epi, args, row_begin, row_end, row_begin2, row_end2, g);
  return dispatch_epi<int32_t>(ctx, A, x, epi, args, row_begin, row_end, row_begin2, row_end2, g);
}
} // namespace mfmgb
