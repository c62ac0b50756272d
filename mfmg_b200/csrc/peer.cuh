// peer.cuh -- device-side primitives of the NVLink peer-memory transport (comm.cu, coarse_dd.cu).
#pragma once
#include "comm.cuh"

namespace mfmgb
{
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// spin until *flag >= want; gives up after timeout_ns and reports through *err (the host checks it at sync points)
__device__ __forceinline__ void wait_flag(const unsigned long long *flag, unsigned long long want,
                                          unsigned long long timeout_ns, int *err)
{
  if (ld_acquire_sys(flag) >= want)
    return;
  const unsigned long long t0 = global_timer_ns();
  while (ld_acquire_sys(flag) < want)
  {
    __nanosleep(40);
    if (global_timer_ns() - t0 > timeout_ns)
    {
      *err = 1;
      __threadfence_system();
      return;
    }
  }
}


struct PeerAllreduceArgs
{
  int nranks, rank;
  unsigned char *const *base; // mapped windows of all ranks
  unsigned char *local;       // this rank's window
  size_t ar_off, flag_off;
  int cap;
  unsigned long long *seq;
  unsigned long long timeout_ns;
  int *err;
};

inline PeerAllreduceArgs peer_allreduce_args(const mfmgb_comm *c)
{
  const mfmgb_peer &p = c->peer;
  return {c->nranks, c->rank, p.base_dev, p.local, p.ar_off, p.ar_flag_off, p.ar_cap, p.ar_seq, p.timeout_ns, p.err_dev};
}

// All-reduce (sum) of buf[0 .. n), n <= cap, executed by ONE CTA (all of its threads call this; buf must already be
// visible to the whole CTA): store the local values into this rank's slot of EVERY rank's window (remote stores over
// NVLink), fence, raise the per-source flags, wait for every rank's flag, sum the slots in rank order.  Slots are
// double-buffered by the parity of the sequence number kept in device memory (graph replay == eager launch); a rank
// cannot be two reductions ahead of another one, because each reduction needs everybody's contribution.
__device__ __forceinline__ void peer_allreduce_cta(double *__restrict__ buf, int n, const PeerAllreduceArgs &a)
{
  const unsigned long long s = a.seq[0] + 1;
  const size_t par = (size_t)(s & 1ull);
  const size_t my_slot = (par * (size_t)a.nranks + (size_t)a.rank) * (size_t)a.cap;
  for (int r = 0; r < a.nranks; ++r)
  {
    double *dst = reinterpret_cast<double *>(a.base[r] + a.ar_off) + my_slot;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      dst[i] = buf[i];
  }
  // No fence per thread: the barrier orders every thread's remote stores before the release stores of the
  // flag-writing lanes, and a release is cumulative over what happens-before it.  (A system fence per thread cost
  // ~20 us here: 32 warps each waiting for their NVLink write acknowledgements one after the other.)
  __syncthreads();
  if ((int)threadIdx.x < a.nranks)
  {
    st_release_sys(reinterpret_cast<unsigned long long *>(a.base[threadIdx.x] + a.flag_off) +
                       (par * (size_t)a.nranks + (size_t)a.rank),
                   s);
    wait_flag(reinterpret_cast<const unsigned long long *>(a.local + a.flag_off) + (par * (size_t)a.nranks + threadIdx.x), s,
              a.timeout_ns, a.err);
  }
  __syncthreads();
  const double *slots = reinterpret_cast<const double *>(a.local + a.ar_off) + par * (size_t)a.nranks * (size_t)a.cap;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
  {
    double acc = 0.;
    for (int r = 0; r < a.nranks; ++r) // fixed rank order: the same bits on every rank
      acc += __ldcg(slots + (size_t)r * (size_t)a.cap + i);
    buf[i] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0)
    a.seq[0] = s;
}
} // namespace mfmgb
