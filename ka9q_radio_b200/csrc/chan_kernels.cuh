// chan_kernels.cuh -- the per-channel half of the fast convolver, batched: for every channel
// (slave) and block, gather its bins from the master spectrum, multiply by the channel's
// frequency response, run the small inverse transform and keep the last olen samples.
// Replaces execute_filter_output's arithmetic (reference filter.c:728-921): the four slicing
// loops (:728-893), ISB (:895-909), Nyquist zero (:911), fftwf_execute(rev_plan) (:914) and the
// "output = buffer + points - olen" discard (:357).
//
// One warp owns one (channel, block): the 600-point NBFM case is two fat in-register stages
// (radix 24 then 25) with only __syncwarp between them; hundreds of channels x several blocks
// go out as one grid.
#pragma once
#include "fft_tile.cuh"

namespace kfft {

constexpr int kChanWarps = 4;  // (channel, block) pairs per CTA

// Host-resolved description of how output bin t (in the reference's walk order, starting at the
// most negative output bin) maps onto the master spectrum.  Covers filter.c:810-893 (REAL
// master, upright or inverted) and :728-793 (COMPLEX master with circular wrap).
struct ChanDesc {
  int plan;        // registry index of the length-`points` inverse plan, < 0: channel disabled
  int points;      // Ns
  int olen;        // Ls
  int zlead;       // walk positions t < zlead are zero
  int ncopy;       // then ncopy bins are taken from the master ...
  int q0;          // ... starting at master bin q0 ...
  int dir;         // ... stepping +1 or -1 (inverted spectrum => conjugate, filter.c:876)
  int flags;       // bit0: ISB
  long resp_off;   // float2 offset of this channel's response
  long out_off;    // float2 offset of this channel's output inside a block's output row
};

struct ChanArgs {
  float2 const *spec;
  long spec_stride;
  int m_bins;        // master bins (wrap modulus for COMPLEX masters)
  int wrap;          // 1: COMPLEX master (q wraps mod m_bins), 0: REAL
  ChanDesc const *desc;
  int const *order;  // descriptor indices to process (one launch per plan), or nullptr:
  int norder;        //   then descriptors chan_base .. chan_base+norder-1
  int chan_base;
  float2 const *resp;
  float2 *out;
  long out_stride;
  int pitch;         // shared-memory floats2 per warp
};

__global__ void __launch_bounds__(kChanWarps * 32) chan_kernel(ChanArgs const a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int const oi = blockIdx.x * kChanWarps + warp;
  if (oi >= a.norder) return;
  ChanDesc const d = a.desc[a.order ? a.order[oi] : a.chan_base + oi];
  if (d.plan < 0) return;
  int const blk = blockIdx.y;
  float2 *col = reinterpret_cast<float2 *>(smem_raw) + warp * a.pitch;
  TilePlan const &pl = c_plans[d.plan];
  int const ns = d.points;
  int const top = (ns + 1) / 2;  // index of the most negative output bin == Nyquist slot

  float2 const *X = a.spec + (long)blk * a.spec_stride;
  float2 const *R = a.resp + d.resp_off;
  auto src_of = [&](int wp, bool &live, bool &cj) -> int {
    int t = wp - top;
    if (t < 0) t += ns;
    int const u = t - d.zlead;
    live = (u >= 0 && u < d.ncopy && wp != top);  // Nyquist slot is forced to zero (filter.c:911)
    cj = d.dir < 0;
    int q = d.q0 + d.dir * u;
    if (a.wrap && q >= a.m_bins) q -= a.m_bins;
    return live ? q : 0;
  };
  constexpr int U = 4;
  int wp = lane;
  for (; wp + (U - 1) * 32 < ns; wp += U * 32) {
    float2 x[U], rr[U];
    bool live[U], cj[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      int const q = src_of(wp + u * 32, live[u], cj[u]);
      x[u] = __ldg(X + q);
      rr[u] = __ldg(R + wp + u * 32);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (cj[u]) x[u].y = -x[u].y;
      float2 const v = cmul(x[u], rr[u]);
      col[wp + u * 32] = live[u] ? v : make_float2(0.f, 0.f);
    }
  }
  for (; wp < ns; wp += 32) {
    bool live, cj;
    int const q = src_of(wp, live, cj);
    float2 x = __ldg(X + q);
    if (cj) x.y = -x.y;
    float2 const v = cmul(x, __ldg(R + wp));
    col[wp] = live ? v : make_float2(0.f, 0.f);
  }
  __syncwarp();
  if (d.flags & 1) {  // ISB: (S[p], S[ns-p]) <- (S[p]+conj S[ns-p], S[ns-p]-conj S[p]); S[0]=0
    for (int p = 1 + lane; p < ns / 2; p += 32) {
      float2 const pos = col[p], neg = col[ns - p];
      col[p] = make_float2(pos.x + neg.x, pos.y - neg.y);
      col[ns - p] = make_float2(neg.x - pos.x, neg.y + pos.y);
    }
    if (lane == 0) {
      col[0] = make_float2(0.f, 0.f);
      col[top] = make_float2(0.f, 0.f);
    }
    __syncwarp();
  }
  tile_fft<true>(pl, col, lane, 32, [] { __syncwarp(); });
  float2 *dst = a.out + (long)blk * a.out_stride + d.out_off;
  int const first = ns - d.olen;
  {
    constexpr int V = 4;
    int i = lane;
    for (; i + (V - 1) * 32 < d.olen; i += V * 32) {
      int slot[V];
      float2 v[V];
#pragma unroll
      for (int u = 0; u < V; u++) slot[u] = __ldg(pl.perm + first + i + u * 32);
#pragma unroll
      for (int u = 0; u < V; u++) v[u] = col[slot[u]];
#pragma unroll
      for (int u = 0; u < V; u++) dst[i + u * 32] = v[u];
    }
    for (; i < d.olen; i += 32) dst[i] = col[__ldg(pl.perm + first + i)];
  }
}

// Forward transform of one response in place (set_filter's fftwf_execute, filter.c:1030):
// one warp, data staged through shared memory.
__global__ void __launch_bounds__(32) response_fft_kernel(float2 *resp, int plan) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *col = reinterpret_cast<float2 *>(smem_raw);
  TilePlan const &pl = c_plans[plan];
  int const lane = threadIdx.x;
  for (int i = lane; i < pl.len; i += 32) col[i] = resp[i];
  __syncwarp();
  tile_fft<false>(pl, col, lane, 32, [] { __syncwarp(); });
  for (int k = lane; k < pl.len; k += 32) resp[k] = col[pl.perm[k]];
}

}  // namespace kfft
