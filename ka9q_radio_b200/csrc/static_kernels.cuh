// static_kernels.cuh -- the three hot kernels specialised at compile time for the transform
// lengths the configured workloads use (fft_static.cuh).  Same arithmetic, same tables and the
// same argument structs as the generic kernels in fwd_kernels.cuh / chan_kernels.cuh, so parity
// tests cover both; what changes is everything around the butterflies:
//   * literal strides / trip counts, unrolled stage loops, arithmetic digit reversal
//   * rows and channel inputs arrive by TMA bulk copies (cp.async.bulk -> mbarrier) straight into
//     shared memory: no register staging, whole rows / slices in flight at once
//   * inter-pass twiddles come from small precomputed tables instead of double-precision sincospi
#pragma once
#include "chan_kernels.cuh"
#include "fft_static.cuh"
#include "fwd_kernels.cuh"

namespace kfft {

struct FwdTables {
  float2 const *twA;    // [n2][nit]  W_nc^{n2*32*it}
  float2 const *twB;    // [n2][32]   W_nc^{n2*r}
  float2 const *rootC;  // [n1/2+1]   W_{2nc}^{k1}   (REAL split only)
  int nit;
};

// ------------------------------------------------------------------ pass 1: columns -----------
template <int FMT, class P>
__global__ void __launch_bounds__(kFwdThreads, 2) fwd_cols_static(Pass1Args const a, FwdTables const tb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *tile = reinterpret_cast<float2 *>(smem_raw);  // [kTile][pitch]
  constexpr int N1 = P::len, RPI = kFwdThreads / kTile /*32*/, NIT = (N1 + RPI - 1) / RPI;
  TilePlan const &pl = c_plans[a.plan];
  int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int const c = tid % kTile, r = tid / kTile;
  int const c0 = blockIdx.x * kTile;
  int const blk = blockIdx.y;
  int const ncols = min(kTile, a.n2 - c0);
  bool const col_ok = c < ncols;
  long const n2g = c0 + c;
  float2 *mycol = tile + c * a.pitch;

  unsigned long long energy = 0;
  unsigned int clips = 0;
  constexpr int U = 8;
  if (col_ok) {
    if (FMT == 1) {
      int const *src = reinterpret_cast<int const *>(a.in) + (long)blk * a.hop + n2g;
#pragma unroll 1
      for (int it0 = 0; it0 < NIT; it0 += U) {
        int w[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          int const n1 = r + RPI * (it0 + u);
          w[u] = (n1 < N1) ? __ldg(src + (long)n1 * a.n2) : 0;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          int const n1 = r + RPI * (it0 + u);
          if (n1 < N1) {
            short lo = (short)(w[u] & 0xffff), hi = (short)((unsigned)w[u] >> 16);
            if (a.derandomize) {
              lo ^= (short)((lo & 1) ? 0xfffe : 0);
              hi ^= (short)((hi & 1) ? 0xfffe : 0);
            }
            if (a.stats && (long)n1 * a.n2 + n2g >= a.first_new) {
              energy += (unsigned long long)((int)lo * lo) + (unsigned long long)((int)hi * hi);
              clips += (lo > 32766 || lo < -32766) + (hi > 32766 || hi < -32766);
            }
            mycol[n1] = make_float2((float)lo * a.scale, (float)hi * a.scale);
          }
        }
      }
    } else {
      float2 const *src = reinterpret_cast<float2 const *>(a.in) + (long)blk * a.hop + n2g;
#pragma unroll 1
      for (int it0 = 0; it0 < NIT; it0 += U) {
        float2 w[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          int const n1 = r + RPI * (it0 + u);
          w[u] = (n1 < N1) ? __ldg(src + (long)n1 * a.n2) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          int const n1 = r + RPI * (it0 + u);
          if (n1 < N1) mycol[n1] = w[u];
        }
      }
    }
  }
  if (FMT == 1 && a.stats) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      energy += __shfl_xor_sync(0xffffffffu, energy, o);
      clips += __shfl_xor_sync(0xffffffffu, clips, o);
    }
    if (lane == 0 && (energy | clips)) {
      atomicAdd(&a.stats[blk].energy, energy);
      atomicAdd(&a.stats[blk].clips, clips);
    }
  }
  __syncthreads();
  if (warp < ncols) StaticFft<P, false>::run(tile + warp * a.pitch, pl.tw, lane);
  __syncthreads();
  if (col_ok) {
    float2 *dst = a.mid + (long)blk * a.nc + n2g;
    float2 const twB = __ldg(tb.twB + n2g * RPI + r);
    float2 const *twA = tb.twA + n2g * tb.nit;
    constexpr int V = 8;
#pragma unroll 1
    for (int it0 = 0; it0 < NIT; it0 += V) {
      float2 v[V];
#pragma unroll
      for (int u = 0; u < V; u++) {
        int const k1 = r + RPI * (it0 + u);
        if (k1 < N1) v[u] = cmul(mycol[static_slot<P>(k1)], cmul(twB, __ldg(twA + it0 + u)));
      }
#pragma unroll
      for (int u = 0; u < V; u++) {
        int const k1 = r + RPI * (it0 + u);
        if (k1 < N1) dst[(long)k1 * a.n2] = v[u];
      }
    }
  }
}

// ------------------------------------------------------------------ pass 2: rows --------------
template <class P, bool REAL_SPLIT>
__global__ void __launch_bounds__(kFwdThreads, 2) fwd_rows_static(Pass2Args const a, FwdTables const tb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *tile = reinterpret_cast<float2 *>(smem_raw);  // [kTile][pitch]
  __shared__ __align__(8) uint64_t bars[kTile];
  constexpr int N2 = P::len;
  static_assert(N2 % 2 == 0, "bulk row copies need 16-byte multiples");
  TilePlan const &pl = c_plans[a.plan];
  int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int const blk = blockIdx.y;
  constexpr int IPC = REAL_SPLIT ? kTile / 2 : kTile;
  RowItem const *items = a.items + (long)blockIdx.x * IPC;
  {
    RowItem const it = items[REAL_SPLIT ? warp >> 1 : warp];
    int row = -1;
    if (REAL_SPLIT) {
      if ((warp & 1) == 0 && it.kind != kRowEmpty) row = it.row_a;
      if ((warp & 1) == 1 && it.kind == kRowPair) row = it.row_b;
    } else if (it.kind == kRowPlain) {
      row = it.row_a;
    }
    float2 *colp = tile + warp * a.pitch;
    if (row >= 0) {
      // one TMA bulk copy brings the whole (contiguous) row; completion lands on this warp's mbarrier
      if (lane == 0) {
        mbar_init(&bars[warp], 1);
        mbar_fence_init();
        mbar_expect_tx(&bars[warp], N2 * 8);
        bulk_g2s(colp, a.mid + (long)blk * a.nc + (long)row * N2, N2 * 8, &bars[warp]);
      }
      __syncwarp();
      mbar_wait(&bars[warp], 0);
      StaticFft<P, false>::run(colp, pl.tw, lane);
    }
  }
  __syncthreads();

  float2 *spec = a.spec + (long)blk * a.spec_stride;
  if (!REAL_SPLIT) {
    int const i = tid % kTile, q0 = tid / kTile;
    RowItem const it = items[i];
    if (it.kind == kRowPlain) {
      float2 const *colp = tile + i * a.pitch;
      constexpr int QS = kFwdThreads / kTile, V = 8;
#pragma unroll 1
      for (int k0 = q0; k0 < N2; k0 += V * QS) {
        float2 v[V];
#pragma unroll
        for (int u = 0; u < V; u++)
          if (k0 + u * QS < N2) v[u] = colp[static_slot<P>(k0 + u * QS)];
#pragma unroll
        for (int u = 0; u < V; u++)
          if (k0 + u * QS < N2) spec[(long)it.row_a + (long)a.n1 * (k0 + u * QS)] = v[u];
      }
    }
    return;
  }
  constexpr int HALF = kTile / 2, QS = kFwdThreads / HALF;
  int const i = tid % HALF, q0 = tid / HALF;
  RowItem const it = items[i];
  if (it.kind == kRowEmpty) return;
  float2 const *ca = tile + (2 * i) * a.pitch;
  float2 const *cb = (it.kind == kRowPair) ? tile + (2 * i + 1) * a.pitch : ca;
  float2 const rootC = __ldg(tb.rootC + it.row_a);
  int const kend = (it.kind == kRowPair) ? N2 : (it.kind == kRowSelf0 ? N2 / 2 + 1 : (N2 + 1) / 2);
  bool const self0 = it.kind == kRowSelf0;
  constexpr int V = 4;
#pragma unroll 1
  for (int k0 = q0; k0 < kend; k0 += V * QS) {
    float2 za[V], zb[V], rd[V];
#pragma unroll
    for (int u = 0; u < V; u++) {
      int const k2 = k0 + u * QS;
      if (k2 < kend) {
        int const k2m = self0 ? (k2 == 0 ? 0 : N2 - k2) : N2 - 1 - k2;
        rd[u] = __ldg(a.rootD + k2);
        za[u] = ca[static_slot<P>(k2)];
        zb[u] = cb[static_slot<P>(k2m)];
      }
    }
#pragma unroll
    for (int u = 0; u < V; u++) {
      int const k2 = k0 + u * QS;
      if (k2 < kend) {
        long const k = (long)it.row_a + (long)a.n1 * k2;
        float2 const w = cmul(rootC, rd[u]);
        float2 const E = make_float2(0.5f * (za[u].x + zb[u].x), 0.5f * (za[u].y - zb[u].y));
        float2 const O = make_float2(0.5f * (za[u].x - zb[u].x), 0.5f * (za[u].y + zb[u].y));
        float2 const Pp = cmul(w, O);
        spec[k] = make_float2(E.x + Pp.y, E.y - Pp.x);
        long const km = a.nc - k;
        if (km != k) spec[km] = make_float2(E.x - Pp.y, -(E.y + Pp.x));
      }
    }
  }
}

// ------------------------------------------------------------------ channels ------------------
// `order` lists the descriptors that share this plan (mixed output rates are launched per plan).
template <class P>
__global__ void __launch_bounds__(kChanWarps * 32) chan_static(ChanArgs const a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[kChanWarps];
  constexpr int NS = P::len, TOP = (NS + 1) / 2;
  static_assert(NS % 2 == 0, "bulk copies need 16-byte multiples");
  constexpr int XS = NS + 4;  // staged slice: up to NS bins + alignment slack
  int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int const oi = blockIdx.x * kChanWarps + warp;
  if (oi >= a.norder) return;
  ChanDesc const d = a.desc[a.order ? a.order[oi] : a.chan_base + oi];
  if (d.plan < 0) return;
  int const blk = blockIdx.y;
  float2 *col = reinterpret_cast<float2 *>(smem_raw) + warp * (NS + XS);
  float2 *xs = col + NS;
  TilePlan const &pl = c_plans[d.plan];
  float2 const *X = a.spec + (long)blk * a.spec_stride;
  float2 const *R = a.resp + d.resp_off;
  float2 *dst = a.out + (long)blk * a.out_stride + d.out_off;

  if (d.ncopy <= 0) {  // nothing of this channel overlaps the master spectrum: zeros (filter.c:823-832)
    for (int i = lane; i < d.olen; i += 32) dst[i] = make_float2(0.f, 0.f);
    return;
  }
  int const qlo = d.dir > 0 ? d.q0 : d.q0 - (d.ncopy - 1);
  bool const wraps = a.wrap && (d.q0 + d.ncopy > a.m_bins);
  int const qa = qlo & ~1;
  if (!wraps) {
    int const qhi = qlo + d.ncopy - 1;
    uint32_t const nx = (uint32_t)(((qhi - qa + 1) + 1) & ~1);
    if (lane == 0) {
      mbar_init(&bars[warp], 1);
      mbar_fence_init();
      mbar_expect_tx(&bars[warp], nx * 8 + NS * 8);
      bulk_g2s(xs, X + qa, nx * 8, &bars[warp]);
      bulk_g2s(col, R, NS * 8, &bars[warp]);
    }
    __syncwarp();
    mbar_wait(&bars[warp], 0);
  } else {  // circular wrap of a COMPLEX master (filter.c:771-772): two pieces, plain loads
    for (int i = lane; i < NS; i += 32) col[i] = __ldg(R + i);
    for (int u = lane; u < d.ncopy; u += 32) {
      int q = d.q0 + u;
      if (q >= a.m_bins) q -= a.m_bins;
      xs[u] = __ldg(X + q);
    }
    __syncwarp();
  }
  // S[wp] = X[q(wp)] * R[wp] in place over the staged response
#pragma unroll 4
  for (int wp = lane; wp < NS; wp += 32) {
    int t = wp - TOP;
    if (t < 0) t += NS;
    int const u = t - d.zlead;
    bool const live = (u >= 0 && u < d.ncopy && wp != TOP);
    int const xi = wraps ? u : (d.q0 + d.dir * u - qa);
    float2 x = xs[live ? xi : 0];
    if (d.dir < 0) x.y = -x.y;
    float2 const v = cmul(x, col[wp]);
    col[wp] = live ? v : make_float2(0.f, 0.f);
  }
  __syncwarp();
  if (d.flags & 1) {
    for (int p = 1 + lane; p < NS / 2; p += 32) {
      float2 const pos = col[p], neg = col[NS - p];
      col[p] = make_float2(pos.x + neg.x, pos.y - neg.y);
      col[NS - p] = make_float2(neg.x - pos.x, neg.y + pos.y);
    }
    if (lane == 0) {
      col[0] = make_float2(0.f, 0.f);
      col[TOP] = make_float2(0.f, 0.f);
    }
    __syncwarp();
  }
  StaticFft<P, true>::run(col, pl.tw, lane);
  int const first = NS - d.olen;
#pragma unroll 4
  for (int i = lane; i < d.olen; i += 32) dst[i] = col[static_slot<P>(first + i)];
}

// does the registry plan have exactly the radices of static plan P?
template <class P> inline bool plan_is(TilePlan const *p) {
  if (p->len != P::len || p->nstages != P::nst) return false;
  for (int i = 0; i < P::nst; i++)
    if (p->radix[i] != P::rad(i)) return false;
  return true;
}

using S1296 = SPlan<1296, 12, 12, 9>;
using S1250 = SPlan<1250, 10, 25, 5>;
using S600 = SPlan<600, 24, 25>;
using S300 = SPlan<300, 20, 15>;
using S1200 = SPlan<1200, 12, 10, 10>;

}  // namespace kfft
