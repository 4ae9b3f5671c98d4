// fft_static.cuh -- compile-time specialised version of the shared-memory column transform.
//
// Same algorithm and table layout as fft_tile.cuh (in-place DIF, digit-reversed result, stage
// twiddles tw[off_i + (t-1)*s_i + j]) but the length and radix sequence are template
// parameters: strides, loop trip counts and the index splits are literals, the stage loop is
// unrolled, and the digit reversal is arithmetic instead of a table load.  Used by the kernels
// instantiated for the sizes the configured workloads hit; every other size takes the generic
// (runtime-plan) kernels.
#pragma once
#include "fft_radix.cuh"

namespace kfft {

template <int LEN, int... RS> struct SPlan {
  static constexpr int len = LEN;
  static constexpr int nst = sizeof...(RS);
  static constexpr int rad_arr[sizeof...(RS) > 0 ? sizeof...(RS) : 1] = {RS...};
  static constexpr int rad(int i) { return rad_arr[i]; }
  static constexpr int nsub(int i) {  // length of the sub-transform entering stage i
    int n = LEN;
    for (int k = 0; k < i; k++) n /= rad_arr[k];
    return n;
  }
  static constexpr int stride(int i) { return nsub(i) / rad_arr[i]; }
  static constexpr int tw_off(int i) {
    int o = 0;
    for (int k = 0; k < i; k++)
      if (stride(k) > 1) o += (rad_arr[k] - 1) * stride(k);
    return o;
  }
  static constexpr bool valid() {
    int p = 1;
    for (int k = 0; k < nst; k++) p *= rad_arr[k];
    return p == LEN;
  }
};

// slot holding X[k] after the last stage: digits of k in the mixed radix (r0, r1, ...)
template <class P, int I = 0> struct SlotOf {
  static __device__ __forceinline__ int run(int rem) {
    constexpr int R = P::rad(I), S = P::stride(I);
    int const q = rem / R;
    return (rem - q * R) * S + SlotOf<P, I + 1>::run(q);
  }
};
template <class P> struct SlotOf<P, P::nst> {
  static __device__ __forceinline__ int run(int) { return 0; }
};
template <class P> __device__ __forceinline__ int static_slot(int k) { return SlotOf<P>::run(k); }

template <class P, bool INV, int I>
__device__ __forceinline__ void static_stage(float2 *__restrict__ col, float2 const *__restrict__ tw, int lane) {
  constexpr int R = P::rad(I), NSUB = P::nsub(I), S = P::stride(I), NB = P::len / R;
  constexpr int ITERS = (NB + 31) / 32;
  float2 const *twi = tw + P::tw_off(I);
#pragma unroll
  for (int it = 0; it < ITERS; it++) {
    int const u = lane + 32 * it;
    if ((NB % 32 == 0) || it + 1 < ITERS || u < NB) {
      int const b = (S == 1) ? u : u / S;
      int const j = (S == 1) ? 0 : u - b * S;
      float2 *p = col + b * NSUB + j;
      float2 x[R];
#pragma unroll
      for (int m = 0; m < R; m++) x[m] = p[m * S];
      Dft<R, INV>::run(x);
      if (S > 1) {
#pragma unroll
        for (int t = 1; t < R; t++) {
          float2 const w = __ldg(twi + (t - 1) * S + j);
          x[t] = INV ? cmulc(x[t], w) : cmul(x[t], w);
        }
      }
#pragma unroll
      for (int t = 0; t < R; t++) p[t * S] = x[t];
    }
  }
}

template <class P, bool INV, int I = 0> struct StaticFft {
  static __device__ __forceinline__ void run(float2 *col, float2 const *tw, int lane) {
    static_stage<P, INV, I>(col, tw, lane);
    __syncwarp();
    StaticFft<P, INV, I + 1>::run(col, tw, lane);
  }
};
template <class P, bool INV> struct StaticFft<P, INV, P::nst> {
  static __device__ __forceinline__ void run(float2 *, float2 const *, int) {}
};

// ---- shared-memory bulk copies (TMA, 1-D): cp.async.bulk + mbarrier -------------------------
__device__ __forceinline__ uint32_t smem_u32(void const *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, void const *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

}  // namespace kfft
