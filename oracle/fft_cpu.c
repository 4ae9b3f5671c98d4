/* oracle/fft_cpu.c -- CPU mixed-radix FFT behind the oracle and the fftw3.h shim.
 * TEST INFRASTRUCTURE, NOT PRODUCT (see fft_cpu.h for the contract being restated).
 */
#define _GNU_SOURCE 1
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "fft_cpu.h"

#define KFFT_LEAF_MAX 8192
#define KFFT_MAXRAD 40
#define KFFT_MAXDEPTH 4

struct kfft_plan {
  int n;
  int depth;             /* recursion depth: selects thread-local scratch slots */
  /* leaf */
  int nrad;
  int rad[KFFT_MAXRAD];
  float complex *twf_f, *twb_f;   /* W_n^t forward / backward, t < n */
  double complex *twf_d, *twb_d;
  /* four-step split (n1 == 0 for a leaf) */
  int n1, n2;
  struct kfft_plan *sub1, *sub2;
  int big_S;
  double complex *big_hi, *big_lo, *big_hi_b, *big_lo_b; /* W_n^t = hi[t/S]*lo[t%S] */
  /* tables for real transforms of length 2n built on this plan: W_{2n}^k, k <= n/2 ... */
  double complex *rtw; /* W_{2n}^k for k = 0..n (forward sign) */
};

/* ---- thread-local growing scratch ------------------------------------------------------- */
#define KFFT_NSLOT (3 * (KFFT_MAXDEPTH + 1) + 2)
static __thread void *Scratch[KFFT_NSLOT];
static __thread size_t Scratch_size[KFFT_NSLOT];
static void *kfft_scratch(int slot, size_t bytes) {
  if (Scratch_size[slot] < bytes) {
    free(Scratch[slot]);
    void *p = NULL;
    if (posix_memalign(&p, 64, bytes + 64) != 0)
      abort();
    Scratch[slot] = p;
    Scratch_size[slot] = bytes;
  }
  return Scratch[slot];
}

static double complex unit_root(long t, long n, int sign) {
  /* exp(sign*2*pi*i*t/n) with exact octant reduction in integers */
  t %= n;
  if (t < 0)
    t += n;
  long double const a = 2.0L * M_PIl * (long double)t / (long double)n;
  return CMPLX((double)cosl(a), (double)(sign * sinl(a)));
}

static int factor_radices(int n, int *rad) {
  int k = 0;
  while (n % 4 == 0) {
    rad[k++] = 4;
    n /= 4;
  }
  if (n % 2 == 0) {
    rad[k++] = 2;
    n /= 2;
  }
  for (int f = 3; f <= 5; f += 2)
    while (n % f == 0) {
      rad[k++] = f;
      n /= f;
    }
  for (int f = 7; (long)f * f <= n; f += 2)
    while (n % f == 0) {
      rad[k++] = f;
      n /= f;
    }
  if (n > 1)
    rad[k++] = n;
  return k;
}

static kfft_plan *plan_rec(int n, int depth) {
  kfft_plan *p = calloc(1, sizeof *p);
  if (!p)
    return NULL;
  p->n = n;
  p->depth = depth;
  /* try a four-step split for long transforms */
  if (n > KFFT_LEAF_MAX && depth < KFFT_MAXDEPTH) {
    int best = 0;
    double const root = sqrt((double)n);
    for (int d = 2; (long)d * d <= n; d++)
      if (n % d == 0)
        best = d; /* largest divisor <= sqrt(n) */
    (void)root;
    if (best > 1) {
      p->n1 = best;
      p->n2 = n / best;
      p->sub1 = plan_rec(p->n1, depth + 1);
      p->sub2 = plan_rec(p->n2, depth + 1);
      int S = 1;
      while ((long)S * S < n)
        S <<= 1;
      p->big_S = S;
      int const nhi = n / S + 1;
      p->big_hi = malloc(sizeof(double complex) * nhi);
      p->big_hi_b = malloc(sizeof(double complex) * nhi);
      p->big_lo = malloc(sizeof(double complex) * S);
      p->big_lo_b = malloc(sizeof(double complex) * S);
      for (int i = 0; i < nhi; i++) {
        p->big_hi[i] = unit_root((long)i * S, n, -1);
        p->big_hi_b[i] = conj(p->big_hi[i]);
      }
      for (int i = 0; i < S; i++) {
        p->big_lo[i] = unit_root(i, n, -1);
        p->big_lo_b[i] = conj(p->big_lo[i]);
      }
      return p;
    }
  }
  p->nrad = factor_radices(n, p->rad);
  p->twf_f = malloc(sizeof(float complex) * n);
  p->twb_f = malloc(sizeof(float complex) * n);
  p->twf_d = malloc(sizeof(double complex) * n);
  p->twb_d = malloc(sizeof(double complex) * n);
  for (int t = 0; t < n; t++) {
    double complex const w = unit_root(t, n, -1);
    p->twf_d[t] = w;
    p->twb_d[t] = conj(w);
    p->twf_f[t] = (float complex)w;
    p->twb_f[t] = (float complex)conj(w);
  }
  return p;
}

kfft_plan *kfft_plan_create(int n) {
  if (n < 1)
    return NULL;
  kfft_plan *p = plan_rec(n, 0);
  if (p) {
    p->rtw = malloc(sizeof(double complex) * ((size_t)n + 1));
    for (int k = 0; k <= n; k++)
      p->rtw[k] = unit_root(k, 2L * n, -1);
  }
  return p;
}

void kfft_plan_destroy(kfft_plan *p) {
  if (!p)
    return;
  kfft_plan_destroy(p->sub1);
  kfft_plan_destroy(p->sub2);
  free(p->twf_f);
  free(p->twb_f);
  free(p->twf_d);
  free(p->twb_d);
  free(p->big_hi);
  free(p->big_lo);
  free(p->big_hi_b);
  free(p->big_lo_b);
  free(p->rtw);
  free(p);
}

int kfft_plan_size(kfft_plan const *p) { return p ? p->n : 0; }

#define REAL float
#define SUFFIX f
#include "fft_cpu_impl.h"
#undef REAL
#undef SUFFIX
#define REAL double
#define SUFFIX d
#include "fft_cpu_impl.h"
#undef REAL
#undef SUFFIX

void kfft_exec_f(kfft_plan const *p, float complex const *in, float complex *out, int sign) {
  exec_f(p, in, out, sign);
}
void kfft_exec_d(kfft_plan const *p, double complex const *in, double complex *out, int sign) {
  exec_d(p, in, out, sign);
}

/* Real input, even length n = 2h: z[j] = x[2j] + i x[2j+1]; Z = FFT_h(z);
 * X[k] = (Z[k] + conj(Z[h-k]))/2 - (i/2) W_n^k (Z[k] - conj(Z[h-k])),  k = 0..h  (Z[h] == Z[0]) */
void kfft_r2c_f(kfft_plan const *half, float const *in, float complex *out) {
  int const h = half->n;
  float complex *z = (float complex *)kfft_scratch(KFFT_NSLOT - 1, sizeof(float complex) * (size_t)h);
  exec_f(half, (float complex const *)in, z, -1);
  for (int k = 0; k <= h; k++) {
    float complex const a = z[k == h ? 0 : k];
    float complex const b = conjf(z[k == 0 ? 0 : h - k]);
    float complex const e = (a + b) * 0.5f;
    float complex const o = (a - b) * 0.5f;
    float complex const w = (float complex)half->rtw[k];
    /* -i*w*o */
    float complex const wo = w * o;
    out[k] = e + CMPLXF(cimagf(wo), -crealf(wo));
  }
}
void kfft_r2c_d(kfft_plan const *half, double const *in, double complex *out) {
  int const h = half->n;
  double complex *z = (double complex *)kfft_scratch(KFFT_NSLOT - 1, sizeof(double complex) * (size_t)h);
  exec_d(half, (double complex const *)in, z, -1);
  for (int k = 0; k <= h; k++) {
    double complex const a = z[k == h ? 0 : k];
    double complex const b = conj(z[k == 0 ? 0 : h - k]);
    double complex const e = (a + b) * 0.5;
    double complex const o = (a - b) * 0.5;
    double complex const wo = half->rtw[k] * o;
    out[k] = e + CMPLX(cimag(wo), -creal(wo));
  }
}
/* Inverse of the above (unnormalised: result = n * ifft): Z[k] = E[k] + i*conj(W_n^k)*O[k] with
 * E = (X[k] + conj(X[h-k])), O = (X[k] - conj(X[h-k])); then x = IFFT_h(Z) interleaved. */
void kfft_c2r_f(kfft_plan const *half, float complex const *in, float *out) {
  int const h = half->n;
  float complex *z = (float complex *)kfft_scratch(KFFT_NSLOT - 1, sizeof(float complex) * (size_t)h);
  for (int k = 0; k < h; k++) {
    /* FFTW's c2r works on the halfcomplex representation, which has no slot for the imaginary parts of the DC and
     * Nyquist bins: whatever the caller left there (the reference leaves X[shift] R[0], filter.c:803-809) is ignored */
    float complex const a = (k == 0) ? (float complex)crealf(in[0]) : in[k];
    float complex const b = (k == 0) ? (float complex)crealf(in[h]) : conjf(in[h - k]);
    float complex const e = a + b;
    float complex const o = a - b;
    float complex const wo = (float complex)conj(half->rtw[k]) * o;
    z[k] = e + CMPLXF(-cimagf(wo), crealf(wo)); /* + i*wo */
  }
  exec_f(half, z, (float complex *)out, +1);
}
