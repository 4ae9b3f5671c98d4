/* oracle/chan_oracle.c -- CPU restatement of the reference's overlap-save channelizer path.
 * TEST INFRASTRUCTURE, NOT PRODUCT (see chan_oracle.h).  Every function names the reference
 * lines it follows; nothing here is called by the shipped CUDA path.
 */
#define _GNU_SOURCE 1
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "chan_oracle.h"
#include "fft_cpu.h"

/* ---------------------------------------------------------------- small plan cache --------- */
/* The reference keeps one FFTW plan per filter; the oracle is stateless, so cache by size. */
#include <pthread.h>
static pthread_mutex_t Plan_lock = PTHREAD_MUTEX_INITIALIZER;
static struct {
  int n;
  kfft_plan *p;
} Plans[64];
static kfft_plan *plan_for(int n) {
  pthread_mutex_lock(&Plan_lock);
  kfft_plan *r = NULL;
  int i;
  for (i = 0; i < 64 && Plans[i].p; i++)
    if (Plans[i].n == n) {
      r = Plans[i].p;
      break;
    }
  if (!r && i < 64) {
    Plans[i].n = n;
    r = Plans[i].p = kfft_plan_create(n);
  }
  pthread_mutex_unlock(&Plan_lock);
  return r;
}

/* ---------------------------------------------------------------- response design ---------- */
/* sin/cos of pi*x with exact range reduction in revolutions (sincospi.c:24-66). */
static void sincos_pi(double x, double *s, double *c) {
  if (!isfinite(x)) {
    *s = *c = NAN;
    return;
  }
  double y = x - 2.0 * floor(0.5 * x); /* [0,2) */
  if (y < 0)
    y += 2.0;
  if (y >= 2.0)
    y -= 2.0;
  int const quadrant = (int)(2.0 * y);   /* 0..3 */
  double z = y - 0.5 * quadrant;         /* [0,0.5) */
  int const swap = z > 0.25;
  if (swap)
    z = 0.5 - z;
  double sz = sin(M_PI * z), cz = cos(M_PI * z);
  if (swap) {
    double const t = sz;
    sz = cz;
    cz = t;
  }
  switch (quadrant) {
  case 0: *s = sz; *c = cz; break;
  case 1: *s = cz; *c = -sz; break;
  case 2: *s = -sz; *c = -cz; break;
  default: *s = -cz; *c = sz; break;
  }
}
static double complex cis_pi(double x) { /* misc.h:273-277 */
  double s, c;
  sincos_pi(x, &s, &c);
  return CMPLX(c, s);
}
/* modified Bessel I0 by its power series, <= 40 terms, 1e-12 relative cutoff (misc.c:416-427) */
static double bessel_i0(double z) {
  double const q = 0.25 * z * z;
  double term = q, sum = 1.0 + q;
  for (int k = 2; k < 40; k++) {
    term *= q / ((double)k * k);
    sum += term;
    if (term < 1e-12 * sum)
      break;
  }
  return sum;
}
static double sinc_pi(double x) { /* misc.h:217-221 */
  return x == 0 ? 1.0 : sin(M_PI * x) / (M_PI * x);
}

int ko_design_response(int points, int olen, int master_points, int master_real, double low, double high,
                       double kaiser_beta, float complex *response) {
  if (isnan(low) || isnan(high) || isnan(kaiser_beta) || response == NULL)
    return -1;
  /* filter.c:976-984: order the edges, clamp to the output Nyquist band */
  if (low > high) {
    double const t = low;
    low = high;
    high = t;
  }
  low = fmin(fmax(low, -0.5), 0.5);
  high = fmin(fmax(high, -0.5), 0.5);
  int const M = points - olen + 1; /* filter.c:986-990 */
  if (M < 2)
    return -1;
  double const half_bw = (high == low) ? 1e-4 : 0.5 * fabs(high - low); /* filter.c:992 */
  double const center = 0.5 * (high + low);

  /* Kaiser window, float storage, symmetric fill, centre tap 1 for odd M (window.c:217-238) */
  float *win = malloc(sizeof(float) * (size_t)M);
  double const inv_i0 = 1.0 / bessel_i0(kaiser_beta);
  double const step = 2.0 / (M - 1);
  for (int n = 0; n < M / 2; n++) {
    double const p = step * n - 1;
    float const w = (float)(bessel_i0(kaiser_beta * sqrt(1 - p * p)) * inv_i0);
    win[n] = win[M - 1 - n] = w;
  }
  if (M & 1)
    win[(M - 1) / 2] = 1;
  /* normalise so the taps sum to M (window.c:240-254) */
  double wsum = 0;
  for (int n = 0; n < M; n++)
    wsum += win[n];
  if (wsum == 0 || !isfinite(wsum)) {
    free(win);
    return -1;
  }
  float const wnorm = (float)(M / wsum);
  for (int n = 0; n < M; n++)
    win[n] *= wnorm;

  /* causal windowed-sinc bandpass in the first M of `points` slots (filter.c:1009-1019) */
  memset(response, 0, sizeof(float complex) * (size_t)points);
  double tap_sum = 0;
  for (int i = 0; i < M; i++) {
    double const n = i - 0.5 * (double)(M - 1);
    double const r = win[i] * 2 * half_bw * sinc_pi(2 * half_bw * n);
    tap_sum += r;
    response[i] = (float complex)(cis_pi(2 * center * n) * r);
  }
  /* gain: sqrt2 for real input, window loss, 1/N of the unnormalised master FFT (filter.c:1020-1028) */
  double const gain = (master_real ? M_SQRT2 : 1.0) / (tap_sum * master_points);
  for (int i = 0; i < M; i++)
    response[i] = (float complex)((double complex)response[i] * gain);
  free(win);
  kfft_exec_f(plan_for(points), response, response, -1); /* filter.c:1007,1030 */
  return 0;
}

/* ---------------------------------------------------------------- forward transform -------- */
int ko_forward_real(int n, float const *window, float complex *spectrum) {
  if (n < 2)
    return -1;
  if (n % 2 == 0) {
    kfft_r2c_f(plan_for(n / 2), window, spectrum);
  } else {
    float complex *t = malloc(sizeof(float complex) * (size_t)n);
    for (int i = 0; i < n; i++)
      t[i] = window[i];
    kfft_exec_f(plan_for(n), t, t, -1);
    memcpy(spectrum, t, sizeof(float complex) * (size_t)(n / 2 + 1));
    free(t);
  }
  return 0;
}
int ko_forward_real_d(int n, double const *window, double complex *spectrum) {
  if (n < 2 || (n & 1))
    return -1;
  kfft_r2c_d(plan_for(n / 2), window, spectrum);
  return 0;
}
int ko_forward_complex(int n, float complex const *window, float complex *spectrum) {
  if (n < 1)
    return -1;
  kfft_exec_f(plan_for(n), window, spectrum, -1);
  return 0;
}

void ko_apply_notches(struct ko_notch *list, float complex *spectrum) { /* filter.c:464-474 */
  if (!list || !spectrum)
    return;
  for (;; list++) {
    list->state += list->alpha * ((double complex)spectrum[list->bin] - list->state);
    spectrum[list->bin] = (float complex)((double complex)spectrum[list->bin] - list->state);
    if (list->bin == 0)
      break;
  }
}

/* ---------------------------------------------------------------- slice x response --------- */
void ko_slice_multiply(int in_type, int m_bins, float complex const *X, int s_bins, float complex const *R,
                       int shift, int isb, float complex *S) {
  int const lo = -(s_bins / 2); /* most negative output bin; lives at index (s_bins+1)/2 */
  if (in_type == KO_REAL) {
    /* filter.c:810-893.  Output bin k in [-floor(Ns/2), ceil(Ns/2)-1] -> index k mod Ns.
     * shift >= 0: upright, q = shift+k; shift < 0: inverted, q = -(shift+k), conjugated.
     * Never folds across DC or past the last master bin: out-of-range -> 0. */
    for (int t = 0; t < s_bins; t++) {
      int const k = lo + t;
      int const wp = ((k % s_bins) + s_bins) % s_bins;
      long const q = (shift >= 0) ? (long)shift + k : -((long)shift + k);
      if (q < 0 || q >= m_bins)
        S[wp] = 0;
      else
        S[wp] = (shift >= 0 ? X[q] : conjf(X[q])) * R[wp];
    }
  } else {
    /* filter.c:728-793 (non-beam).  A walk, because outside |shift| < N/2 the reference's
     * behaviour is defined by its loop, not by a formula (SURVEY.md 8a note). */
    int wp = (s_bins + 1) / 2;
    long rp = (long)shift - s_bins / 2;
    int const top = (s_bins + 1) / 2;       /* write index at which the output is complete */
    int const m_nyq = (m_bins + 1) / 2;     /* master read index at which copying stops */
    int t = 0;
    /* leading zeros while below the master's most negative bin */
    while (t < s_bins && rp < -(long)m_nyq) {
      S[wp] = 0;
      rp++;
      t++;
      if (++wp == s_bins)
        wp = 0;
    }
    if (t < s_bins) {
      if (rp < 0)
        rp += m_bins;
      if (rp >= 0 && rp < m_bins) {
        do { /* copy until the output is full or the master hits its Nyquist index */
          S[wp] = X[rp] * R[wp];
          t++;
          if (++rp == m_bins)
            rp = 0;
          if (++wp == s_bins)
            wp = 0;
        } while (wp != top && rp != m_nyq);
      }
      while (wp != top) { /* whatever is left is zero */
        S[wp] = 0;
        if (++wp == s_bins)
          wp = 0;
      }
    }
  }
  if (isb) { /* filter.c:895-909 */
    for (int p = 1, dn = s_bins - 1; p < s_bins / 2; p++, dn--) {
      float complex const pos = S[p], neg = S[dn];
      S[p] = pos + conjf(neg);
      S[dn] = neg - conjf(pos);
    }
    S[0] = 0;
  }
  S[(s_bins + 1) / 2] = 0; /* filter.c:911 */
}

int ko_channel_block(int in_type, int m_bins, float complex const *X, int points, float complex const *R,
                     int shift, int isb, float complex *full) {
  float complex *S = malloc(sizeof(float complex) * (size_t)points);
  if (!S)
    return -1;
  ko_slice_multiply(in_type, m_bins, X, points, R, shift, isb, S);
  kfft_exec_f(plan_for(points), S, full, +1); /* filter.c:914, FFTW_BACKWARD, unnormalised */
  free(S);
  return 0;
}

/* ---------------------------------------------------------------- ingest ------------------- */
int ko_convert_i16(float *dst, int16_t const *src, int n, float scale, uint64_t *energy, int randomize) {
  /* rx888.c:753-767 (portable) / :694-750 (AVX2, the path that runs on x86-64).  The de-randomiser
   * follows the AVX2 lanes (16-bit shifts, rx888.c:707-712): lsb set -> flip bits 1..15. */
  int clips = 0;
  for (int i = 0; i < n; i++) {
    int16_t x = src[i];
    if (randomize)
      x ^= (int16_t)((int16_t)(x << 15) >> 14); /* lsb set -> flip all other bits */
    if (energy)
      *energy += (uint64_t)((int32_t)x * x);
    clips += (x > 32766 || x < -32766);
    dst[i] = (float)x * scale;
  }
  return clips;
}

void ko_block_window_real(float const *stream, int L, int M, int b, float *window) {
  long const start = (long)b * L - (M - 1); /* filter.c:244/259: ring zeroed, writes start M-1 in */
  for (long i = 0; i < (long)L + M - 1; i++)
    window[i] = (start + i < 0) ? 0.0f : stream[start + i];
}
void ko_block_window_complex(float complex const *stream, int L, int M, int b, float complex *window) {
  long const start = (long)b * L - (M - 1);
  for (long i = 0; i < (long)L + M - 1; i++)
    window[i] = (start + i < 0) ? 0.0f : stream[start + i];
}

/* ---------------------------------------------------------------- synthetic source --------- */
/* xoshiro256** with splitmix64 seeding (public-domain algorithm, as used at gauss.c:17-62) */
struct xo {
  uint64_t s[4];
};
static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static void xo_seed(struct xo *g, uint64_t seed) {
  for (int i = 0; i < 4; i++) {
    uint64_t z = (seed += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    g->s[i] = z ^ (z >> 31);
  }
  if (!(g->s[0] | g->s[1] | g->s[2] | g->s[3]))
    g->s[0] = 1;
}
static uint64_t xo_next(struct xo *g) {
  uint64_t const out = rotl(g->s[1] * 5, 7) * 9, t = g->s[1] << 17;
  g->s[2] ^= g->s[0];
  g->s[3] ^= g->s[1];
  g->s[1] ^= g->s[2];
  g->s[0] ^= g->s[3];
  g->s[2] ^= t;
  g->s[3] = rotl(g->s[3], 45);
  return out;
}
/* popcount-sum Gaussian approximation (gauss.c:102-110) */
static double gauss_from(struct xo *g) {
  uint64_t const u = xo_next(g);
  double x = __builtin_popcountll(u * 0x2c1b3c6dULL) + __builtin_popcountll(u * 0x297a2d39ULL) - 64;
  x += (double)(int64_t)u * (1 / 9223372036854775808.);
  return x * 0.1765469659009499;
}
/* complex rotator with periodic renormalisation (osc.c:15,28-70) */
struct rot {
  double complex phasor, step;
  int steps;
};
static void rot_init(struct rot *r, double cycles_per_sample) {
  r->phasor = 1;
  r->steps = 16384;
  r->step = (cycles_per_sample != 0) ? cis_pi(2 * cycles_per_sample) : 1; /* osc.c:40-43: freq 0 keeps step 1 */
}
static double complex rot_step(struct rot *r) {
  if (--r->steps <= 0) {
    r->steps = 16384;
    r->phasor /= cabs(r->phasor);
  }
  double complex const out = r->phasor;
  r->phasor *= r->step;
  return out;
}
struct ko_siggen {
  struct xo rng;
  struct rot carrier;
};
ko_siggen *ko_siggen_new(double cycles_per_sample) {
  ko_siggen *g = calloc(1, sizeof *g);
  xo_seed(&g->rng, 1); /* gauss.c:95-100 */
  rot_init(&g->carrier, cycles_per_sample);
  return g;
}
void ko_siggen_free(ko_siggen *g) { free(g); }
void ko_siggen_real(ko_siggen *g, float *dst, long n, double amplitude, double noise, double scale) {
  for (long i = 0; i < n; i++) { /* sig_gen.c:292-296 */
    double const samp = amplitude * creal(rot_step(&g->carrier)) + noise * gauss_from(&g->rng);
    dst[i] = (float)(samp * scale);
  }
}
void ko_siggen_complex(ko_siggen *g, float complex *dst, long n, double amplitude, double noise, double scale) {
  for (long i = 0; i < n; i++) { /* sig_gen.c:318-322, misc.h:399-403 (real part drawn first) */
    double complex const car = amplitude * rot_step(&g->carrier);
    double const nr = gauss_from(&g->rng), ni = gauss_from(&g->rng);
    dst[i] = (float complex)((car + noise * CMPLX(nr, ni)) * scale);
  }
}
void ko_siggen_tones_i16(int16_t *dst, long n, int ntones, double const *cycles_per_sample,
                         double const *amplitude, double noise, uint64_t seed) {
  struct xo rng;
  xo_seed(&rng, seed);
  struct rot *osc = calloc((size_t)ntones, sizeof *osc);
  for (int k = 0; k < ntones; k++)
    rot_init(&osc[k], cycles_per_sample[k]);
  for (long i = 0; i < n; i++) {
    double x = noise * gauss_from(&rng);
    for (int k = 0; k < ntones; k++)
      x += amplitude[k] * creal(rot_step(&osc[k]));
    long v = lrint(32767.0 * x);
    dst[i] = (int16_t)(v > 32767 ? 32767 : v < -32767 ? -32767 : v);
  }
  free(osc);
}

int ko_compute_tuning(int N, double samprate, double freq, int *shift, double *remainder) {
  double const hz_per_bin = samprate / N; /* radio.c:1175-1199 */
  int const r = (int)lrint(freq / hz_per_bin);
  if (shift)
    *shift = r;
  if (remainder)
    *remainder = fma(-(double)r, hz_per_bin, freq);
  return abs(r) >= N / 2 ? -1 : 0;
}
