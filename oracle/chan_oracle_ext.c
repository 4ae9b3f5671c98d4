/* oracle/chan_oracle_ext.c -- CPU restatement of the remaining slice variants of execute_filter_output and of the
 * per-channel steps that follow it in the reference's downconvert().
 * TEST INFRASTRUCTURE, NOT PRODUCT (see chan_oracle.h).  Pinned against the reference's own filter.c / radio.c /
 * osc.c compiled unmodified (oracle/_ref, tests/test_oracle_ext_cpu.py).
 */
#define _GNU_SOURCE 1
#include <complex.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "chan_oracle.h"
#include "fft_cpu.h"

static int modulo(int x, int const m) { /* misc.h: always non-negative remainder */
  x = x % m;
  return x < 0 ? x + m : x;
}

/* ---------------------------------------------------------------- slice variants ------------- */
/* COMPLEX master, COMPLEX out, beam == true (filter.c:756-775; weights filter.c:922-929).  Same walk as the plain
 * variant; in-domain shifts only (|shift| < N/2): what the reference leaves untouched after its loop is zero here. */
void ko_slice_beam(int m_bins, float complex const *X, int s_bins, float complex const *R, int shift,
                   double complex alpha, double complex beta, float complex *S) {
  int const top = (s_bins + 1) / 2, m_nyq = (m_bins + 1) / 2;
  int wp = top;
  long rp = (long)shift - s_bins / 2;
  for (int i = 0; i < s_bins; i++)
    S[i] = 0;
  int t = 0;
  while (t < s_bins && rp < -(long)m_nyq) { /* filter.c:733-743 */
    rp++;
    t++;
    if (++wp == s_bins)
      wp = 0;
  }
  if (t < s_bins) {
    if (rp < 0)
      rp += m_bins;
    if (rp >= 0 && rp < m_bins) {
      do {
        if (rp == 0 || rp == m_bins / 2) /* filter.c:765-767 */
          S[wp] = (float complex)(crealf(X[rp]) * alpha * R[wp] + cimagf(X[rp]) * beta * R[wp]);
        else /* filter.c:769-770 */
          S[wp] = (float complex)((alpha * X[rp] + beta * conjf(X[m_bins - rp])) * R[wp]);
        if (++rp == m_bins)
          rp = 0;
        if (++wp == s_bins)
          wp = 0;
      } while (wp != top && rp != m_nyq);
    }
  }
  S[(s_bins + 1) / 2] = 0; /* filter.c:911 */
}

/* REAL output slaves: s_bins = points/2 + 1 positive-frequency bins (filter.c:374).
 * REAL master  (filter.c:803-809): S[si] = X[si+shift] R[si], zero outside [0, m_bins)
 * COMPLEX master (filter.c:794-802, "UNTESTED" in the reference): S[si] = R[si] (X[mi mod m] + conj X[(m-mi) mod m]),
 * mi = si + shift in [-m/2, m/2).  Then the Nyquist-zero line filter.c:911 hits index (s_bins+1)/2 -- for a REAL slave
 * that is a bin near points/4, in the middle of the band; it is what the reference does, so it is what we do. */
void ko_slice_realout(int in_type, int m_bins, float complex const *X, int points, float complex const *R, int shift,
                      float complex *S) {
  int const s_bins = points / 2 + 1;
  for (int si = 0; si < s_bins; si++) {
    int const mi = si + shift;
    float complex v = 0;
    if (in_type == KO_REAL) {
      if (mi >= 0 && mi < m_bins)
        v = X[mi] * R[si];
    } else if (mi >= -m_bins / 2 && mi < m_bins / 2)
      v = R[si] * (X[modulo(mi, m_bins)] + conjf(X[modulo(m_bins - mi, m_bins)]));
    S[si] = v;
  }
  S[(s_bins + 1) / 2] = 0;
}

/* full[0..points): the c2r inverse (filter.c:914 with rev_plan = plan_c2r, filter.c:386); user part = last olen */
int ko_channel_block_realout(int in_type, int m_bins, float complex const *X, int points, float complex const *R,
                             int shift, float *full) {
  if (points < 2 || (points & 1))
    return -1;
  float complex *S = malloc(sizeof(float complex) * (size_t)(points / 2 + 1));
  if (!S)
    return -1;
  ko_slice_realout(in_type, m_bins, X, points, R, shift, S);
  kfft_plan *p = kfft_plan_create(points / 2);
  kfft_c2r_f(p, S, full);
  kfft_plan_destroy(p);
  free(S);
  return 0;
}

int ko_channel_block_beam(int m_bins, float complex const *X, int points, float complex const *R, int shift,
                          double are, double aim, double bre, double bim, float complex *full) {
  float complex *S = malloc(sizeof(float complex) * (size_t)points);
  if (!S)
    return -1;
  ko_slice_beam(m_bins, X, points, R, shift, CMPLX(are, aim), CMPLX(bre, bim), S);
  kfft_plan *p = kfft_plan_create(points);
  kfft_exec_f(p, S, full, +1);
  kfft_plan_destroy(p);
  free(S);
  return 0;
}

/* ---------------------------------------------------------------- fine tuning ---------------- */
/* exp(i pi x) as sincospi.c:24-66 / misc.h:273-277 do it (exact reduction in half-turns) */
static double complex cis_pi(double x) {
  double y = x - 2.0 * floor(0.5 * x);
  if (y < 0)
    y += 2.0;
  if (y >= 2.0)
    y -= 2.0;
  int const quadrant = (int)(2.0 * y);
  double z = y - 0.5 * quadrant;
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
  case 0: return CMPLX(cz, sz);
  case 1: return CMPLX(-sz, cz);
  case 2: return CMPLX(-cz, -sz);
  default: return CMPLX(sz, -cz);
  }
}

/* struct osc + set_osc/step_osc (osc.c:15,18-70) */
static int phasor_ok(double complex x) { return !(isnan(creal(x)) || isnan(cimag(x)) || creal(x) * creal(x) + cimag(x) * cimag(x) < 0.9); }
void ko_osc_set(struct ko_osc *o, double f, double r) {
  if (!phasor_ok(o->phasor)) {
    o->phasor = 1;
    o->steps = 16384;
    o->freq = 0;
    o->rate = 0;
    o->phasor_step = 1;
    o->phasor_step_step = 1;
  }
  if (f != o->freq) {
    o->freq = f;
    o->phasor_step = cis_pi(2 * o->freq);
  }
  if (r != o->rate) {
    o->rate = r;
    o->phasor_step_step = cis_pi(2 * o->rate);
  }
}
double complex ko_osc_step(struct ko_osc *o) {
  if (--o->steps <= 0) { /* renorm_osc, osc.c:46-57 */
    if (!phasor_ok(o->phasor))
      o->phasor = 1;
    o->steps = 16384;
    o->phasor /= cabs(o->phasor);
    if (o->rate != 0)
      o->phasor_step /= cabs(o->phasor_step);
  }
  double complex const r = o->phasor;
  if (o->rate != 0)
    o->phasor_step *= o->phasor_step_step;
  o->phasor *= o->phasor_step;
  return r;
}

void ko_finetune_init(struct ko_finetune *s) {
  memset(s, 0, sizeof *s);
  s->remainder = NAN;      /* modes.c:265 */
  s->bin_shift = -1000999; /* modes.c:266, the "something bizarre" radio.c:1487 asks for */
  s->phase_adjust = 1;
}

/* One block of radio.c:1476-1501 + :1515-1520 on the olen fresh samples y[] of a channel:
 * set_osc on retune, block phase adjust (a) every block and (b) once per shift change, per-sample rotation,
 * then the mean power.  L, M: the MASTER's block and impulse lengths (V = 1 + L/(M-1), radio.c:1490). */
double ko_finetune_block(struct ko_finetune *s, int L, int M, int shift, double remainder, double out_samprate,
                         double doppler_rate, float complex *y, int olen) {
  if (shift != s->bin_shift || isnan(s->remainder) || remainder != s->remainder) {
    ko_osc_set(&s->fine, -remainder / out_samprate, doppler_rate / (out_samprate * out_samprate));
    s->remainder = remainder;
  }
  if (shift != s->bin_shift) {
    int const V = 1 + (L / (M - 1));
    s->phase_adjust = cis_pi(2.0 * (shift % V) / (double)V);
    s->fine.phasor *= cis_pi((shift - s->bin_shift) / (-2.0 * (V - 1)));
    s->bin_shift = shift;
  }
  s->fine.phasor *= s->phase_adjust;
  for (int n = 0; n < olen; n++)
    y[n] = (float complex)((double complex)y[n] * ko_osc_step(&s->fine));
  double energy = 0;
  for (int n = 0; n < olen; n++)
    energy += crealf(y[n]) * crealf(y[n]) + cimagf(y[n]) * cimagf(y[n]);
  return energy / olen;
}

/* ---------------------------------------------------------------- noise estimate ------------- */
/* quickselect / quantile exactly as radio.c:1722-1775 (same pivot rule, same interpolation) */
static void dswap(double *a, double *b) {
  double const t = *a;
  *a = *b;
  *b = t;
}
static int partition(double *arr, int left, int right, int pivot_index) {
  double const pv = arr[pivot_index];
  dswap(&arr[pivot_index], &arr[right]);
  int store = left;
  for (int i = left; i < right; i++)
    if (arr[i] < pv) {
      dswap(&arr[store], &arr[i]);
      store++;
    }
  dswap(&arr[right], &arr[store]);
  return store;
}
static double quickselect(double *arr, int left, int right, int k) {
  while (left < right) {
    int const pi = left + (right - left) / 2;
    int const pn = partition(arr, left, right, pi);
    if (pn == k)
      return arr[k];
    else if (k < pn)
      right = pn - 1;
    else
      left = pn + 1;
  }
  return arr[left];
}
static double quantile(double *a, int n, double p) {
  if (n == 0)
    return NAN;
  double const pos = p * (n - 1);
  int const i = (int)floor(pos);
  double const frac = pos - i;
  double const q1 = quickselect(a, 0, n - 1, i);
  if (frac == 0.0)
    return q1;
  double const q2 = quickselect(a, 0, n - 1, i + 1);
  return q1 + frac * (q2 - q1);
}

/* radio.c:1783-1866: N0 estimate from >= 1000 master bins around the channel.  Returns W/Hz in the
 * reference's scaling (per master bin, / (bins * samprate)).  The COMPLEX-master branch stops filling at the master's
 * Nyquist bin (radio.c:1832-1833) and the reference then reads uninitialised stack: callers stay inside that domain. */
double ko_estimate_noise(int in_type, int m_bins, float complex const *X, int s_bins, int shift, double samprate) {
  if (s_bins <= 0)
    return 0;
  int nbins = s_bins < 1000 ? 1000 : s_bins; /* Min_noise_bins, radio.c:76 */
  double *e = calloc((size_t)nbins, sizeof *e);
  if (in_type == KO_REAL) {
    int mbin = abs(shift) - nbins / 2;
    if (mbin < 0)
      mbin = 0;
    else if (mbin + nbins > m_bins)
      mbin = m_bins - nbins;
    for (int i = 0; i < nbins; i++, mbin++)
      e[i] = crealf(X[mbin]) * crealf(X[mbin]) + cimagf(X[mbin]) * cimagf(X[mbin]);
  } else {
    int mbin = shift - nbins / 2;
    if (mbin < 0)
      mbin += m_bins;
    else if (mbin >= m_bins)
      mbin -= m_bins;
    if (mbin < 0 || mbin >= m_bins) {
      free(e);
      return 0;
    }
    for (int i = 0; i < nbins; i++) {
      e[i] = crealf(X[mbin]) * crealf(X[mbin]) + cimagf(X[mbin]) * cimagf(X[mbin]);
      if (++mbin == m_bins)
        mbin = 0;
      if (mbin == m_bins / 2)
        break;
    }
  }
  double const NQ = 0.10, N_cutoff = 1.5; /* radio.c:73-74 */
  double const z = N_cutoff * (-log(1 - NQ));
  double const correction = 1 / (1 - z * exp(-z) / (1 - exp(-z)));
  double const en = N_cutoff * quantile(e, nbins, NQ);
  double energy = 0;
  int noisebins = 0;
  for (int i = 0; i < nbins; i++)
    if (e[i] <= en) {
      energy += e[i];
      noisebins++;
    }
  free(e);
  if (noisebins == 0)
    return 0;
  energy /= noisebins;
  return energy * correction / ((double)m_bins * samprate);
}

/* ---------------------------------------------------------------- Airspy 12-bit packed ingest ---------- */
/* airspy-unpack.c:106-130 (portable version; the AVX2 one :17-104 computes the same): 8 offset-binary 12-bit samples in
 * three 32-bit words, most significant first; x = s - 2048, float = scale * x, energy += x*x, clip count x == 2047 or
 * x <= -2047.  sampcount must be a multiple of 8.  Returns the clip count. */
int ko_airspy_unpack(float *dst, uint32_t const *up, int sampcount, float scale, uint64_t *energy) {
  int over = 0;
  for (int i = 0; i < sampcount; i += 8, up += 3, dst += 8) {
    uint32_t s[8];
    s[0] = up[0] >> 20;
    s[1] = up[0] >> 8;
    s[2] = (up[0] << 4) | (up[1] >> 28);
    s[3] = up[1] >> 16;
    s[4] = up[1] >> 4;
    s[5] = (up[1] << 8) | (up[2] >> 24);
    s[6] = up[2] >> 12;
    s[7] = up[2];
    for (int j = 0; j < 8; j++) {
      int const x = (int)(s[j] & 0xfff) - 2048;
      over += (x == 2047 || x <= -2047);
      dst[j] = scale * (float)x;
      if (energy)
        *energy += (uint64_t)((int64_t)x * x);
    }
  }
  return over;
}

/* ---------------------------------------------------------------- FM front half -------------- */
/* fm.c:104-131 (amplitude statistics of the variance-based SNR estimator: mean of cabsf, then the sum of squared
 * deviations in a second pass) and fm.c:205-231 (plain quadrature discriminator, no PLL, no threshold extension):
 *   s = x[n] conj(x[n-1]) in double complex, phase = carg(s) / pi, x[-1] = phase_memory carried across blocks (0 at start).
 * demod_fm() cannot be isolated from the rest of radiod (squelch, PL, de-emphasis, RTP output in one loop), so this part
 * is restated from the source only: PARITY UNPINNED for this function. */
void ko_fm_front(float complex const *x, int n, double complex *phase_memory, float *baseband, double *avg_amp,
                 double *variance_sum) {
  double avg = 0;
  double *amp = malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++)
    avg += amp[i] = cabsf(x[i]); /* fm.c:117 */
  avg /= n;
  double var = 0;
  for (int i = 0; i < n; i++) /* fm.c:122-123 */
    var += (amp[i] - avg) * (amp[i] - avg);
  free(amp);
  if (avg_amp)
    *avg_amp = avg;
  if (variance_sum)
    *variance_sum = var;
  double complex pm = *phase_memory;
  for (int i = 0; i < n; i++) { /* fm.c:211-229 with fm.threshold == false */
    double complex const s = x[i] * conj(pm);
    baseband[i] = (float)(M_1_PI * carg(s));
    pm = x[i];
  }
  *phase_memory = pm;
}
