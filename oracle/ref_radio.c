/* oracle/ref_radio.c -- drives the reference's OWN downconvert() (radio.c:1410-1523: tuning, execute_filter_output,
 * estimate_noise, fine-tuning oscillator, block phase, baseband power) for the oracle.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Compiled only into oracle/_ref/libka9qradio.so together with the reference's
 * unmodified src/radio.c (compiled where it lies, against declaration-only stubs of iniparser.h / opus.h /
 * fftw3.h) and the same filter.c, osc.c, misc.c ... objects as libka9qref.so.  radio.c's other entry points
 * (config parsing, threads, status) are never called; the functions they would need from files that are not part
 * of the path are satisfied by aborting stubs in ref_radio_stubs.c.
 *
 * `Frontend` is a process-wide global of radio.c (radio.c:82), so one session at a time.
 */
#define _GNU_SOURCE 1
#include <complex.h>
#include <math.h>
#include <pthread.h>
#include <stdbool.h>
#include <stdlib.h>
#include <string.h>
#include "radio.h" /* the reference's header, through -iquote /root/reference/src */

extern double Blocktime; /* radio.c:130 */

struct rr_session {
  int nchan;
  chan_t *chan[64];
  bool open;
};
static struct rr_session S;

int rr_open(int L, int M, int in_type, double samprate, double frequency) {
  if (S.open)
    return -1;
  memset(&S, 0, sizeof S);
  N_worker_threads = 0; /* inline forward transform: deterministic */
  memset(&Frontend.in, 0, sizeof Frontend.in);
  Frontend.samprate = samprate;
  Frontend.frequency = frequency;
  Frontend.L = L;
  Frontend.M = M;
  Frontend.isreal = (in_type == REAL);
  Blocktime = L / samprate; /* radio.c:587 */
  pthread_mutex_init(&Frontend.status_mutex, NULL);
  pthread_cond_init(&Frontend.status_cond, NULL);
  if (create_filter_input(&Frontend.in, L, M, (enum filtertype)in_type) != 0) /* radio.c:599 */
    return -1;
  S.open = true;
  return 0;
}

/* one channel as fm.c:27-34 / radio.c:1559-1611 set it up; freq = carrier frequency in Hz */
int rr_add_channel(int olen, double out_samprate, double freq, double low, double high, double beta) {
  if (!S.open || S.nchan == 64)
    return -1;
  chan_t *c = calloc(1, sizeof *c);
  c->frontend = &Frontend;
  c->tune.freq = freq;
  c->output.samprate = (int)lrint(out_samprate);
  c->filter.min_IF = low * out_samprate;
  c->filter.max_IF = high * out_samprate;
  c->filter.kaiser_beta = beta;
  c->filter.remainder = NAN;     /* modes.c:265 */
  c->filter.bin_shift = -1000999; /* modes.c:266 */
  c->sig.n0 = NAN;
  if (create_filter_output(&c->filter.out, &Frontend.in, olen, COMPLEX) != 0 ||
      set_filter(&c->filter.out, low, high, beta) != 0) {
    free(c);
    return -1;
  }
  S.chan[S.nchan] = c;
  return S.nchan++;
}
int rr_set_freq(int ch, double freq, double doppler, double doppler_rate) {
  S.chan[ch]->tune.freq = freq;
  S.chan[ch]->tune.doppler = doppler;
  S.chan[ch]->tune.doppler_rate = doppler_rate;
  return 0;
}
int rr_write_real(float const *x, int n) { return write_rfilter(&Frontend.in, x, n); }
int rr_write_complex(float complex const *x, int n) { return write_cfilter(&Frontend.in, x, n); }

/* radio.c:1410: returns downconvert()'s value; baseband (olen samples after the fine-tuning rotation),
 * bb_power (radio.c:1515-1520), n0 (radio.c:1468-1474), the shift/remainder it computed (filter.bin_shift, .remainder) */
int rr_downconvert(int ch, float complex *baseband, double *bb_power, double *n0, int *shift, double *remainder) {
  chan_t *c = S.chan[ch];
  int const r = downconvert(c);
  if (r == 0 && c->baseband) {
    if (baseband)
      memcpy(baseband, c->baseband, sizeof(float complex) * (size_t)c->sampcount);
    if (bb_power)
      *bb_power = c->sig.bb_power;
    if (n0)
      *n0 = c->sig.n0;
    if (shift)
      *shift = c->filter.bin_shift;
    if (remainder)
      *remainder = c->filter.remainder;
  }
  return r;
}
/* the last block's spectrum, for feeding the restatement */
int rr_get_spectrum(float complex *dst) {
  unsigned const job = Frontend.in.next_jobnum - 1;
  memcpy(dst, Frontend.in.fdomain[job % ND], sizeof(float complex) * (size_t)Frontend.in.bins);
  return Frontend.in.bins;
}
int rr_master_bins(void) { return Frontend.in.bins; }

void rr_close(void) {
  if (!S.open)
    return;
  for (int i = 0; i < S.nchan; i++) {
    delete_filter_output(&S.chan[i]->filter.out);
    free(S.chan[i]);
  }
  delete_filter_input(&Frontend.in);
  S.open = false;
  S.nchan = 0;
}

/* the reference's own oscillator, step by step (osc.c:28-70), for pinning ko_osc_* */
void rr_osc_run(double f, double r, long n, double complex *out) {
  struct osc o;
  memset(&o, 0, sizeof o);
  set_osc(&o, f, r);
  for (long i = 0; i < n; i++)
    out[i] = step_osc(&o);
}
