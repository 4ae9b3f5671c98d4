/* oracle/ref_driver.c -- thin ctypes-friendly driver around the reference's OWN filter path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  This file is compiled only into oracle/_ref/libka9qref.so
 * together with the reference's unmodified src/filter.c, window.c, misc.c, sched.c, sincospi.c,
 * sincospif.c, osc.c, gauss.c (compiled where they lie under /root/reference, never copied)
 * and oracle/fftw_shim.c.  It drives the filter.h surface the way radiod does:
 *   setup      radio.c:582-620   (L, M, create_filter_input, notch list)
 *   producer   rx888.c:800-826 / sig_gen.c:288-317   (write samples, write_rfilter)
 *   consumer   radio.c:1460      (execute_filter_output(&chan->filter.out, shift))
 *   channel    fm.c:27-34, radio.c:1559-1611  (create_filter_output, set_filter)
 * so the rest of the oracle (and the GPU parity tests) can be pinned against the real thing.
 */
#define _GNU_SOURCE 1
#include <assert.h>
#include <complex.h>
#include <math.h>
#include <pthread.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "filter.h" /* the reference's header, found through -iquote /root/reference/src */
#include "osc.h"
#include "airspy.h"

int Verbose = 0; /* referenced by misc.c */
char const *App_path = "ka9q-oracle";

extern _Thread_local bool Rand_init; /* gauss.c:15 */

struct ref_session {
  struct filter_in in;
  int nchan, cap;
  struct filter_out **out;
  struct notch_state *notches;
};

/* nworkers = 0 => inline forward FFT (deterministic; what the parity tests use). */
struct ref_session *ref_open(int L, int M, int in_type, int nworkers) {
  struct ref_session *s = calloc(1, sizeof *s);
  if (!s)
    return NULL;
  N_worker_threads = nworkers; /* owned by filter.c:44, set by radio.c:303 */
  if (create_filter_input(&s->in, L, M, (enum filtertype)in_type) != 0) {
    free(s);
    return NULL;
  }
  return s;
}

/* bins[] as radio.c:608-620 builds it: caller-listed spur bins, DC entry (bin 0) appended last. */
int ref_set_notches(struct ref_session *s, int const *bins, int nbins, double alpha) {
  free(s->notches);
  s->notches = calloc((size_t)nbins + 1, sizeof *s->notches);
  for (int i = 0; i < nbins; i++) {
    s->notches[i].bin = bins[i];
    s->notches[i].alpha = alpha;
  }
  s->notches[nbins].bin = 0;
  s->notches[nbins].alpha = alpha;
  s->in.notches = s->notches;
  return 0;
}

int ref_add_channel(struct ref_session *s, int olen, int out_type, double low, double high, double beta) {
  if (s->nchan == s->cap) {
    s->cap = s->cap ? 2 * s->cap : 64;
    s->out = realloc(s->out, sizeof(*s->out) * (size_t)s->cap);
  }
  struct filter_out *o = calloc(1, sizeof *o);
  if (create_filter_output(o, &s->in, olen, (enum filtertype)out_type) != 0) {
    free(o);
    return -1;
  }
  if (out_type != SPECTRUM && set_filter(o, low, high, beta) != 0) {
    delete_filter_output(o);
    free(o);
    return -1;
  }
  s->out[s->nchan] = o;
  return s->nchan++;
}
int ref_retune_channel(struct ref_session *s, int ch, double low, double high, double beta) {
  return set_filter(s->out[ch], low, high, beta);
}
int ref_channel_points(struct ref_session *s, int ch) { return s->out[ch]->points; }
int ref_get_response(struct ref_session *s, int ch, float complex *dst) {
  memcpy(dst, s->out[ch]->response, sizeof(float complex) * (size_t)s->out[ch]->bins);
  return s->out[ch]->bins;
}
int ref_set_isb(struct ref_session *s, int ch, int isb) {
  s->out[ch]->isb = isb != 0;
  return 0;
}
int ref_set_beam(struct ref_session *s, int ch, int beam, double ire, double iim, double qre, double qim) {
  s->out[ch]->beam = beam != 0;
  return set_filter_weights(s->out[ch], CMPLX(ire, iim), CMPLX(qre, qim));
}

/* Copy n samples in and fire blocks exactly as write_rfilter/write_cfilter do. */
int ref_write_real(struct ref_session *s, float const *x, int n) { return write_rfilter(&s->in, x, n); }
int ref_write_complex(struct ref_session *s, float complex const *x, int n) { return write_cfilter(&s->in, x, n); }

/* Spectrum of the most recently completed block (inline mode). */
int ref_get_spectrum(struct ref_session *s, float complex *dst) {
  unsigned const job = s->in.next_jobnum - 1;
  memcpy(dst, s->in.fdomain[job % ND], sizeof(float complex) * (size_t)s->in.bins);
  return s->in.bins;
}
int ref_master_bins(struct ref_session *s) { return s->in.bins; }

/* One channel, one block.  dst receives the user-visible output (olen samples); full (if not
 * NULL) the whole Ns-point inverse transform; fdom (if not NULL) the slave's frequency domain
 * product after slicing (filter.c:728-911). Returns execute_filter_output()'s value. */
int ref_execute_channel(struct ref_session *s, int ch, int shift, float complex *dst, float complex *full,
                        float complex *fdom) {
  struct filter_out *o = s->out[ch];
  int const r = execute_filter_output(o, shift);
  if (o->out_type == COMPLEX) {
    if (dst)
      memcpy(dst, o->output.c, sizeof(float complex) * (size_t)o->olen);
    if (full)
      memcpy(full, o->output_buffer.c, sizeof(float complex) * (size_t)o->points);
  } else if (o->out_type == REAL) {
    if (dst)
      memcpy(dst, o->output.r, sizeof(float) * (size_t)o->olen);
    if (full)
      memcpy(full, o->output_buffer.r, sizeof(float) * (size_t)o->points);
  }
  if (fdom && o->fdomain)
    memcpy(fdom, o->fdomain, sizeof(float complex) * (size_t)o->bins);
  return r;
}
unsigned ref_channel_drops(struct ref_session *s, int ch) { return s->out[ch]->block_drops; }

void ref_close(struct ref_session *s) {
  if (!s)
    return;
  for (int i = 0; i < s->nchan; i++) {
    delete_filter_output(s->out[i]);
    free(s->out[i]);
  }
  free(s->out);
  s->in.notches = NULL;
  /* worker threads (if any) are detached and idle; the master can go */
  delete_filter_input(&s->in);
  free(s->notches);
  free(s);
}

/* ---- the reference's own synthetic source: sig_gen.c:288-296 (real) / :318-322 (complex) ---- */
/* x[i] = (amplitude*Re(step_osc) + noise*real_gauss()) * scale, xoshiro seed 1 (gauss.c:95-100). */
void ref_siggen_real(float *dst, long n, double amplitude, double noise, double cycles_per_sample, double scale,
                     int reseed) {
  if (reseed)
    Rand_init = false;
  rand_init();
  struct osc carrier = {0};
  set_osc(&carrier, cycles_per_sample, 0.0);
  for (long i = 0; i < n; i++) {
    double const samp = amplitude * creal(step_osc(&carrier)) + noise * real_gauss();
    dst[i] = (float)(samp * scale);
  }
}
void ref_siggen_complex(float complex *dst, long n, double amplitude, double noise, double cycles_per_sample,
                        double scale, int reseed) {
  if (reseed)
    Rand_init = false;
  rand_init();
  struct osc carrier = {0};
  set_osc(&carrier, cycles_per_sample, 0.0);
  for (long i = 0; i < n; i++) {
    double complex const samp = amplitude * step_osc(&carrier) + noise * complex_gauss();
    dst[i] = (float complex)(samp * scale);
  }
}

/* ---- CPU baseline: run the reference the way radiod does, timed ------------------------------ */
/* One producer (this thread) calling write_rfilter(in,NULL,L) after filling the ring in place
 * (rx888.c:800-826), `nworkers` run_fft threads (filter.c:485), one pthread per channel looping
 * execute_filter_output (radio.c:996,1460).  `input` holds `nin` float samples and is replayed
 * cyclically.  For COMPLEX masters `input` is interleaved I/Q (nin complex samples).
 * Returns elapsed seconds for nblocks blocks (after `warm` untimed blocks). */
struct chan_arg {
  struct filter_out *o;
  int shift;
  int nblocks;
  volatile double sink;
};
static void *chan_thread(void *p) {
  struct chan_arg *a = p;
  double acc = 0;
  for (int b = 0; b < a->nblocks; b++) {
    execute_filter_output(a->o, a->shift);
    acc += crealf(a->o->output.c[0]);
  }
  a->sink = acc;
  return NULL;
}
static double ref_bench_run(struct ref_session *s, int L, int in_type, int nchan, int const *shifts, void const *input, long nin,
                            int nblocks, int nworkers, unsigned *drops_out);
double ref_bench(int L, int M, int in_type, int nchan, int olen, int const *shifts, double low, double high,
                 double beta, void const *input, long nin, int nblocks, int nworkers, unsigned *drops_out) {
  struct ref_session *s = ref_open(L, M, in_type, nworkers);
  if (!s)
    return -1;
  for (int i = 0; i < nchan; i++)
    if (ref_add_channel(s, olen, COMPLEX, low, high, beta) < 0)
      return -1;
  return ref_bench_run(s, L, in_type, nchan, shifts, input, nin, nblocks, nworkers, drops_out);
}
/* same with per-channel output lengths and filters (mixed-rate banks, BASELINE cfg-3) */
double ref_bench_mixed(int L, int M, int in_type, int nchan, int const *olen, int const *shifts, double const *low,
                       double const *high, double const *beta, void const *input, long nin, int nblocks, int nworkers,
                       unsigned *drops_out) {
  struct ref_session *s = ref_open(L, M, in_type, nworkers);
  if (!s)
    return -1;
  for (int i = 0; i < nchan; i++)
    if (ref_add_channel(s, olen[i], COMPLEX, low[i], high[i], beta[i]) < 0)
      return -1;
  return ref_bench_run(s, L, in_type, nchan, shifts, input, nin, nblocks, nworkers, drops_out);
}
static double ref_bench_run(struct ref_session *s, int L, int in_type, int nchan, int const *shifts, void const *input, long nin,
                            int nblocks, int nworkers, unsigned *drops_out) {
  struct chan_arg *args = calloc((size_t)nchan, sizeof *args);
  pthread_t *tids = calloc((size_t)nchan, sizeof *tids);
  pthread_attr_t attr;
  pthread_attr_init(&attr);
  pthread_attr_setstacksize(&attr, 256 * 1024);
  for (int i = 0; i < nchan; i++) {
    args[i].o = s->out[i];
    args[i].shift = shifts[i];
    args[i].nblocks = nblocks;
    s->out[i]->next_jobnum = s->in.next_jobnum;
  }
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int i = 0; i < nchan; i++)
    pthread_create(&tids[i], &attr, chan_thread, &args[i]);
  long pos = 0;
  size_t const esz = (in_type == COMPLEX) ? sizeof(float complex) : sizeof(float);
  for (int b = 0; b < nblocks; b++) {
    /* producer writes straight into the ring like a front-end driver, then publishes */
    long left = L;
    char *w = (in_type == COMPLEX) ? (char *)s->in.input_write_pointer.c : (char *)s->in.input_write_pointer.r;
    while (left > 0) {
      long const chunk = (nin - pos < left) ? nin - pos : left;
      memcpy(w, (char const *)input + (size_t)pos * esz, (size_t)chunk * esz);
      w += (size_t)chunk * esz;
      pos = (pos + chunk) % nin;
      left -= chunk;
    }
    if (in_type == COMPLEX)
      write_cfilter(&s->in, NULL, L);
    else
      write_rfilter(&s->in, NULL, L);
    /* never run more than ND-1 blocks ahead of the slowest consumer: a real front end is paced by
     * its ADC; an unpaced producer would just lap the ring and make every channel drop blocks */
    if (nworkers > 0 || nchan > 0) {
      for (;;) {
        unsigned minjob = s->in.next_jobnum;
        for (int i = 0; i < nchan; i++) {
          unsigned const j = *(volatile unsigned *)&s->out[i]->next_jobnum;
          if ((int)(j - minjob) < 0)
            minjob = j;
        }
        if ((int)(s->in.next_jobnum - minjob) < ND - 1)
          break;
        struct timespec ts = {0, 20000};
        nanosleep(&ts, NULL);
      }
    }
  }
  for (int i = 0; i < nchan; i++)
    pthread_join(tids[i], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  unsigned drops = 0;
  for (int i = 0; i < nchan; i++)
    drops += s->out[i]->block_drops;
  if (drops_out)
    *drops_out = drops;
  free(args);
  free(tids);
  /* let detached worker threads go idle before tearing the master down */
  struct timespec ts = {0, 50 * 1000 * 1000};
  nanosleep(&ts, NULL);
  ref_close(s);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* the reference's own unpackers (airspy-unpack.c), both variants, for pinning ko_airspy_unpack */
int ref_airspy_unpack(float *dst, uint32_t const *up, int sampcount, float scale, uint64_t *energy, int avx2) {
#if defined(__x86_64__)
  if (avx2 && __builtin_cpu_supports("avx2"))
    return airspy_unpack_avx2(dst, up, sampcount, scale, energy);
#endif
  (void)avx2;
  return airspy_unpack(dst, up, sampcount, scale, energy);
}
