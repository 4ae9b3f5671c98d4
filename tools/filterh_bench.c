/* tools/filterh_bench.c -- bench.py's "e2e_filter_h" leg: the hot path driven ONLY through the reference-facing
 * filter.h symbols of libka9qgpu.so (create_filter_input/output, set_filter, write_i16filter,
 * execute_filter_output_batch), the way a patched radiod would: one producer thread handing raw int16 ADC words
 * from HOST memory to the master (rx888.c:797-826 -> write_i16filter), one consumer thread standing in for the channel
 * threads (radio.c:996,1460).  Everything between the host sample buffer and the host output buffers is inside the
 * timed region: ring copy, H2D, kernels, D2H, per-slave delivery.  BENCH INFRASTRUCTURE, not product, not oracle.
 */
#define _GNU_SOURCE 1
#include <complex.h>
#include <pthread.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include "ka9q_gpu_filter.h"

static double now_s(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

struct run {
  struct filter_in in;
  struct filter_out *out;
  struct filter_out **outp;
  int *shifts;
  int nchan, k, nblocks, warm, inplace;
  int16_t const *stream;
  long stream_blocks; /* blocks available in `stream`, replayed cyclically */
  long words_per_block;
  float scale;
  double *t_write, *t_done; /* per block: producer hand-over time, consumer completion time */
  volatile double sink;
};

static void *producer(void *p) {
  struct run *r = p;
  int const total = r->warm + r->nblocks;
  long pos = 0;
  for (int b = 0; b < total;) {
    int k = r->k;
    if (k > total - b)
      k = total - b;
    if (!r->inplace && pos + k > r->stream_blocks)
      pos = 0;
    /* ADC pacing stand-in: never overwrite a ring slot a slave has not consumed (an unpaced producer would only
     * make every channel drop blocks, filter.c:690-701) */
    for (;;) {
      unsigned const need = r->in.next_jobnum + (unsigned)k;
      unsigned const c0 = *(volatile unsigned *)&r->out[0].next_jobnum;
      unsigned const c1 = *(volatile unsigned *)&r->out[r->nchan - 1].next_jobnum;
      unsigned const low = (int)(c0 - c1) < 0 ? c0 : c1;
      if ((int)(need - low) <= ND - 1) /* one slot of slack: in zero-copy mode the consumer still reads the row it was handed */
        break;
      struct timespec ts = {0, 5000};
      nanosleep(&ts, NULL);
    }
    double const t = now_s();
    for (int j = 0; j < k; j++)
      r->t_write[b + j] = t;
    /* inplace: the samples are already in the pinned ring (a driver's DMA target, filter_i16_write_pointer): publish only */
    write_i16filter(&r->in, r->inplace ? NULL : r->stream + pos * r->words_per_block,
                    (int)(k * (r->words_per_block / (r->in.in_type == COMPLEX ? 2 : 1))), r->scale, false);
    pos += k;
    b += k;
  }
  return NULL;
}

static void *consumer(void *p) {
  struct run *r = p;
  int const total = r->warm + r->nblocks;
  double acc = 0;
  for (int b = 0; b < total; b++) {
    execute_filter_output_batch(r->outp, r->shifts, r->nchan);
    for (int i = 0; i < r->nchan; i += 37) /* touch the delivered samples */
      acc += crealf(r->out[i].output.c[0]) + cimagf(r->out[i].output.c[r->out[i].olen - 1]);
    r->t_done[b] = now_s();
  }
  r->sink = acc;
  return NULL;
}

/* int16 words in the library's raw-ingest ring: ND windows, rounded up to whole pages (filter_abi.c: page_round) */
long kgf_ring_words(int L, int M, int in_type) {
  size_t const esz = (in_type == COMPLEX) ? 2 * sizeof(int16_t) : sizeof(int16_t);
  size_t const pg = (size_t)sysconf(_SC_PAGESIZE);
  size_t const bytes = ((size_t)ND * (size_t)(L + M - 1) * esz + pg - 1) / pg * pg;
  return (long)(bytes / sizeof(int16_t));
}

/* Returns seconds for `nblocks` blocks (after `warm` untimed ones), < 0 on error.
 * check_out (nchan * max_olen complex) receives every slave's output of the LAST block; *last_stream_block its index in
 * `stream`; lat_ms[2] = mean and max hand-over -> delivered latency per block; *drops the dropped blocks. */
double kgf_e2e_run(int L, int M, int in_type, int nchan, int const *olen, int const *shifts, double const *low, double const *high,
                   double const *beta, int16_t const *stream, long stream_blocks, int blocks_per_write, int warm, int nblocks,
                   float scale, float complex *check_out, int max_olen, long *last_stream_block, double *lat_ms, unsigned *drops,
                   int inplace) {
  struct run r;
  memset(&r, 0, sizeof r);
  N_worker_threads = 1; /* not inline: producer and consumers are different threads */
  if (blocks_per_write < 1 || blocks_per_write > ND - 1 || create_filter_input(&r.in, L, M, (enum filtertype)in_type) != 0)
    return -1;
  r.out = calloc((size_t)nchan, sizeof *r.out);
  r.outp = calloc((size_t)nchan, sizeof *r.outp);
  r.shifts = calloc((size_t)nchan, sizeof *r.shifts);
  for (int i = 0; i < nchan; i++) {
    r.outp[i] = &r.out[i];
    r.shifts[i] = shifts[i];
    if (create_filter_output(&r.out[i], &r.in, olen[i], COMPLEX) != 0 || set_filter(&r.out[i], low[i], high[i], beta[i]) != 0)
      return -2;
  }
  r.nchan = nchan;
  r.k = blocks_per_write;
  r.nblocks = nblocks;
  r.warm = warm;
  r.stream = stream;
  r.stream_blocks = stream_blocks;
  r.words_per_block = (long)L * (in_type == COMPLEX ? 2 : 1);
  r.scale = scale;
  r.inplace = 0;
  long const ring_words = kgf_ring_words(L, M, in_type);
  if (inplace) {
    /* Fill the whole pinned ring once (ring_words int16 words, starting at the write pointer and running through the
     * mirror mapping at the ring's end), as a front end's DMA would keep doing, and let the producer publish the samples
     * block by block without touching them again: ring[(hist + i) mod R] = stream[i], so block b's window starts at
     * stream word (b * words_per_block - hist) mod R. */
    int16_t *w = filter_i16_write_pointer(&r.in);
    if (w == NULL || stream_blocks * r.words_per_block < ring_words)
      return -3;
    memcpy(w, stream, sizeof(int16_t) * (size_t)ring_words);
    r.inplace = 1;
  }
  r.t_write = calloc((size_t)(warm + nblocks), sizeof(double));
  r.t_done = calloc((size_t)(warm + nblocks), sizeof(double));
  for (int i = 0; i < nchan; i++)
    r.out[i].next_jobnum = r.in.next_jobnum;
  pthread_t tp, tc;
  pthread_create(&tc, NULL, consumer, &r);
  pthread_create(&tp, NULL, producer, &r);
  pthread_join(tp, NULL);
  pthread_join(tc, NULL);
  double const secs = r.t_done[warm + nblocks - 1] - (warm ? r.t_done[warm - 1] : r.t_write[0]);
  double sum = 0, mx = 0;
  for (int b = warm; b < warm + nblocks; b++) {
    double const d = r.t_done[b] - r.t_write[b];
    sum += d;
    if (d > mx)
      mx = d;
  }
  if (lat_ms) {
    lat_ms[0] = 1e3 * sum / nblocks;
    lat_ms[1] = 1e3 * mx;
  }
  unsigned dr = 0;
  for (int i = 0; i < nchan; i++) {
    dr += r.out[i].block_drops;
    if (check_out)
      memcpy(check_out + (size_t)i * (size_t)max_olen, r.out[i].output.c, sizeof(float complex) * (size_t)r.out[i].olen);
  }
  if (drops)
    *drops = dr;
  if (last_stream_block) { /* replay the producer's cursor */
    long pos = 0, last = 0;
    int const total = warm + nblocks;
    for (int b = 0; b < total;) {
      int k = blocks_per_write;
      if (k > total - b)
        k = total - b;
      if (!inplace && pos + k > stream_blocks)
        pos = 0;
      last = pos + k - 1;
      pos += k;
      b += k;
    }
    if (inplace) { /* word index (into `stream`, cyclic with period ring_words) of the last block's window start */
      long const hist = (long)(M - 1) * (in_type == COMPLEX ? 2 : 1);
      long st = ((long)(total - 1) * r.words_per_block - hist) % ring_words;
      if (st < 0)
        st += ring_words;
      last = st;
    }
    *last_stream_block = last;
  }
  for (int i = 0; i < nchan; i++)
    delete_filter_output(&r.out[i]);
  delete_filter_input(&r.in);
  free(r.out);
  free(r.outp);
  free(r.shifts);
  free(r.t_write);
  free(r.t_done);
  return secs;
}
