/* filter_abi.c -- the reference's filter.h surface served by the B200 kernels.
 *
 * Host code stays C, exactly as north_star asks: this file owns the caller-visible state the
 * untouched ka9q-radio sources expect (mirrored input ring, ND-deep job bookkeeping, per-slave
 * output buffers, mutex/condvar hand-off) and forwards the arithmetic to the C-ABI in
 * include/ka9q_gpu.h.  One master = one device pipeline on one stream:
 *
 *   execute_filter_input  : H2D(window) -> fwd_cols -> fwd_rows -> notch -> chan (ALL slaves,
 *                           batched with their last shifts) -> D2H(outputs) [-> D2H(spectrum)]
 *   execute_filter_output : wait for that block's event, copy the slave's slice out of the
 *                           pinned batch buffer; a slave whose shift/filter changed since the
 *                           block was issued is recomputed alone (retunes are rare, radio.c:1491)
 *
 * The two FFTW plan slots of the reference structs (fwd_plan / rev_plan, only ever touched by
 * filter.c itself) carry the contexts.  Reference lines are cited per function.
 */
#define _GNU_SOURCE 1
#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <errno.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>

#include "ka9q_gpu.h"
#include "ka9q_gpu_filter.h"

/* globals filter.c owns in the reference (filter.c:40-48, :476-479) */
char const *Wisdom_file;
int N_worker_threads = 1;
int N_internal_threads = 1;
int FFTW_planning_level = (1 << 5); /* FFTW_PATIENT; meaningless here, kept for radio.c:310-318 */
int64_t Min_fft_time = INT64_MAX;
int64_t Max_fft_time = 0;
int64_t Avg_fft_time = 0;
int64_t Mean_dev = 0;

#define KGF_MAX_SLAVES 2048 /* radiod allows 2000 channels (radio.h:356) */

struct slave_ctx {
  int idx; /* bank slot */
  unsigned version;
};

struct master_ctx {
  kgpu_master *km;
  kgpu_bank *bank;
  cudaStream_t st, st_one;
  size_t esz; /* bytes per input element (float: 4, float complex: 8) */
  void *d_win[ND];
  float complex *d_spec; /* ND * spec_stride */
  long spec_stride;
  float complex *d_out[ND], *h_out[ND];
  long out_cap; /* float2 capacity of each d_out/h_out */
  float complex *d_one, *h_one;
  int one_cap;
  cudaEvent_t t0[ND], done[ND];
  bool timed[ND];
  pthread_mutex_t mu;
  struct filter_out *slots[KGF_MAX_SLAVES];
  unsigned ver[KGF_MAX_SLAVES];
  int cur_shift[KGF_MAX_SLAVES]; /* shift / isb last pushed into the bank */
  bool cur_isb[KGF_MAX_SLAVES];
  /* what the batched launch of each ring slot used */
  int snap_shift[ND][KGF_MAX_SLAVES];
  unsigned snap_ver[ND][KGF_MAX_SLAVES];
  long snap_off[ND][KGF_MAX_SLAVES];
  bool snap_isb[ND][KGF_MAX_SLAVES];
  bool snap_ok[ND][KGF_MAX_SLAVES];
  int nslots; /* highest used + 1 */
  bool spectrum_d2h;
  struct notch_state *notches_seen;
  /* raw int16 ingest (extension) */
  bool i16_mode;
  void *i16_ring;
  size_t i16_ring_size, i16_esz;
  char *i16_wp, *i16_rp;
  float i16_scale;
  bool i16_derand;
};

/* ---------------------------------------------------------------- mirrored ring ------------- */
/* Same contract as the reference's mirror_alloc (misc.c:635-682): `size` bytes followed by a
 * second mapping of the same pages, so a window that starts near the end stays contiguous. */
static size_t page_round(size_t n) {
  size_t const pg = (size_t)sysconf(_SC_PAGESIZE);
  return (n + pg - 1) / pg * pg;
}
static void *ring_alloc(size_t size) {
  int const fd = memfd_create("ka9q-gpu-ring", 0);
  if (fd < 0)
    return NULL;
  if (ftruncate(fd, (off_t)size) != 0) {
    close(fd);
    return NULL;
  }
  char *base = mmap(NULL, 2 * size, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (base == MAP_FAILED) {
    close(fd);
    return NULL;
  }
  void *a = mmap(base, size, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED, fd, 0);
  void *b = mmap(base + size, size, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED, fd, 0);
  close(fd);
  if (a == MAP_FAILED || b == MAP_FAILED) {
    munmap(base, 2 * size);
    return NULL;
  }
  memset(base, 0, size);
  /* pin the primary view so the per-block H2D is a true async DMA (the copy never reads through
   * the mirror view, see window_h2d); harmless if the driver refuses */
  if (cudaHostRegister(base, size, cudaHostRegisterPortable) != cudaSuccess)
    (void)cudaGetLastError();
  return base;
}
static void ring_free(void *base, size_t size) {
  if (!base)
    return;
  if (cudaHostUnregister(base) != cudaSuccess)
    (void)cudaGetLastError();
  munmap(base, 2 * size);
}

static int kgf_fail(char const *where) {
  fprintf(stderr, "ka9q-gpu filter: %s failed: kgpu: \"%s\" cuda: %s\n", where, kgpu_last_error(),
          cudaGetErrorString(cudaGetLastError()));
  return -1;
}

/* H2D of one FFT window that starts inside the primary view and may run past its end: the part
 * beyond the end is the start of the ring again (that is what the mirror view shows the CPU). */
static int window_h2d(void *dst, void const *src, size_t bytes, void const *ring, size_t ring_size, cudaStream_t st) {
  char const *end = (char const *)ring + ring_size;
  if ((char const *)src + bytes <= end)
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st) == cudaSuccess ? 0 : -1;
  size_t const first = (size_t)(end - (char const *)src);
  if (cudaMemcpyAsync(dst, src, first, cudaMemcpyHostToDevice, st) != cudaSuccess)
    return -1;
  return cudaMemcpyAsync((char *)dst + first, ring, bytes - first, cudaMemcpyHostToDevice, st) == cudaSuccess ? 0 : -1;
}

static void *cache_aligned(size_t bytes) {
  void *p = NULL;
  return posix_memalign(&p, 64, bytes ? bytes : 64) == 0 ? p : NULL;
}

static void master_teardown(struct filter_in *master) {
  struct master_ctx *c = (struct master_ctx *)master->fwd_plan;
  if (c) {
    cudaStreamSynchronize(c->st);
    cudaStreamSynchronize(c->st_one);
    for (int i = 0; i < ND; i++) {
      cudaFree(c->d_win[i]);
      cudaFree(c->d_out[i]);
      cudaFreeHost(c->h_out[i]);
      cudaEventDestroy(c->t0[i]);
      cudaEventDestroy(c->done[i]);
    }
    cudaFree(c->d_spec);
    cudaFree(c->d_one);
    cudaFreeHost(c->h_one);
    kgpu_bank_destroy(c->bank);
    kgpu_master_destroy(c->km);
    cudaStreamDestroy(c->st);
    cudaStreamDestroy(c->st_one);
    ring_free(c->i16_ring, c->i16_ring_size);
    pthread_mutex_destroy(&c->mu);
    free(c);
    master->fwd_plan = NULL;
  }
  for (int i = 0; i < ND; i++) {
    if (master->fdomain[i])
      cudaFreeHost(master->fdomain[i]);
    master->fdomain[i] = NULL;
  }
  ring_free(master->input_buffer, master->input_buffer_size);
  master->input_buffer = NULL;
}

/* ---------------------------------------------------------------- create_filter_input ------- */
/* filter.c:186-269 */
int create_filter_input(struct filter_in *master, int const L, int const M, enum filtertype const in_type) {
  if (master == NULL || L <= 0 || M <= 0)
    return -1;
  if (master->init && master->ilen == L && master->impulse_length == M && in_type == master->in_type)
    return 0; /* unchanged (filter.c:191) */
  if (in_type != REAL && in_type != COMPLEX)
    return -1;
  int const N = L + M - 1;
  int const bins = (in_type == COMPLEX) ? N : N / 2 + 1;
  if (bins < 2)
    return -1;
  if (master->init && master->fwd_plan)
    master_teardown(master);

  struct master_ctx *c = calloc(1, sizeof *c);
  if (!c)
    return -1;
  c->km = kgpu_master_create(L, M, in_type == REAL ? KGPU_REAL : KGPU_COMPLEX);
  if (!c->km) {
    fprintf(stderr, "create_filter_input(L=%d M=%d): %s\n", L, M, kgpu_last_error());
    free(c);
    return -1;
  }
  c->bank = kgpu_bank_create(c->km, KGF_MAX_SLAVES);
  c->esz = (in_type == COMPLEX) ? sizeof(float complex) : sizeof(float);
  c->spec_stride = kgpu_master_spec_stride(c->km);
  char const *env = getenv("KA9Q_GPU_SPECTRUM_D2H");
  c->spectrum_d2h = !(env && env[0] == '0'); /* radio.c:1799-1831 reads master->fdomain on the host */
  pthread_mutex_init(&c->mu, NULL);
  bool ok = c->bank != NULL;
  ok = ok && cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&c->st_one, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaMalloc((void **)&c->d_spec, sizeof(float complex) * (size_t)c->spec_stride * ND) == cudaSuccess;
  for (int i = 0; ok && i < ND; i++) {
    ok = ok && cudaMalloc(&c->d_win[i], c->esz * (size_t)N) == cudaSuccess;
    ok = ok && cudaEventCreate(&c->t0[i]) == cudaSuccess;
    ok = ok && cudaEventCreate(&c->done[i]) == cudaSuccess;
  }
  master->points = N;
  master->perform_inline = (N_worker_threads == 0);
  master->bins = bins;
  master->ilen = L;
  master->impulse_length = M;
  master->in_type = in_type;
  master->wcnt = 0;
  master->next_jobnum = 0;
  master->sample_index = 0;
  for (int i = 0; ok && i < ND; i++) {
    ok = ok && cudaHostAlloc((void **)&master->fdomain[i], sizeof(float complex) * (size_t)bins, cudaHostAllocPortable) ==
                   cudaSuccess;
    master->completed_jobs[i] = UINT_MAX; /* filter.c:214 */
  }
  master->input_buffer_size = page_round((size_t)ND * N * c->esz);
  master->input_buffer = ok ? ring_alloc(master->input_buffer_size) : NULL;
  ok = ok && master->input_buffer != NULL;
  master->fwd_plan = (fftwf_plan)c;
  if (!ok) {
    fprintf(stderr, "create_filter_input(L=%d M=%d): device/host allocation failed: %s\n", L, M,
            cudaGetErrorString(cudaGetLastError()));
    master_teardown(master);
    return -1;
  }
  /* read pointer at the start, write pointer M-1 samples in: the zero history (filter.c:243-244) */
  if (in_type == COMPLEX) {
    master->input_read_pointer.c = master->input_buffer;
    master->input_write_pointer.c = master->input_read_pointer.c + (M - 1);
    master->input_read_pointer.r = master->input_write_pointer.r = NULL;
  } else {
    master->input_read_pointer.r = master->input_buffer;
    master->input_write_pointer.r = master->input_read_pointer.r + (M - 1);
    master->input_read_pointer.c = master->input_write_pointer.c = NULL;
  }
  if (!master->init) {
    pthread_mutex_init(&master->filter_mutex, NULL);
    pthread_cond_init(&master->filter_cond, NULL);
    master->init = true;
  }
  master->owner = pthread_self();
  return 0;
}

/* ---------------------------------------------------------------- create_filter_output ------ */
/* filter.c:298-415 */
int create_filter_output(struct filter_out *slave, struct filter_in *master, int len, enum filtertype out_type) {
  if (master == NULL || slave == NULL || (out_type != SPECTRUM && len <= 0) || master->fwd_plan == NULL)
    return -1;
  if (slave->master == master && slave->olen == len && slave->out_type == out_type && slave->init)
    goto done;
  if (out_type == REAL) {
    fprintf(stderr, "create_filter_output: REAL output slaves are not served by the GPU backend\n");
    return -1;
  }
  if (out_type == SPECTRUM)
    len = 0;
  struct master_ctx *c = (struct master_ctx *)master->fwd_plan;
  int const N = master->ilen + master->impulse_length - 1, L = master->ilen;
  if (((long)len * N % L) != 0) {
    fprintf(stderr, "Invalid filter output length %d for input N=%d, L=%d\n", len, N, L);
    return -1;
  }
  if (!slave->init) {
    pthread_mutex_init(&slave->response_mutex, NULL);
    slave->init = true;
  } else {
    pthread_mutex_lock(&slave->response_mutex);
    free(slave->response);
    slave->response = NULL;
    pthread_mutex_unlock(&slave->response_mutex);
    free(slave->fdomain);
    slave->fdomain = NULL;
    free(slave->output_buffer.c);
    slave->output_buffer.c = NULL;
    slave->output.c = NULL;
  }
  slave->olen = len;
  slave->points = (int)((long)len * N / L);
  slave->master = master;
  slave->out_type = out_type;
  set_filter_weights(slave, 1.0, 0.0);
  struct slave_ctx *sc = (struct slave_ctx *)slave->rev_plan;
  if (out_type == COMPLEX) {
    pthread_mutex_lock(&c->mu);
    if (!sc) {
      int idx = -1;
      for (int i = 0; i < KGF_MAX_SLAVES; i++)
        if (c->slots[i] == NULL) {
          idx = i;
          break;
        }
      if (idx < 0) {
        pthread_mutex_unlock(&c->mu);
        return -1;
      }
      sc = calloc(1, sizeof *sc);
      sc->idx = idx;
      c->slots[idx] = slave;
      if (idx + 1 > c->nslots)
        c->nslots = idx + 1;
      slave->rev_plan = (fftwf_plan)sc;
    }
    cudaStreamSynchronize(c->st);
    int const pts = kgpu_bank_define(c->bank, sc->idx, len);
    c->ver[sc->idx]++;
    pthread_mutex_unlock(&c->mu);
    if (pts != slave->points) {
      fprintf(stderr, "create_filter_output: %s\n", kgpu_last_error());
      return -1;
    }
    slave->bins = slave->points;
    slave->fdomain = cache_aligned(sizeof(float complex) * (size_t)slave->bins);
    slave->output_buffer.c = cache_aligned(sizeof(float complex) * (size_t)slave->points);
    if (!slave->fdomain || !slave->output_buffer.c)
      return -1;
    memset(slave->output_buffer.c, 0, sizeof(float complex) * (size_t)slave->points);
    slave->output.c = slave->output_buffer.c + slave->bins - len; /* filter.c:357 */
  }
done:;
  slave->next_jobnum = master->next_jobnum;
  return 0;
}

/* ---------------------------------------------------------------- execute_filter_input ------ */
static int grow_out(struct master_ctx *c, long need) {
  if (need <= c->out_cap)
    return 0;
  cudaStreamSynchronize(c->st);
  long const cap = need + need / 4 + 1024;
  for (int i = 0; i < ND; i++) {
    cudaFree(c->d_out[i]);
    cudaFreeHost(c->h_out[i]);
    c->d_out[i] = NULL;
    c->h_out[i] = NULL;
    if (cudaMalloc((void **)&c->d_out[i], sizeof(float complex) * (size_t)cap) != cudaSuccess ||
        cudaHostAlloc((void **)&c->h_out[i], sizeof(float complex) * (size_t)cap, cudaHostAllocPortable) != cudaSuccess)
      return -1;
    memset(c->snap_ok[i], 0, sizeof c->snap_ok[i]);
  }
  c->out_cap = cap;
  return 0;
}

static void sync_notches(struct filter_in *f, struct master_ctx *c) {
  if (f->notches == c->notches_seen)
    return;
  c->notches_seen = f->notches;
  int bins[64];
  double alpha[64];
  int n = 0;
  if (f->notches)
    for (struct notch_state *p = f->notches; n < 64; p++) { /* list ends at bin 0 (filter.c:470) */
      bins[n] = p->bin;
      alpha[n] = p->alpha;
      n++;
      if (p->bin == 0)
        break;
    }
  cudaStreamSynchronize(c->st);
  kgpu_master_set_notches(c->km, bins, alpha, n);
}

/* filter.c:558-651 (+ run_fft :485-555) */
int execute_filter_input(struct filter_in *const f) {
  if (f == NULL || f->fwd_plan == NULL)
    return -1;
  struct master_ctx *c = (struct master_ctx *)f->fwd_plan;
  int const N = f->points;
  pthread_mutex_lock(&c->mu);
  unsigned const jobnum = f->next_jobnum;
  int const slot = (int)(jobnum % ND);
  /* the ring slot's previous occupant (job - ND) must have drained */
  cudaEventSynchronize(c->done[slot]);
  if (c->timed[slot]) { /* forward+channels device time of that older job, for main.c:154-164 */
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->t0[slot], c->done[slot]) == cudaSuccess) {
      int64_t const ns = (int64_t)(ms * 1e6f);
      if (ns > Max_fft_time)
        Max_fft_time = ns;
      if (ns < Min_fft_time)
        Min_fft_time = ns;
      int64_t const dev = ns - Avg_fft_time;
      Avg_fft_time += dev >> 4;
      Mean_dev += (llabs(dev) - Mean_dev) >> 4;
    }
  }
  sync_notches(f, c);
  int rc = 0;
  cudaEventRecord(c->t0[slot], c->st);
  void const *src;
  int fmt = KGPU_FMT_F32;
  float scale = 1.0f;
  size_t bytes;
  if (c->i16_mode) {
    src = c->i16_rp;
    bytes = c->i16_esz * (size_t)N;
    fmt = KGPU_FMT_I16;
    scale = c->i16_scale;
    c->i16_rp += c->i16_esz * (size_t)f->ilen;
    if (c->i16_rp >= (char *)c->i16_ring + c->i16_ring_size)
      c->i16_rp -= c->i16_ring_size;
  } else if (f->in_type == COMPLEX) {
    src = f->input_read_pointer.c;
    bytes = sizeof(float complex) * (size_t)N;
    f->input_read_pointer.c += f->ilen;
    kgf_ring_wrap((void **)&f->input_read_pointer.c, f->input_buffer, f->input_buffer_size);
  } else {
    src = f->input_read_pointer.r;
    bytes = sizeof(float) * (size_t)N;
    f->input_read_pointer.r += f->ilen;
    kgf_ring_wrap((void **)&f->input_read_pointer.r, f->input_buffer, f->input_buffer_size);
  }
  float complex *spec = c->d_spec + (size_t)slot * (size_t)c->spec_stride;
  if (window_h2d(c->d_win[slot], src, bytes, c->i16_mode ? c->i16_ring : f->input_buffer,
                 c->i16_mode ? c->i16_ring_size : f->input_buffer_size, c->st) != 0)
    rc = kgf_fail("execute_filter_input: H2D of the window");
  if (rc == 0 && kgpu_forward(c->km, c->d_win[slot], fmt, scale, c->i16_derand, 1, spec, NULL, c->st) != 0)
    rc = kgf_fail("execute_filter_input: kgpu_forward");
  if (rc == 0 && f->notches)
    kgpu_apply_notches(c->km, spec, 1, c->st);
  /* every slave, batched, with the shift it used last (radio.c:1491: shifts move only on retune) */
  if (rc == 0 && c->nslots > 0) {
    kgpu_bank_commit(c->bank, c->st);
    long const stride = kgpu_bank_out_stride(c->bank);
    if (stride > 0 && grow_out(c, stride) == 0) {
      if (kgpu_bank_run(c->bank, spec, 1, c->d_out[slot], c->st) == 0) {
        cudaMemcpyAsync(c->h_out[slot], c->d_out[slot], sizeof(float complex) * (size_t)stride, cudaMemcpyDeviceToHost,
                        c->st);
        for (int i = 0; i < c->nslots; i++) {
          struct filter_out *o = c->slots[i];
          c->snap_ok[slot][i] = false;
          if (!o || o->out_type != COMPLEX || !o->response)
            continue;
          c->snap_shift[slot][i] = c->cur_shift[i];
          c->snap_isb[slot][i] = c->cur_isb[i];
          c->snap_off[slot][i] = kgpu_bank_out_offset(c->bank, i);
          c->snap_ver[slot][i] = c->ver[i];
          c->snap_ok[slot][i] = true;
        }
      }
    }
  }
  if (rc == 0 && c->spectrum_d2h)
    cudaMemcpyAsync(f->fdomain[slot], spec, sizeof(float complex) * (size_t)f->bins, cudaMemcpyDeviceToHost, c->st);
  cudaEventRecord(c->done[slot], c->st);
  c->timed[slot] = true;
  pthread_mutex_unlock(&c->mu);

  pthread_mutex_lock(&f->filter_mutex);
  f->owner = pthread_self();
  f->next_jobnum++;
  f->samples_by_job[slot] = f->sample_index;
  f->completed_jobs[slot] = jobnum; /* "complete" == issued; consumers wait on the slot's event */
  pthread_cond_broadcast(&f->filter_cond);
  pthread_mutex_unlock(&f->filter_mutex);
  f->sample_index += (uint64_t)f->ilen;
  if (f->perform_inline)
    cudaEventSynchronize(c->done[slot]);
  return rc;
}

/* ---------------------------------------------------------------- execute_filter_output ----- */
static int ensure_one(struct master_ctx *c, int olen) {
  if (olen <= c->one_cap)
    return 0;
  cudaStreamSynchronize(c->st_one);
  cudaFree(c->d_one);
  cudaFreeHost(c->h_one);
  c->d_one = NULL;
  c->h_one = NULL;
  c->one_cap = 0;
  if (cudaMalloc((void **)&c->d_one, sizeof(float complex) * (size_t)olen) != cudaSuccess ||
      cudaHostAlloc((void **)&c->h_one, sizeof(float complex) * (size_t)olen, cudaHostAllocPortable) != cudaSuccess)
    return -1;
  c->one_cap = olen;
  return 0;
}

/* filter.c:663-921 */
int execute_filter_output(struct filter_out *const slave, int const shift) {
  if (slave == NULL)
    return -1;
  struct filter_in *const master = slave->master;
  if (master == NULL || master->fwd_plan == NULL) /* transient, filter.c:670-671 */
    return -1;
  struct master_ctx *c = (struct master_ctx *)master->fwd_plan;

  pthread_mutex_lock(&master->filter_mutex);
  if (pthread_equal(master->owner, pthread_self())) {
    slave->next_jobnum = master->next_jobnum - 1; /* same thread wrote the input: take the latest (filter.c:681-683) */
  } else {
    while ((int)(slave->next_jobnum - master->completed_jobs[slave->next_jobnum % ND]) > 0)
      pthread_cond_wait(&master->filter_cond, &master->filter_mutex);
    int const behind = (int)(master->completed_jobs[slave->next_jobnum % ND] - slave->next_jobnum);
    if (behind >= ND) { /* lapped: a block of zeros and a drop (filter.c:690-701) */
      pthread_mutex_unlock(&master->filter_mutex);
      slave->block_drops++;
      slave->next_jobnum++;
      if (slave->output_buffer.c != NULL)
        memset(slave->output_buffer.c, 0, sizeof(float complex) * (size_t)slave->points);
      return 0;
    }
  }
  unsigned const job = slave->next_jobnum;
  int const slot = (int)(job % ND);
  slave->sample_index = master->samples_by_job[slot];
  slave->next_jobnum++;
  pthread_mutex_unlock(&master->filter_mutex);

  if (cudaEventSynchronize(c->done[slot]) != cudaSuccess)
    return kgf_fail("execute_filter_output: waiting for the block");
  if (slave->out_type == SPECTRUM)
    return 0; /* the caller reads master->fdomain[] itself (filter.c:368-371, spectrum.c:318) */
  struct slave_ctx *sc = (struct slave_ctx *)slave->rev_plan;
  if (sc == NULL || slave->output.c == NULL)
    return -1;
  if (slave->response == NULL) /* no filter yet: leave the output alone (filter.c:715-718) */
    return 0;
  if (slave->beam) {
    fprintf(stderr, "execute_filter_output: beam synthesis is not served by the GPU backend\n");
    return -1;
  }
  int rc = 0;
  pthread_mutex_lock(&c->mu);
  int const i = sc->idx;
  if (c->snap_ok[slot][i] && c->snap_shift[slot][i] == shift && c->snap_ver[slot][i] == c->ver[i] &&
      c->snap_isb[slot][i] == slave->isb) {
    memcpy(slave->output.c, c->h_out[slot] + c->snap_off[slot][i], sizeof(float complex) * (size_t)slave->olen);
  } else {
    /* this slave's parameters moved after the block was issued (or it is new): redo it alone
     * from the block's spectrum, and let the next batched launches use the new shift */
    cudaStreamSynchronize(c->st);
    c->cur_shift[i] = shift;
    c->cur_isb[i] = slave->isb;
    kgpu_bank_set_shift(c->bank, i, shift);
    kgpu_bank_set_flags(c->bank, i, slave->isb ? KGPU_CHAN_ISB : 0);
    float complex const *spec = c->d_spec + (size_t)slot * (size_t)c->spec_stride;
    if (ensure_one(c, slave->olen) != 0)
      rc = kgf_fail("execute_filter_output: scratch allocation");
    else if (kgpu_bank_run_one(c->bank, i, spec, c->d_one, c->st_one) != 0)
      rc = kgf_fail("execute_filter_output: kgpu_bank_run_one");
    else if (cudaMemcpyAsync(c->h_one, c->d_one, sizeof(float complex) * (size_t)slave->olen, cudaMemcpyDeviceToHost,
                             c->st_one) != cudaSuccess ||
             cudaStreamSynchronize(c->st_one) != cudaSuccess)
      rc = kgf_fail("execute_filter_output: D2H of the recomputed channel");
    else
      memcpy(slave->output.c, c->h_one, sizeof(float complex) * (size_t)slave->olen);
  }
  pthread_mutex_unlock(&c->mu);
  return rc;
}

int execute_filter_output_batch(struct filter_out *const *slaves, int const *shifts, int n) {
  int rc = 0;
  for (int i = 0; i < n; i++)
    if (execute_filter_output(slaves[i], shifts[i]) != 0)
      rc = -1;
  return rc;
}

/* ---------------------------------------------------------------- set_filter ---------------- */
/* filter.c:968-1045 */
int set_filter(struct filter_out *const slave, double low, double high, double const kaiser_beta) {
  if (slave == NULL || low != low || high != high || kaiser_beta != kaiser_beta || slave->master == NULL)
    return -1;
  struct master_ctx *c = (struct master_ctx *)slave->master->fwd_plan;
  struct slave_ctx *sc = (struct slave_ctx *)slave->rev_plan;
  if (c == NULL || sc == NULL)
    return -1;
  float complex *host = cache_aligned(sizeof(float complex) * (size_t)slave->points);
  if (!host)
    return -1;
  pthread_mutex_lock(&c->mu);
  cudaStreamSynchronize(c->st);
  int rc = kgpu_bank_set_filter(c->bank, sc->idx, low, high, kaiser_beta);
  if (rc == 0)
    rc = kgpu_bank_get_response(c->bank, sc->idx, (float *)host) > 0 ? 0 : -1;
  if (rc == 0)
    c->ver[sc->idx]++;
  pthread_mutex_unlock(&c->mu);
  if (rc != 0) {
    free(host);
    return -1;
  }
  pthread_mutex_lock(&slave->response_mutex); /* hot swap (filter.c:1039-1043) */
  float complex *old = slave->response;
  slave->response = host;
  pthread_mutex_unlock(&slave->response_mutex);
  free(old);
  return 0;
}

int set_filter_weights(struct filter_out *out, double complex i_weight, double complex q_weight) { /* filter.c:922-929 */
  if (out == NULL)
    return -1;
  out->alpha = 0.5 * i_weight - I * q_weight;
  out->beta = 0.5 * i_weight + I * q_weight;
  return 0;
}

/* ---------------------------------------------------------------- delete -------------------- */
int delete_filter_output(struct filter_out *slave) { /* filter.c:943-957 */
  if (slave == NULL)
    return -1;
  struct slave_ctx *sc = (struct slave_ctx *)slave->rev_plan;
  if (sc && slave->master && slave->master->fwd_plan) {
    struct master_ctx *c = (struct master_ctx *)slave->master->fwd_plan;
    pthread_mutex_lock(&c->mu);
    cudaStreamSynchronize(c->st);
    kgpu_bank_enable(c->bank, sc->idx, 0);
    c->slots[sc->idx] = NULL;
    for (int s = 0; s < ND; s++)
      c->snap_ok[s][sc->idx] = false;
    pthread_mutex_unlock(&c->mu);
  }
  free(sc);
  if (slave->init)
    pthread_mutex_destroy(&slave->response_mutex);
  free(slave->output_buffer.c);
  free(slave->output_buffer.r);
  free(slave->response);
  free(slave->fdomain);
  memset(slave, 0, sizeof *slave);
  return 0;
}
int delete_filter_input(struct filter_in *master) { /* filter.c:930-942 */
  if (master == NULL)
    return -1;
  master_teardown(master);
  if (master->init) {
    pthread_mutex_destroy(&master->filter_mutex);
    pthread_cond_destroy(&master->filter_cond);
  }
  memset(master, 0, sizeof *master);
  return 0;
}

/* ---------------------------------------------------------------- write_*filter ------------- */
int write_cfilter(struct filter_in *f, float complex const *buffer, int size) { /* filter.c:1093-1113 */
  if (f == NULL)
    return -1;
  if ((f->wcnt + size) * sizeof *buffer >= f->input_buffer_size)
    return -1;
  if (buffer != NULL)
    memcpy(f->input_write_pointer.c, buffer, (size_t)size * sizeof *buffer);
  f->input_write_pointer.c += size;
  kgf_ring_wrap((void **)&f->input_write_pointer.c, f->input_buffer, f->input_buffer_size);
  f->wcnt += size;
  int fired = 0;
  while (f->wcnt >= f->ilen) {
    f->wcnt -= f->ilen;
    execute_filter_input(f);
    fired = 1;
  }
  return fired;
}
int write_rfilter(struct filter_in *f, float const *buffer, int size) { /* filter.c:1114-1134 */
  if (f == NULL)
    return -1;
  if ((f->wcnt + size) * sizeof *buffer >= f->input_buffer_size)
    return -1;
  if (buffer != NULL)
    memcpy(f->input_write_pointer.r, buffer, (size_t)size * sizeof *buffer);
  f->input_write_pointer.r += size;
  kgf_ring_wrap((void **)&f->input_write_pointer.r, f->input_buffer, f->input_buffer_size);
  f->wcnt += size;
  int fired = 0;
  while (f->wcnt >= f->ilen) {
    f->wcnt -= f->ilen;
    execute_filter_input(f);
    fired = 1;
  }
  return fired;
}
/* EXTENSION: raw ADC words straight to the device; conversion (rx888.c:753-767) happens in fwd_cols */
int write_i16filter(struct filter_in *f, int16_t const *samples, int n, float scale, bool derandomize) {
  if (f == NULL || f->fwd_plan == NULL || samples == NULL || n < 0)
    return -1;
  struct master_ctx *c = (struct master_ctx *)f->fwd_plan;
  if (!c->i16_mode) {
    c->i16_esz = (f->in_type == COMPLEX) ? 2 * sizeof(int16_t) : sizeof(int16_t);
    c->i16_ring_size = page_round((size_t)ND * (size_t)f->points * c->i16_esz);
    c->i16_ring = ring_alloc(c->i16_ring_size);
    if (!c->i16_ring)
      return -1;
    c->i16_rp = c->i16_ring;
    c->i16_wp = c->i16_rp + c->i16_esz * (size_t)(f->impulse_length - 1);
    c->i16_mode = true;
  }
  if (((size_t)f->wcnt + (size_t)n) * c->i16_esz >= c->i16_ring_size)
    return -1;
  c->i16_scale = scale;
  c->i16_derand = derandomize;
  memcpy(c->i16_wp, samples, (size_t)n * c->i16_esz);
  c->i16_wp += (size_t)n * c->i16_esz;
  if (c->i16_wp >= (char *)c->i16_ring + c->i16_ring_size)
    c->i16_wp -= c->i16_ring_size;
  f->wcnt += n;
  int fired = 0;
  while (f->wcnt >= f->ilen) {
    f->wcnt -= f->ilen;
    execute_filter_input(f);
    fired = 1;
  }
  return fired;
}

/* ---------------------------------------------------------------- housekeeping -------------- */
void *run_fft(void *p) { /* filter.c:485: the CPU FFT worker pool has no GPU counterpart */
  (void)p;
  return NULL;
}
void suggest(int size, int dir, int clex) { /* filter.c:1136-1144: wisdom hints are meaningless here */
  (void)size;
  (void)dir;
  (void)clex;
}
long gcd(long a, long b) {
  while (b != 0) {
    long const t = a % b;
    a = b;
    b = t;
  }
  return a;
}
long lcm(long a, long b) {
  if (a <= 0 || b <= 0)
    return 0;
  return a / gcd(a, b) * b;
}
/* "good" now means: plannable by the device transform (factors 2,3,5,7) */
bool goodchoice(long n) {
  if (n <= 0)
    return false;
  static int const primes[4] = {2, 3, 5, 7};
  for (int i = 0; i < 4; i++)
    while (n % primes[i] == 0)
      n /= primes[i];
  return n == 1;
}
int ceil_pow2(uint32_t x) {
  uint32_t p = 1;
  while (p < x && p < 0x80000000u)
    p <<= 1;
  return (int)p;
}

/* spectrum.c's own analysis FFTs (filter.h:112-115): forwarded to the host's FFTW when present */
static void *fftw_handle(void) {
  static void *h;
  static int tried;
  if (!tried) {
    tried = 1;
    h = dlopen("libfftw3f.so.3", RTLD_NOW | RTLD_GLOBAL);
  }
  return h;
}
fftwf_plan plan_complex(int N, float complex *in, float complex *out, int direction) {
  void *h = fftw_handle();
  fftwf_plan (*fn)(int, float complex *, float complex *, int, unsigned) = h ? dlsym(h, "fftwf_plan_dft_1d") : NULL;
  return fn ? fn(N, in, out, direction, 1u << 6 /* FFTW_ESTIMATE */) : NULL;
}
fftwf_plan plan_r2c(int N, float *in, float complex *out) {
  void *h = fftw_handle();
  fftwf_plan (*fn)(int, float *, float complex *, unsigned) = h ? dlsym(h, "fftwf_plan_dft_r2c_1d") : NULL;
  return fn ? fn(N, in, out, 1u << 6) : NULL;
}
fftwf_plan plan_c2r(int N, float complex *in, float *out) {
  void *h = fftw_handle();
  fftwf_plan (*fn)(int, float complex *, float *, unsigned) = h ? dlsym(h, "fftwf_plan_dft_c2r_1d") : NULL;
  return fn ? fn(N, in, out, 1u << 6) : NULL;
}
void destroy_plan(fftwf_plan *plan) {
  if (plan == NULL || *plan == NULL)
    return;
  void *h = fftw_handle();
  void (*fn)(fftwf_plan) = h ? dlsym(h, "fftwf_destroy_plan") : NULL;
  if (fn)
    fn(*plan);
  *plan = NULL;
}
