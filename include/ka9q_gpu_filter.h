/* include/ka9q_gpu_filter.h -- the reference-facing surface of libka9qgpu.so.
 *
 * This header is layout- and name-compatible with ka9q-radio's src/filter.h (commit 4e0033b4):
 * a program built against the reference header (radiod's radio.c / fm.c / linear.c, the front-end
 * drivers) links against libka9qgpu.so instead of filter.o and runs unmodified; tests/
 * test_filter_abi.py checks every field offset against the reference header where it is present.
 * Use this header only when building NEW code without the ka9q-radio tree.
 *
 *   entry point                   replaces (reference file:line)
 *   create_filter_input           filter.c:186-269   master: ring + forward plan  -> device plans
 *   create_filter_output          filter.c:298-415   slave: response/ifft plan    -> bank slot
 *   execute_filter_input          filter.c:558-651   queue forward FFT            -> H2D + 2 kernels
 *   execute_filter_output         filter.c:663-921   wait, slice*response, IFFT   -> batched kernel
 *   set_filter                    filter.c:968-1045  Kaiser design + FFT          -> host design + device FFT
 *   set_filter_weights            filter.c:922-929
 *   write_rfilter / write_cfilter filter.c:1093-1134 ring advance, fire blocks
 *   delete_filter_input/output    filter.c:930-957
 *   write_i16filter               (EXTENSION, not in the reference) raw int16 ingest: fuses
 *                                 rx888.c:753-767 convert() into the first FFT pass
 *
 * Semantics kept: return 0 / -1 (write_*: 1 if a block fired), ND-deep spectrum ring with
 * lap -> zeros + block_drops++ (filter.c:690-701), owner-thread shortcut (filter.c:681-683),
 * missing response -> 0 with stale output (filter.c:715-718), caller-owned structs zeroed by
 * delete_*.  Not supported on the GPU path (return -1): REAL output slaves (wfm/stereod only),
 * beam synthesis (filter.c:756-775).
 */
#ifndef KA9Q_GPU_FILTER_H
#define KA9Q_GPU_FILTER_H 1
#include <assert.h>
#include <complex.h>
#include <pthread.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#error "C header (uses C99 complex); bind through include/ka9q_gpu.h from C++"
#endif

/* FFTW's opaque plan handle appears in the reference structs; the GPU library reuses the two
 * slots for its own context pointers.  Same typedef FFTW itself uses, so both may be included. */
typedef struct fftwf_plan_s *fftwf_plan;

enum filtertype { NONE, COMPLEX, REAL, SPECTRUM };

struct rc {
  float *r;
  float complex *c;
};
struct notch_state {
  int bin;
  double complex state;
  double alpha;
};

#define ND 4 /* depth of the spectrum ring */

struct filter_in {
  enum filtertype in_type;
  int points;         /* N = L + M - 1 */
  int ilen;           /* L */
  int bins;           /* N (complex) or N/2+1 (real) */
  int impulse_length; /* M */
  int wcnt;
  void *input_buffer; /* mirrored host ring the drivers write through input_write_pointer */
  size_t input_buffer_size;
  struct rc input_write_pointer;
  struct rc input_read_pointer;
  fftwf_plan fwd_plan; /* GPU library: master context */
  pthread_mutex_t filter_mutex;
  pthread_cond_t filter_cond;
  struct notch_state *notches; /* assigned by the caller (radio.c:601) */
  float complex *fdomain[ND];  /* host copies of the block spectra (pinned) */
  unsigned int next_jobnum;
  unsigned int completed_jobs[ND];
  bool perform_inline;
  uint64_t sample_index;
  uint64_t samples_by_job[ND];
  bool init;
  pthread_t owner;
};

struct filter_out {
  struct filter_in *master;
  enum filtertype out_type;
  int points;
  int olen;
  int bins;
  double complex alpha;
  double complex beta;
  float complex *fdomain;
  float complex *response; /* host copy of the device response */
  pthread_mutex_t response_mutex;
  struct rc output_buffer;
  struct rc output;
  fftwf_plan rev_plan; /* GPU library: slave context */
  unsigned next_jobnum;
  unsigned block_drops;
  int rcnt;
  uint64_t sample_index;
  bool beam;
  bool isb;
  bool init;
};

/* globals the reference's filter.c owns and radio.c/main.c touch (filter.h:17-24, filter.c:476-479) */
extern char const *Wisdom_file;
extern int N_worker_threads;
extern int N_internal_threads;
extern int FFTW_planning_level;
extern double FFTW_plan_timelimit;
extern int64_t Min_fft_time, Max_fft_time, Avg_fft_time, Mean_dev;

int create_filter_input(struct filter_in *master, int L, int M, enum filtertype in_type);
int create_filter_output(struct filter_out *slave, struct filter_in *master, int olen, enum filtertype out_type);
int execute_filter_input(struct filter_in *master);
int execute_filter_output(struct filter_out *slave, int shift);
int delete_filter_input(struct filter_in *master);
int delete_filter_output(struct filter_out *slave);
int set_filter(struct filter_out *slave, double low, double high, double kaiser_beta);
int set_filter_weights(struct filter_out *slave, double complex i_weight, double complex q_weight);
int write_cfilter(struct filter_in *master, float complex const *samples, int n);
int write_rfilter(struct filter_in *master, float const *samples, int n);
/* EXTENSION: raw ADC words.  n int16 samples (REAL master) or n I/Q pairs (COMPLEX master); a
 * master fed this way must not also be fed through write_rfilter/write_cfilter. */
int write_i16filter(struct filter_in *master, int16_t const *samples, int n, float scale, bool derandomize);
/* Where a driver may deposit the next raw samples itself (pinned, mirrored ring: a block's worth stays contiguous),
 * e.g. as the libusb transfer buffer of rx888.c:797-826; publish with write_i16filter(master, NULL, n, scale, derand). */
int16_t *filter_i16_write_pointer(struct filter_in *master);
/* EXTENSION: serve many slaves with one call (what 1024 channel threads would each do): one wait per block. */
int execute_filter_output_batch(struct filter_out *const *slaves, int const *shifts, int n);
/* EXTENSION (downconvert()'s per-sample work, radio.c:1476-1501 and :1515-1520, on the device): execute_filter_output
 * plus the fine-tuning oscillator (set_osc / step_osc, osc.c:28-70), the block phase correction for shifts that are
 * not multiples of the overlap factor, and the baseband power.  shift, remainder: compute_tuning's results
 * (radio.c:1175-1199); samprate: the slave's output rate; doppler_rate: Hz/s.  *bb_power = chan->sig.bb_power.
 * output.c then holds what radio.c:1499-1501 would have left there; the caller skips that loop. */
int execute_filter_output_tuned(struct filter_out *slave, int shift, double remainder, double samprate, double doppler_rate,
                                double *bb_power);
int filter_output_untune(struct filter_out *slave); /* back to plain execute_filter_output semantics */
/* EXTENSION (estimate_noise, radio.c:1783-1866, on the device): once enabled (samprate = Frontend.samprate), every block
 * carries one noise-density estimate per slave; filter_noise_estimate() returns the one belonging to the block the
 * slave's last execute_filter_output* delivered (NAN if that block had to be recomputed alone after a retune). */
int filter_input_enable_noise(struct filter_in *master, double samprate);
double filter_noise_estimate(struct filter_out const *slave);

/* housekeeping the reference exports from filter.c */
void *run_fft(void *);
void suggest(int size, int dir, int clex);
long gcd(long a, long b);
long lcm(long a, long b);
bool goodchoice(long n);
int ceil_pow2(uint32_t x);
/* spectrum.c plans its own analysis FFTs through these; served by libfftw3f.so.3 when the host
 * has it (dlopen), NULL otherwise */
fftwf_plan plan_complex(int N, float complex *in, float complex *out, int direction);
fftwf_plan plan_r2c(int N, float *in, float complex *out);
fftwf_plan plan_c2r(int N, float complex *in, float *out);
void destroy_plan(fftwf_plan *plan);

/* ---- header-inline sample interface (filter.h:121-163) -------------------------------------- */
static inline void kgf_ring_wrap(void **p, void *base, size_t size) {
  if ((uint8_t *)*p >= (uint8_t *)base + size)
    *p = (uint8_t *)*p - size;
}
static inline int put_cfilter(struct filter_in *f, float complex s) {
  *f->input_write_pointer.c++ = s;
  kgf_ring_wrap((void **)&f->input_write_pointer.c, f->input_buffer, f->input_buffer_size);
  if (++f->wcnt < f->ilen)
    return 0;
  f->wcnt -= f->ilen;
  execute_filter_input(f);
  return 1;
}
static inline int put_rfilter(struct filter_in *f, float s) {
  *f->input_write_pointer.r++ = s;
  kgf_ring_wrap((void **)&f->input_write_pointer.r, f->input_buffer, f->input_buffer_size);
  if (++f->wcnt < f->ilen)
    return 0;
  f->wcnt -= f->ilen;
  execute_filter_input(f);
  return 1;
}
static inline float complex read_cfilter(struct filter_out *f, int rotate) {
  if (f->rcnt == 0) {
    execute_filter_output(f, rotate);
    f->rcnt = f->olen;
  }
  return f->output.c[f->olen - f->rcnt--];
}
static inline float read_rfilter(struct filter_out *f, int rotate) {
  if (f->rcnt == 0) {
    execute_filter_output(f, rotate);
    f->rcnt = f->olen;
  }
  return f->output.r[f->olen - f->rcnt--];
}
#endif
