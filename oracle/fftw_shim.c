/* oracle/fftw_shim.c -- the dozen FFTW3f entry points src/filter.c calls, implemented on
 * oracle/fft_cpu.c.  TEST INFRASTRUCTURE, NOT PRODUCT.  Linked only into oracle/_ref/ (the
 * reference's own filter.c compiled unmodified) so that library has a DFT to call.
 * Reference call sites: filter.c:104-109,125-130,146-151 (plans), :505,:508,:573,:582,:914,:1030
 * (executes), :172 (destroy), :1051-1080 (version, threads, wisdom).
 */
#include <stdlib.h>
#include "fftw3.h"
#include "fft_cpu.h"

const char fftwf_version[] = "ka9q-oracle-shim (not FFTW; oracle/fft_cpu.c)";

enum kind { K_C2C, K_R2C, K_C2R, K_R2C_ODD };
struct kshim_plan_s {
  enum kind kind;
  int n, sign;
  kfft_plan *fft; /* length n (c2c, odd r2c) or n/2 (even r2c / c2r) */
  void *in, *out;
};

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags) {
  if (flags & FFTW_WISDOM_ONLY)
    return NULL; /* no wisdom exists: filter.c then replans with FFTW_ESTIMATE (filter.c:107-110) */
  struct kshim_plan_s *p = calloc(1, sizeof *p);
  p->kind = K_C2C;
  p->n = n;
  p->sign = sign;
  p->fft = kfft_plan_create(n);
  p->in = in;
  p->out = out;
  return p;
}
fftwf_plan fftwf_plan_dft_r2c_1d(int n, float *in, fftwf_complex *out, unsigned flags) {
  if (flags & FFTW_WISDOM_ONLY)
    return NULL;
  struct kshim_plan_s *p = calloc(1, sizeof *p);
  p->n = n;
  p->sign = FFTW_FORWARD;
  if (n % 2 == 0) {
    p->kind = K_R2C;
    p->fft = kfft_plan_create(n / 2);
  } else {
    p->kind = K_R2C_ODD;
    p->fft = kfft_plan_create(n);
  }
  p->in = in;
  p->out = out;
  return p;
}
fftwf_plan fftwf_plan_dft_c2r_1d(int n, fftwf_complex *in, float *out, unsigned flags) {
  if (flags & FFTW_WISDOM_ONLY)
    return NULL;
  if (n % 2 != 0)
    return NULL; /* odd c2r never occurs on the configured paths */
  struct kshim_plan_s *p = calloc(1, sizeof *p);
  p->kind = K_C2R;
  p->n = n;
  p->sign = FFTW_BACKWARD;
  p->fft = kfft_plan_create(n / 2);
  p->in = in;
  p->out = out;
  return p;
}
static void run(struct kshim_plan_s const *p, void *in, void *out) {
  switch (p->kind) {
  case K_C2C:
    kfft_exec_f(p->fft, (float complex const *)in, (float complex *)out, p->sign);
    break;
  case K_R2C:
    kfft_r2c_f(p->fft, (float const *)in, (float complex *)out);
    break;
  case K_C2R:
    kfft_c2r_f(p->fft, (float complex const *)in, (float *)out);
    break;
  case K_R2C_ODD: {
    int const n = p->n;
    float complex *tmp = malloc(sizeof(float complex) * (size_t)n);
    for (int i = 0; i < n; i++)
      tmp[i] = ((float const *)in)[i];
    kfft_exec_f(p->fft, tmp, tmp, -1);
    for (int i = 0; i <= n / 2; i++)
      ((float complex *)out)[i] = tmp[i];
    free(tmp);
  } break;
  }
}
void fftwf_execute(const fftwf_plan p) { run(p, p->in, p->out); }
void fftwf_execute_dft(const fftwf_plan p, fftwf_complex *in, fftwf_complex *out) { run(p, in, out); }
void fftwf_execute_dft_r2c(const fftwf_plan p, float *in, fftwf_complex *out) { run(p, in, out); }
void fftwf_destroy_plan(fftwf_plan p) {
  if (!p)
    return;
  kfft_plan_destroy(p->fft);
  free(p);
}
int fftwf_init_threads(void) { return 1; }
void fftwf_plan_with_nthreads(int nthreads) { (void)nthreads; }
int fftwf_import_system_wisdom(void) { return 0; }
int fftwf_import_wisdom_from_filename(const char *filename) {
  (void)filename;
  return 0;
}
