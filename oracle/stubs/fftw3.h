/* oracle/stubs/fftw3.h -- declaration-only stand-in for FFTW 3's single-precision API.
 * TEST INFRASTRUCTURE.  FFTW3 (external dependency of the reference, src/Makefile:290) is not
 * installed in this image; this header declares exactly the entry points the reference's
 * src/filter.c uses so that file compiles UNMODIFIED, and oracle/fftw_shim.c implements them
 * on top of oracle/fft_cpu.c.  Semantics restated from the FFTW 3.3 manual:
 *   - transforms are unnormalised; FFTW_FORWARD = -1 exponent sign, FFTW_BACKWARD = +1
 *   - r2c returns n/2+1 bins; c2r consumes n/2+1 bins
 *   - fftwf_execute() uses the arrays given at plan time; fftwf_execute_dft*() new arrays
 */
#ifndef KA9Q_ORACLE_FFTW3_SHIM_H
#define KA9Q_ORACLE_FFTW3_SHIM_H 1
#include <complex.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef float complex fftwf_complex;
typedef struct kshim_plan_s *fftwf_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_EXHAUSTIVE (1U << 3)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)
#define FFTW_WISDOM_ONLY (1U << 21)

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
fftwf_plan fftwf_plan_dft_r2c_1d(int n, float *in, fftwf_complex *out, unsigned flags);
fftwf_plan fftwf_plan_dft_c2r_1d(int n, fftwf_complex *in, float *out, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_execute_dft(const fftwf_plan p, fftwf_complex *in, fftwf_complex *out);
void fftwf_execute_dft_r2c(const fftwf_plan p, float *in, fftwf_complex *out);
void fftwf_destroy_plan(fftwf_plan p);
int fftwf_init_threads(void);
void fftwf_plan_with_nthreads(int nthreads);
int fftwf_import_system_wisdom(void);
int fftwf_import_wisdom_from_filename(const char *filename);
extern const char fftwf_version[];
#ifdef __cplusplus
}
#endif
#endif
