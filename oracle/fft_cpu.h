/* oracle/fft_cpu.h -- CPU mixed-radix FFT used ONLY as test/oracle infrastructure.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Nothing under ka9q_radio_b200/ may link or call this.
 *
 * Why it exists: every DFT in the reference hot path is delegated to FFTW3 single
 * precision (reference src/filter.c:106,127,148 plans; :505,508,573,582,914,1030 executes;
 * link line src/Makefile:290 `-lfftw3f_threads -lfftw3f`; Debian 12 ships FFTW 3.3.10 per
 * docs/FFTW3.md:137-140).  FFTW is not vendored under /root/reference and is not installed
 * in the build image, so the oracle restates FFTW's *published contract* instead:
 *
 *   fftwf_plan_dft_1d(n, in, out, sign, ..)      Y[k] = sum_j X[j] exp(sign*2*pi*i*j*k/n), unnormalised
 *   fftwf_plan_dft_r2c_1d(n, in, out, ..)        sign = -1, returns bins 0..n/2 (n/2+1 outputs)
 *   fftwf_plan_dft_c2r_1d(n, in, out, ..)        sign = +1, input bins 0..n/2 (Hermitian half)
 *
 * Arithmetic: data in the requested precision (float or double), twiddles computed in
 * double and rounded once.  The float path is what the reference would see from FFTW (same
 * contract, different but equally valid summation order); the double path is the "truth"
 * used by tests to bound both the float oracle's and the GPU's rounding error.
 */
#ifndef KA9Q_ORACLE_FFT_CPU_H
#define KA9Q_ORACLE_FFT_CPU_H 1
#include <complex.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kfft_plan kfft_plan;

/* Complex length-n transform plan; n >= 1, any factorisation (2,3,4,5 fast; other primes O(r^2)). */
kfft_plan *kfft_plan_create(int n);
void kfft_plan_destroy(kfft_plan *p);
int kfft_plan_size(kfft_plan const *p);

/* sign = -1 forward, +1 backward; unnormalised; out may equal in. Thread-safe on a shared plan. */
void kfft_exec_f(kfft_plan const *p, float complex const *in, float complex *out, int sign);
void kfft_exec_d(kfft_plan const *p, double complex const *in, double complex *out, int sign);

/* Real transforms of length n = 2*kfft_plan_size(half) (n even).  r2c: n reals -> n/2+1 bins.
 * c2r: n/2+1 bins -> n reals (unnormalised, destroys nothing). */
void kfft_r2c_f(kfft_plan const *half, float const *in, float complex *out);
void kfft_r2c_d(kfft_plan const *half, double const *in, double complex *out);
void kfft_c2r_f(kfft_plan const *half, float complex const *in, float *out);

#ifdef __cplusplus
}
#endif
#endif
