/* oracle/chan_oracle.h -- CPU restatement of the ka9q-radio overlap-save channelizer path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.  Parity status: PINNED against the
 * reference's own filter.c compiled unmodified (oracle/_ref, see oracle/Makefile) by
 * tests/test_oracle_vs_reference.py, and against the committed fixtures in tests/golden/.
 * The DFT itself is FFTW3's (external, absent): restated in oracle/fft_cpu.c, see its header.
 */
#ifndef KA9Q_CHAN_ORACLE_H
#define KA9Q_CHAN_ORACLE_H 1
#include <complex.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { KO_COMPLEX = 1, KO_REAL = 2 }; /* same numbering as enum filtertype, filter.h:29-34 */

/* Response design (filter.c:968-1045, window.c:217-254, misc.c:416-427, misc.h:217-221,
 * sincospi.c:24-66).  points = Ns, olen = Ls, master_points = N of the master, master_real != 0
 * adds the sqrt(2) for real input (filter.c:1024).  Writes Ns complex floats. Returns 0 / -1. */
int ko_design_response(int points, int olen, int master_points, int master_real, double low, double high,
                       double kaiser_beta, float complex *response);

/* Forward transform of one N-sample window (filter.c:505,508): REAL -> N/2+1 bins, COMPLEX -> N. */
int ko_forward_real(int n, float const *window, float complex *spectrum);
int ko_forward_complex(int n, float complex const *window, float complex *spectrum);
/* Same in double precision ("truth" for error budgeting; not part of the reference). */
int ko_forward_real_d(int n, double const *window, double complex *spectrum);

/* Notch filter state update on a fresh spectrum (filter.c:464-474); list ends at bin 0. */
struct ko_notch {
  int bin;
  double complex state;
  double alpha;
};
void ko_apply_notches(struct ko_notch *list, float complex *spectrum);

/* Bin slice x response (filter.c:728-893, closed form in SURVEY.md 8a, literal loop for the
 * complex-input out-of-domain cases), ISB (filter.c:895-909), Nyquist zero (filter.c:911).
 * in_type = KO_REAL|KO_COMPLEX master; output always COMPLEX (the configured paths). */
void ko_slice_multiply(int in_type, int m_bins, float complex const *m_fdomain, int s_bins,
                       float complex const *response, int shift, int isb, float complex *s_fdomain);

/* One channel, one block: slice, multiply, unnormalised inverse DFT of size points
 * (filter.c:914), returns all `points` samples in `full`; the user part is the last olen
 * (filter.c:357). */
int ko_channel_block(int in_type, int m_bins, float complex const *m_fdomain, int points,
                     float complex const *response, int shift, int isb, float complex *full);

/* int16 -> float ingest (rx888.c:753-767): optional de-randomise, scale, energy, clip count. */
int ko_convert_i16(float *dst, int16_t const *src, int n, float scale, uint64_t *energy, int randomize);

/* Overlap-save streaming front half (filter.c:186-269 ring with M-1 zero prefix, :558-651 block
 * stepping): given the whole input stream (nblocks*L samples), produce block b's N-sample
 * window.  Window b = samples [b*L-(M-1), b*L+L) with zeros before the stream start. */
void ko_block_window_real(float const *stream, int L, int M, int b, float *window);
void ko_block_window_complex(float complex const *stream, int L, int M, int b, float complex *window);

/* Synthetic source (sig_gen.c:288-296 real, :318-322 complex; osc.c:28-70 rotator with
 * renormalisation every 16384 steps; gauss.c:46-110 xoshiro256** seeded with 1). */
typedef struct ko_siggen ko_siggen;
ko_siggen *ko_siggen_new(double cycles_per_sample);
void ko_siggen_free(ko_siggen *g);
void ko_siggen_real(ko_siggen *g, float *dst, long n, double amplitude, double noise, double scale);
void ko_siggen_complex(ko_siggen *g, float complex *dst, long n, double amplitude, double noise, double scale);
/* multi-tone variant for the wideband configs (same oscillator/noise primitives): sums ntones
 * rotators; int16 output = lrint(32767*x) clamped, as an ADC would deliver. */
void ko_siggen_tones_i16(int16_t *dst, long n, int ntones, double const *cycles_per_sample,
                         double const *amplitude, double noise, uint64_t seed);

/* tuning (radio.c:1175-1199): shift = lrint(f/(fs/N)); returns -1 if |shift| >= N/2 */
int ko_compute_tuning(int N, double samprate, double freq, int *shift, double *remainder);

/* ---- oracle/chan_oracle_ext.c: the remaining slice variants and the per-channel steps after the filter ---- */
/* beam synthesis, COMPLEX master -> COMPLEX slave (filter.c:756-775, weights :922-929) */
void ko_slice_beam(int m_bins, float complex const *X, int s_bins, float complex const *R, int shift,
                   double complex alpha, double complex beta, float complex *S);
int ko_channel_block_beam(int m_bins, float complex const *X, int points, float complex const *R, int shift,
                          double are, double aim, double bre, double bim, float complex *full);
/* REAL output slaves (filter.c:794-809), c2r inverse (filter.c:386,914): S has points/2+1 bins, full has points reals */
void ko_slice_realout(int in_type, int m_bins, float complex const *X, int points, float complex const *R, int shift,
                      float complex *S);
int ko_channel_block_realout(int in_type, int m_bins, float complex const *X, int points, float complex const *R,
                             int shift, float *full);

/* fine-tuning oscillator (osc.h:12-19, osc.c:18-70) and the per-block logic of radio.c:1476-1501, :1515-1520 */
struct ko_osc {
  double freq, rate;
  double complex phasor, phasor_step, phasor_step_step;
  int steps;
};
void ko_osc_set(struct ko_osc *o, double f, double r);
double complex ko_osc_step(struct ko_osc *o);
struct ko_finetune {
  struct ko_osc fine;
  double remainder;
  int bin_shift;
  double complex phase_adjust;
};
void ko_finetune_init(struct ko_finetune *s);
double ko_finetune_block(struct ko_finetune *s, int L, int M, int shift, double remainder, double out_samprate,
                         double doppler_rate, float complex *y, int olen); /* returns bb_power */

/* noise density estimate from the master spectrum (radio.c:1783-1866; quantile :1722-1775) */
double ko_estimate_noise(int in_type, int m_bins, float complex const *X, int s_bins, int shift, double samprate);

/* FM discriminator front half (fm.c:104-131 amplitude statistics, fm.c:205-231 arg(x[n] conj x[n-1]) / pi); restated only */
void ko_fm_front(float complex const *x, int n, double complex *phase_memory, float *baseband, double *avg_amp,
                 double *variance_sum);

/* Airspy R2 / HydraSDR packed 12-bit ingest (airspy-unpack.c:106-130) */
int ko_airspy_unpack(float *dst, uint32_t const *packed, int sampcount, float scale, uint64_t *energy);

#ifdef __cplusplus
}
#endif
#endif
