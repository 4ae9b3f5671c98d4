/* oracle/fft_cpu_impl.h -- precision-generic body of the oracle FFT (included twice by
 * fft_cpu.c with REAL=float / REAL=double).  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Leaf transforms (n <= KFFT_LEAF_MAX) are Stockham autosort passes, radix 4/2/3/5 (+ generic
 * odd primes).  Larger n use the four-step split n = n1*n2 so every sub-transform stays in cache.
 */
#ifndef REAL
#error "include from fft_cpu.c only"
#endif

#define CPLX REAL complex
#define FN_(a, b) a##b
#define FN(a, b) FN_(a, b)
#define F(name) FN(name, SUFFIX)

/* ---- one Stockham pass: x (length n*s viewed as [r][m][s]) -> y ([m][r][s]) ------------- */
/* n = current sub-length, s = stride (n*s = leaf length), r = radix, m = n/r.
 * y[q + s*(r*p + k)] = W_n^{p*k} * sum_j x[q + s*(p + m*j)] * W_r^{j*k}                      */
static void F(pass_)(int r, int n, int s, int ts, CPLX const *restrict x, CPLX *restrict y,
                     CPLX const *restrict tw /* W_leaf^t, t < leaf, sign applied */, int sgn) {
  /* s = element stride of this pass (includes the batch width when several interleaved transforms
   * are processed at once), ts = twiddle-table stride = product of the radices already done */
  int const m = n / r;
  REAL const sg = (REAL)sgn; /* -1 forward, +1 backward: sign of the imaginary rotations */
  switch (r) {
  case 2:
    for (int p = 0; p < m; p++) {
      CPLX const w1 = tw[(long)p * ts];
      for (int q = 0; q < s; q++) {
        CPLX const a = x[q + s * p], b = x[q + s * (p + m)];
        y[q + s * (2 * p)] = a + b;
        y[q + s * (2 * p + 1)] = (a - b) * w1;
      }
    }
    break;
  case 3: {
    REAL const c = (REAL)-0.5, sn = sg * (REAL)0.86602540378443864676;
    for (int p = 0; p < m; p++) {
      CPLX const w1 = tw[(long)p * ts], w2 = tw[(long)2 * p * ts];
      for (int q = 0; q < s; q++) {
        CPLX const a = x[q + s * p], b = x[q + s * (p + m)], d = x[q + s * (p + 2 * m)];
        CPLX const t1 = b + d, t2 = a + c * t1, t3 = sn * (b - d);
        CPLX const jt3 = CMPLX(-cimag(t3), creal(t3)); /* i*t3 */
        y[q + s * (3 * p)] = a + t1;
        y[q + s * (3 * p + 1)] = (t2 + jt3) * w1;
        y[q + s * (3 * p + 2)] = (t2 - jt3) * w2;
      }
    }
  } break;
  case 4:
    for (int p = 0; p < m; p++) {
      CPLX const w1 = tw[(long)p * ts], w2 = tw[(long)2 * p * ts], w3 = tw[(long)3 * p * ts];
      for (int q = 0; q < s; q++) {
        CPLX const a = x[q + s * p], b = x[q + s * (p + m)], c = x[q + s * (p + 2 * m)],
                   d = x[q + s * (p + 3 * m)];
        CPLX const apc = a + c, amc = a - c, bpd = b + d, bmd = b - d;
        /* sign*i*(b-d) */
        CPLX const jbmd = CMPLX(-sg * cimag(bmd), sg * creal(bmd));
        y[q + s * (4 * p)] = apc + bpd;
        y[q + s * (4 * p + 1)] = (amc + jbmd) * w1;
        y[q + s * (4 * p + 2)] = (apc - bpd) * w2;
        y[q + s * (4 * p + 3)] = (amc - jbmd) * w3;
      }
    }
    break;
  case 5: {
    REAL const c1 = (REAL)0.30901699437494742410, c2 = (REAL)-0.80901699437494742410;
    REAL const s1 = sg * (REAL)0.95105651629515357212, s2 = sg * (REAL)0.58778525229247312917;
    for (int p = 0; p < m; p++) {
      CPLX const w1 = tw[(long)p * ts], w2 = tw[(long)2 * p * ts], w3 = tw[(long)3 * p * ts],
                 w4 = tw[(long)4 * p * ts];
      for (int q = 0; q < s; q++) {
        CPLX const a = x[q + s * p], b = x[q + s * (p + m)], c = x[q + s * (p + 2 * m)],
                   d = x[q + s * (p + 3 * m)], e = x[q + s * (p + 4 * m)];
        CPLX const t1 = b + e, t2 = c + d, t3 = b - e, t4 = c - d;
        CPLX const u1 = a + c1 * t1 + c2 * t2, u2 = a + c2 * t1 + c1 * t2;
        CPLX const v1 = s1 * t3 + s2 * t4, v2 = s2 * t3 - s1 * t4;
        CPLX const jv1 = CMPLX(-cimag(v1), creal(v1)), jv2 = CMPLX(-cimag(v2), creal(v2));
        y[q + s * (5 * p)] = a + t1 + t2;
        y[q + s * (5 * p + 1)] = (u1 + jv1) * w1;
        y[q + s * (5 * p + 2)] = (u2 + jv2) * w2;
        y[q + s * (5 * p + 3)] = (u2 - jv2) * w3;
        y[q + s * (5 * p + 4)] = (u1 - jv1) * w4;
      }
    }
  } break;
  default: { /* generic radix (odd primes > 5): O(r^2) using the leaf table for W_r */
    long const leaf = (long)n * ts;
    long const rstep = leaf / r; /* W_r^1 == tw[rstep] */
    for (int p = 0; p < m; p++)
      for (int q = 0; q < s; q++)
        for (int k = 0; k < r; k++) {
          CPLX acc = 0;
          for (int j = 0; j < r; j++)
            acc += x[q + s * (p + m * j)] * tw[(((long)j * k) % r) * rstep];
          y[q + s * (r * p + k)] = acc * tw[((long)p * k * ts) % leaf];
        }
  } break;
  }
}

/* Leaf transform: ping-pong between `out` and `work`; the last pass always writes `out`. */
static void F(leaf_)(struct kfft_plan const *p, CPLX const *in, CPLX *out, CPLX *work, int sign) {
  int const n = p->n;
  if (n == 1) {
    out[0] = in[0];
    return;
  }
  CPLX const *tw = (sign < 0) ? (CPLX const *)F(p->twf_) : (CPLX const *)F(p->twb_);
  int const np = p->nrad;
  /* pass i writes `out` when (np-1-i) is even, `work` otherwise */
  CPLX const *src = in;
  if (in == out && (np & 1)) { /* pass 0 would write the array it reads */
    memcpy(work, in, sizeof(CPLX) * (size_t)n);
    src = work;
  }
  int cur = n, s = 1;
  for (int i = 0; i < np; i++) {
    int const r = p->rad[i];
    CPLX *dst = ((np - 1 - i) & 1) ? work : out;
    if (dst == src) { /* only possible for the copied-input case with np odd: src=work,dst=out */
      dst = (dst == work) ? out : work;
    }
    F(pass_)(r, cur, s, s, src, dst, tw, sign);
    src = dst;
    cur /= r;
    s *= r;
  }
}

/* V interleaved leaf transforms at once: data[i][v], v < V contiguous.  in != out; result in out.
 * Same passes with every stride multiplied by V, so the innermost loop is always >= V long. */
static void F(leaf_batched_)(struct kfft_plan const *p, CPLX const *in, CPLX *out, CPLX *work, int sign, int V) {
  int const n = p->n;
  if (n == 1) {
    memcpy(out, in, sizeof(CPLX) * (size_t)V);
    return;
  }
  CPLX const *tw = (sign < 0) ? (CPLX const *)F(p->twf_) : (CPLX const *)F(p->twb_);
  int const np = p->nrad;
  CPLX const *src = in;
  int cur = n, ts = 1;
  for (int i = 0; i < np; i++) {
    int const r = p->rad[i];
    CPLX *dst = ((np - 1 - i) & 1) ? work : out;
    F(pass_)(r, cur, ts * V, ts, src, dst, tw, sign);
    src = dst;
    cur /= r;
    ts *= r;
  }
}

static void F(exec_)(struct kfft_plan const *p, CPLX const *in, CPLX *out, int sign);

/* Four-step: n = n1*n2, input viewed [n1][n2]; X[k1 + n1*k2].  Both steps are column transforms
 * done V columns at a time with the V columns interleaved (SIMD across columns):
 *   step 1: FFT over i1 for V adjacent i2, times W_n^{i2*k1}, stored transposed as tmp[i2][k1]
 *   step 2: FFT over i2 for V adjacent k1 -> out[k2*n1 + k1] directly in natural order          */
static void F(big_)(struct kfft_plan const *p, CPLX const *in, CPLX *out, int sign) {
  int const n1 = p->n1, n2 = p->n2;
  long const n = (long)n1 * n2;
  enum { V = 16 };
  int const nmax = n1 > n2 ? n1 : n2;
  CPLX *tmp = (CPLX *)kfft_scratch(0 + 3 * p->depth, sizeof(CPLX) * (size_t)n);
  CPLX *bufs = (CPLX *)kfft_scratch(1 + 3 * p->depth, sizeof(CPLX) * (size_t)V * nmax * 3);
  CPLX *A = bufs, *B = bufs + (size_t)V * nmax, *W = bufs + (size_t)2 * V * nmax;
  double complex const *hi = (sign < 0) ? p->big_hi : p->big_hi_b;
  double complex const *lo = (sign < 0) ? p->big_lo : p->big_lo_b;
  int const S = p->big_S;
  if (p->sub1->n1 != 0 || p->sub2->n1 != 0) { /* sub-transforms too long for one leaf: plain recursion */
    for (int i2 = 0; i2 < n2; i2++) {
      for (int i1 = 0; i1 < n1; i1++) A[i1] = in[(size_t)i1 * n2 + i2];
      F(exec_)(p->sub1, A, B, sign);
      for (int k1 = 0; k1 < n1; k1++) {
        long const t = (long)i2 * k1;
        tmp[(size_t)i2 * n1 + k1] = (CPLX)((double complex)B[k1] * (hi[t / S] * lo[t % S]));
      }
    }
    for (int k1 = 0; k1 < n1; k1++) {
      for (int i2 = 0; i2 < n2; i2++) A[i2] = tmp[(size_t)i2 * n1 + k1];
      F(exec_)(p->sub2, A, B, sign);
      for (int k2 = 0; k2 < n2; k2++) out[(size_t)k2 * n1 + k1] = B[k2];
    }
    return;
  }
  for (int c0 = 0; c0 < n2; c0 += V) {
    int const v = (n2 - c0 < V) ? n2 - c0 : V;
    for (int i1 = 0; i1 < n1; i1++) memcpy(A + (size_t)i1 * v, in + (size_t)i1 * n2 + c0, sizeof(CPLX) * (size_t)v);
    F(leaf_batched_)(p->sub1, A, B, W, sign, v); /* B[k1][v] */
    for (int j = 0; j < v; j++) {
      long const i2 = c0 + j;
      double complex const om = hi[i2 / S] * lo[i2 % S];
      double complex w = 1.0;
      CPLX *dst = tmp + (size_t)i2 * n1;
      for (int k1 = 0; k1 < n1; k1++) {
        if ((k1 & 31) == 0) { /* exact table value every 32 steps, double recurrence in between */
          long const t = i2 * k1;
          w = hi[t / S] * lo[t % S];
        }
        REAL const wr = (REAL)creal(w), wi = (REAL)cimag(w);
        CPLX const x = B[(size_t)k1 * v + j];
        dst[k1] = CMPLX(creal(x) * wr - cimag(x) * wi, creal(x) * wi + cimag(x) * wr);
        w *= om;
      }
    }
  }
  for (int g0 = 0; g0 < n1; g0 += V) {
    int const v = (n1 - g0 < V) ? n1 - g0 : V;
    for (int i2 = 0; i2 < n2; i2++) memcpy(A + (size_t)i2 * v, tmp + (size_t)i2 * n1 + g0, sizeof(CPLX) * (size_t)v);
    F(leaf_batched_)(p->sub2, A, B, W, sign, v); /* B[k2][v] */
    for (int k2 = 0; k2 < n2; k2++) memcpy(out + (size_t)k2 * n1 + g0, B + (size_t)k2 * v, sizeof(CPLX) * (size_t)v);
  }
}

static void F(exec_)(struct kfft_plan const *p, CPLX const *in, CPLX *out, int sign) {
  if (p->n1 == 0) {
    CPLX *work = (CPLX *)kfft_scratch(2 + 3 * p->depth, sizeof(CPLX) * (size_t)p->n);
    F(leaf_)(p, in, out, work, sign);
  } else {
    F(big_)(p, in, out, sign);
  }
}

#undef CPLX
#undef FN_
#undef FN
#undef F
