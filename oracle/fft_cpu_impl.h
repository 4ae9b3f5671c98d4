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
static void F(pass_)(int r, int n, int s, CPLX const *restrict x, CPLX *restrict y,
                     CPLX const *restrict tw /* W_leaf^t, t < leaf, sign applied */, int sgn) {
  int const m = n / r;
  REAL const sg = (REAL)sgn; /* -1 forward, +1 backward: sign of the imaginary rotations */
  switch (r) {
  case 2:
    for (int p = 0; p < m; p++) {
      CPLX const w1 = tw[(long)p * s];
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
      CPLX const w1 = tw[(long)p * s], w2 = tw[(long)2 * p * s];
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
      CPLX const w1 = tw[(long)p * s], w2 = tw[(long)2 * p * s], w3 = tw[(long)3 * p * s];
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
      CPLX const w1 = tw[(long)p * s], w2 = tw[(long)2 * p * s], w3 = tw[(long)3 * p * s],
                 w4 = tw[(long)4 * p * s];
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
    long const leaf = (long)n * s;
    long const rstep = leaf / r; /* W_r^1 == tw[rstep] */
    for (int p = 0; p < m; p++)
      for (int q = 0; q < s; q++)
        for (int k = 0; k < r; k++) {
          CPLX acc = 0;
          for (int j = 0; j < r; j++)
            acc += x[q + s * (p + m * j)] * tw[(((long)j * k) % r) * rstep];
          y[q + s * (r * p + k)] = acc * tw[((long)p * k * s) % leaf];
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
    F(pass_)(r, cur, s, src, dst, tw, sign);
    src = dst;
    cur /= r;
    s *= r;
  }
}

static void F(exec_)(struct kfft_plan const *p, CPLX const *in, CPLX *out, int sign);

/* Four-step: n = n1*n2, input viewed [n1][n2]; X[k1 + n1*k2]. */
static void F(big_)(struct kfft_plan const *p, CPLX const *in, CPLX *out, int sign) {
  int const n1 = p->n1, n2 = p->n2;
  long const n = (long)n1 * n2;
  enum { TB = 16 };
  CPLX *tmp = (CPLX *)kfft_scratch(0 + 3 * p->depth, sizeof(CPLX) * (size_t)n);
  CPLX *col = (CPLX *)kfft_scratch(1 + 3 * p->depth, sizeof(CPLX) * (size_t)TB * n1 * 2);
  double complex const *hi = (sign < 0) ? p->big_hi : p->big_hi_b;
  double complex const *lo = (sign < 0) ? p->big_lo : p->big_lo_b;
  int const S = p->big_S;
  /* step 1: column transforms of length n1 (stride n2), twiddle, store [k1][i2] */
  for (int c0 = 0; c0 < n2; c0 += TB) {
    int const tb = (n2 - c0 < TB) ? n2 - c0 : TB;
    for (int i1 = 0; i1 < n1; i1++)
      for (int j = 0; j < tb; j++)
        col[(size_t)j * n1 + i1] = in[(size_t)i1 * n2 + c0 + j];
    for (int j = 0; j < tb; j++) {
      CPLX *cj = col + (size_t)j * n1;
      CPLX *cw = col + (size_t)(TB + j) * n1;
      F(exec_)(p->sub1, cj, cw, sign); /* result in cw */
      long const i2 = c0 + j;
      for (int k1 = 0; k1 < n1; k1++) {
        long const t = i2 * k1; /* < n */
        double complex const w = hi[t / S] * lo[t % S];
        cw[k1] = (CPLX)((double complex)cw[k1] * w);
      }
    }
    for (int k1 = 0; k1 < n1; k1++)
      for (int j = 0; j < tb; j++)
        tmp[(size_t)k1 * n2 + c0 + j] = col[(size_t)(TB + j) * n1 + k1];
  }
  /* step 2: row transforms of length n2 in place */
  for (int k1 = 0; k1 < n1; k1++)
    F(exec_)(p->sub2, tmp + (size_t)k1 * n2, tmp + (size_t)k1 * n2, sign);
  /* step 3: out[k1 + n1*k2] = tmp[k1][k2]  (blocked transpose) */
  enum { BT = 32 };
  for (int ka = 0; ka < n1; ka += BT)
    for (int kb = 0; kb < n2; kb += BT) {
      int const ea = ka + BT < n1 ? ka + BT : n1, eb = kb + BT < n2 ? kb + BT : n2;
      for (int k2 = kb; k2 < eb; k2++)
        for (int k1 = ka; k1 < ea; k1++)
          out[(size_t)k1 + (size_t)n1 * k2] = tmp[(size_t)k1 * n2 + k2];
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
