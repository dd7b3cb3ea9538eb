/*
 * CPU oracle (C) for the numtel/ntru-circom hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product never does.
 *
 * It is a structural restatement of /root/reference/index.js: the same
 * algorithm with the same operation counts (the reference itself is
 * JavaScript and cannot run in this image -- no node), so that it can serve
 * both as a second checker and as the "port" CPU baseline:
 *
 *   multiply  -> index.js:319-355 (+ fft index.js:277-316): float64 radix-2
 *                Cooley-Tukey with the running twiddle w *= wlen, Math.round,
 *                ((x % p) + p) % p, trim
 *   divide    -> index.js:358-401 (+ modInverse 224-232, degree 210-215):
 *                schoolbook long division, degree() rescan and the O(p)
 *                linear-search modInverse on every quotient term
 *   add       -> index.js:235-244
 *   encrypt   -> index.js:87-110   (r injected instead of WebCrypto)
 *   decrypt   -> index.js:111-140  (incl. the (x > q/2 ? x+1 : x) % p lift)
 *
 * JavaScript numbers are doubles; every integer on this path stays below
 * 2^53 so int64 arithmetic reproduces `%`, `*`, `-` exactly (C's % truncates
 * like JS's).  Parity pinning: see oracle/ntru_oracle.py's header; this file
 * is checked against that module and the upstream KATs in tests/.
 *
 * Build: make -C oracle   (gcc -O2 -pthread -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef int64_t i64;

/* index.js:210-215 */
static int degree(const i64 *poly, int len) {
  for (int i = len - 1; i >= 0; i--)
    if (poly[i] != 0) return i;
  return -1;
}

/* index.js:218-221 -- returns the trimmed length (>= 1), writing [0] for the zero polynomial */
static int trim(i64 *poly, int len) {
  int d = degree(poly, len);
  if (d >= 0) return d + 1;
  poly[0] = 0;
  return 1;
}

/* index.js:224-232 -- returns 0 for "null" (no inverse) */
static i64 mod_inverse(i64 a, i64 p) {
  a = ((a % p) + p) % p;
  for (i64 x = 1; x < p; x++)
    if ((a * x) % p == 1) return x;
  return 0;
}

/* index.js:277-316 */
static void fft(double *re, double *im, int n, int invert) {
  for (int i = 1, j = 0; i < n; i++) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j -= bit;
    j += bit;
    if (i < j) {
      double t = re[i]; re[i] = re[j]; re[j] = t;
      t = im[i]; im[i] = im[j]; im[j] = t;
    }
  }
  for (int len = 2; len <= n; len <<= 1) {
    double angle = (2 * M_PI / len) * (invert ? -1 : 1);
    double wlr = cos(angle), wli = sin(angle);
    for (int i = 0; i < n; i += len) {
      double wr = 1, wi = 0;
      for (int j = 0; j < len / 2; j++) {
        double ur = re[i + j], ui = im[i + j];
        double xr = re[i + j + len / 2], xi = im[i + j + len / 2];
        double vr = xr * wr - xi * wi, vi = xr * wi + xi * wr;
        re[i + j] = ur + vr; im[i + j] = ui + vi;
        re[i + j + len / 2] = ur - vr; im[i + j + len / 2] = ui - vi;
        double nwr = wr * wlr - wi * wli, nwi = wr * wli + wi * wlr;
        wr = nwr; wi = nwi;
      }
    }
  }
  if (invert)
    for (int i = 0; i < n; i++) { re[i] /= n; im[i] /= n; }
}

/* index.js:319-355.  out must hold la+lb-1 (>=1) entries; returns trimmed length.
 * *margin (optional) receives max |Re - round(Re)|. */
int oracle_multiply(const i64 *a, int la, const i64 *b, int lb, i64 p, i64 *out, double *margin) {
  if (la == 0 || lb == 0) { out[0] = 0; return 1; }
  int n = 1;
  while (n < la + lb - 1) n <<= 1;
  double *buf = (double *)malloc(sizeof(double) * 4 * (size_t)n);
  double *ar = buf, *ai = buf + n, *br = buf + 2 * n, *bi = buf + 3 * n;
  for (int i = 0; i < n; i++) {
    ar[i] = i < la ? (double)a[i] : 0; ai[i] = 0;
    br[i] = i < lb ? (double)b[i] : 0; bi[i] = 0;
  }
  fft(ar, ai, n, 0);
  fft(br, bi, n, 0);
  for (int i = 0; i < n; i++) {
    double r = ar[i] * br[i] - ai[i] * bi[i];
    double m = ar[i] * bi[i] + ai[i] * br[i];
    ar[i] = r; ai[i] = m;
  }
  fft(ar, ai, n, 1);
  int rl = la + lb - 1;
  double worst = 0;
  for (int i = 0; i < rl; i++) {
    double rd = floor(ar[i] + 0.5);           /* Math.round */
    double d = fabs(ar[i] - rd);
    if (d > worst) worst = d;
    i64 v = (i64)rd;
    out[i] = ((v % p) + p) % p;
  }
  if (margin && worst > *margin) *margin = worst;
  free(buf);
  return trim(out, rl);
}

/* index.js:235-244; out holds max(la,lb) entries (>=1); returns trimmed length */
int oracle_add(const i64 *a, int la, const i64 *b, int lb, i64 p, i64 *out) {
  int n = la > lb ? la : lb;
  for (int i = 0; i < n; i++) {
    i64 ca = i < la ? a[i] : 0, cb = i < lb ? b[i] : 0;
    out[i] = ((ca + cb) % p + p) % p;
  }
  if (n == 0) { out[0] = 0; return 1; }
  return trim(out, n);
}

/* index.js:358-401.  quotient holds max(1, la) entries, remainder holds la+lb entries.
 * Returns 0 ok, -1 "Cannot divide by zero polynomial.", -2 "No inverse exists for division." */
int oracle_divide(const i64 *a, int la, const i64 *b, int lb, i64 p,
                  i64 *quotient, int *lq, i64 *remainder, int *lr) {
  int deg_divisor = degree(b, lb);
  if (deg_divisor == -1) return -1;
  int cap = la + lb;
  int dl = la;                                   /* dividend = a.slice() */
  memcpy(remainder, a, sizeof(i64) * (size_t)la);
  memset(remainder + la, 0, sizeof(i64) * (size_t)(cap - la));
  int ql = degree(a, la) - deg_divisor + 1;
  if (ql < 0) ql = 0;
  for (int i = 0; i < (ql > 0 ? ql : 1); i++) quotient[i] = 0;
  while (degree(remainder, dl) >= deg_divisor) {
    int deg_dividend = degree(remainder, dl);
    i64 lead_dividend = remainder[deg_dividend];
    i64 lead_divisor = b[deg_divisor];
    i64 inv = mod_inverse(lead_divisor, p);
    if (inv == 0) return -2;
    i64 coeff = (lead_dividend * inv) % p;
    int deg_diff = deg_dividend - deg_divisor;
    quotient[deg_diff] = coeff;
    for (int i = 0; i <= deg_divisor; i++) {
      int idx = i + deg_diff;
      if (idx >= dl) dl = idx + 1;               /* (dividend[index] || 0) growth */
      i64 v = (remainder[idx] - coeff * b[i]) % p;
      if (v < 0) v += p;
      remainder[idx] = v;
    }
  }
  *lq = trim(quotient, ql);
  *lr = dl > 0 ? trim(remainder, dl) : trim(remainder, 0);
  return 0;
}

/* scratch for one encrypt/decrypt */
typedef struct {
  i64 *prod, *sum, *quo, *rem, *I, *tmp;
} scratch_t;

static void scratch_init(scratch_t *s, int N) {
  size_t n = (size_t)(4 * N + 8);
  s->prod = (i64 *)malloc(sizeof(i64) * n);
  s->sum = (i64 *)malloc(sizeof(i64) * n);
  s->quo = (i64 *)malloc(sizeof(i64) * n);
  s->rem = (i64 *)malloc(sizeof(i64) * n);
  s->tmp = (i64 *)malloc(sizeof(i64) * n);
  s->I = (i64 *)calloc((size_t)N + 1, sizeof(i64));
  s->I[0] = 1;                                   /* index.js:25-27 */
  s->I[N] = -1;
}

static void scratch_free(scratch_t *s) {
  free(s->prod); free(s->sum); free(s->quo); free(s->rem); free(s->tmp); free(s->I);
}

static void expand_to(const i64 *src, int len, i64 *dst, int outlen) {
  for (int i = 0; i < outlen; i++) dst[i] = i < len ? src[i] : 0;
}

/* index.js:87-110 with r injected.  h: trimmed public key (hl entries), r: N entries in {0,1,2},
 * m: ml <= N entries.  Outputs are the expandArray'd fields: value N, quotientE N+1, remainderE N+1. */
static int encrypt_one(int N, i64 q, const i64 *h, int hl, const i64 *r, const i64 *m, int ml,
                       i64 *value, i64 *quotientE, i64 *remainderE, scratch_t *s, double *margin) {
  int pl = oracle_multiply(r, N, h, hl, q, s->prod, margin);           /* :90 */
  int sl = oracle_add(m, ml, s->prod, pl, q, s->sum);                    /* :91 */
  int lq, lr;
  int rc = oracle_divide(s->sum, sl, s->I, N + 1, q, s->quo, &lq, s->rem, &lr);   /* :92 */
  if (rc) return rc;
  for (int i = 0; i < lq; i++) s->quo[i] %= q;                           /* :101 */
  if (value) expand_to(s->rem, lr, value, N);
  if (quotientE) expand_to(s->quo, lq, quotientE, N + 1);
  if (remainderE) expand_to(s->rem, lr, remainderE, N + 1);
  return 0;
}

/* index.js:111-140.  f: N entries in {-1,0,1}; fp: trimmed (fpl entries); e: el <= N entries. */
static int decrypt_one(int N, i64 q, i64 p, const i64 *f, const i64 *fp, int fpl, const i64 *e, int el,
                       i64 *value, i64 *quotient1, i64 *remainder1, i64 *quotient2, i64 *remainder2,
                       scratch_t *s, double *margin) {
  for (int i = 0; i < N; i++) s->tmp[i] = f[i] == -1 ? q - 1 : f[i];     /* :112 */
  int al = oracle_multiply(s->tmp, N, e, el, q, s->prod, margin);        /* :113 */
  int lq, lr;
  int rc = oracle_divide(s->prod, al, s->I, N + 1, q, s->quo, &lq, s->rem, &lr);  /* :114 */
  if (rc) return rc;
  if (quotient1) expand_to(s->quo, lq, quotient1, N + 1);
  if (remainder1) expand_to(s->rem, lr, remainder1, N + 1);
  for (int i = 0; i < lr; i++) {                                         /* :117 */
    i64 x = s->rem[i];
    s->tmp[i] = (2 * x > q) ? (x + 1) % p : x % p;                       /* x > q/2 */
  }
  int cl = oracle_multiply(fp, fpl, s->tmp, lr, p, s->prod, margin);     /* :118 */
  s->I[N] = -1;
  rc = oracle_divide(s->prod, cl, s->I, N + 1, p, s->quo, &lq, s->rem, &lr);      /* :119 */
  if (rc) return rc;
  if (value) expand_to(s->rem, lr, value, N);
  if (quotient2) expand_to(s->quo, lq, quotient2, N + 1);
  if (remainder2) expand_to(s->rem, lr, remainder2, N + 1);
  return 0;
}

static int trimmed_len(const i64 *a, int n) {
  int d = degree(a, n);
  return d >= 0 ? d + 1 : 1;
}

/*
 * Batched drivers (pthreads; this image's gcc has no libgomp).  Arrays are row-major with
 * the stated pitch; key arrays are shared when key_stride == 0, else per-row with pitch N.
 * m and e rows are passed un-trimmed (N entries) and trimmed here the way a caller of the
 * JS API would hold them (trailing zeros are irrelevant to every function on the path).
 * nthreads <= 0: all online cores.  Returns 0, or the first error code.
 */
typedef struct {
  int kind;                 /* 0 encrypt, 1 decrypt */
  int N; i64 q, p;
  const i64 *k0, *k1; int key_stride;
  long b0, b1;
  const i64 *in0, *in1;
  i64 *o0, *o1, *o2, *o3, *o4;
  int err; double margin;
} job_t;

static void *job_run(void *arg) {
  job_t *j = (job_t *)arg;
  scratch_t s;
  scratch_init(&s, j->N);
  int N = j->N;
  for (long b = j->b0; b < j->b1; b++) {
    int rc;
    if (j->kind == 0) {
      const i64 *hb = j->k0 + (size_t)j->key_stride * b;
      const i64 *mb = j->in1 + (size_t)N * b;
      rc = encrypt_one(N, j->q, hb, trimmed_len(hb, N), j->in0 + (size_t)N * b, mb, trimmed_len(mb, N),
                       j->o0 ? j->o0 + (size_t)N * b : 0,
                       j->o1 ? j->o1 + (size_t)(N + 1) * b : 0,
                       j->o2 ? j->o2 + (size_t)(N + 1) * b : 0, &s, &j->margin);
    } else {
      const i64 *fb = j->k0 + (size_t)j->key_stride * b, *fpb = j->k1 + (size_t)j->key_stride * b;
      const i64 *eb = j->in0 + (size_t)N * b;
      rc = decrypt_one(N, j->q, j->p, fb, fpb, trimmed_len(fpb, N), eb, trimmed_len(eb, N),
                       j->o0 ? j->o0 + (size_t)N * b : 0,
                       j->o1 ? j->o1 + (size_t)(N + 1) * b : 0,
                       j->o2 ? j->o2 + (size_t)(N + 1) * b : 0,
                       j->o3 ? j->o3 + (size_t)(N + 1) * b : 0,
                       j->o4 ? j->o4 + (size_t)(N + 1) * b : 0, &s, &j->margin);
    }
    if (rc && !j->err) j->err = rc;
  }
  scratch_free(&s);
  return 0;
}

int oracle_num_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

static int run_jobs(job_t proto, long B, int nthreads, double *margin_out) {
  if (nthreads <= 0) nthreads = oracle_num_threads();
  if (nthreads > B) nthreads = B > 0 ? (int)B : 1;
  job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)nthreads);
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = proto;
    jobs[t].b0 = B * t / nthreads;
    jobs[t].b1 = B * (t + 1) / nthreads;
    jobs[t].err = 0; jobs[t].margin = 0;
    if (t > 0) pthread_create(&th[t], 0, job_run, &jobs[t]);
  }
  job_run(&jobs[0]);
  int err = jobs[0].err;
  double margin = jobs[0].margin;
  for (int t = 1; t < nthreads; t++) {
    pthread_join(th[t], 0);
    if (jobs[t].err && !err) err = jobs[t].err;
    if (jobs[t].margin > margin) margin = jobs[t].margin;
  }
  free(jobs); free(th);
  if (margin_out) *margin_out = margin;
  return err;
}

int oracle_encrypt_batch(int N, i64 q, const i64 *h, int key_stride, long B, const i64 *r, const i64 *m,
                         i64 *value, i64 *quotientE, i64 *remainderE, int nthreads, double *margin_out) {
  job_t j;
  memset(&j, 0, sizeof j);
  j.kind = 0; j.N = N; j.q = q; j.k0 = h; j.key_stride = key_stride;
  j.in0 = r; j.in1 = m; j.o0 = value; j.o1 = quotientE; j.o2 = remainderE;
  return run_jobs(j, B, nthreads, margin_out);
}

int oracle_decrypt_batch(int N, i64 q, i64 p, const i64 *f, const i64 *fp, int key_stride, long B,
                         const i64 *e, i64 *value, i64 *quotient1, i64 *remainder1,
                         i64 *quotient2, i64 *remainder2, int nthreads, double *margin_out) {
  job_t j;
  memset(&j, 0, sizeof j);
  j.kind = 1; j.N = N; j.q = q; j.p = p; j.k0 = f; j.k1 = fp; j.key_stride = key_stride;
  j.in0 = e; j.o0 = value; j.o1 = quotient1; j.o2 = remainder1; j.o3 = quotient2; j.o4 = remainder2;
  return run_jobs(j, B, nthreads, margin_out);
}

/* Left fold of addPolynomials over B rows (test/reference.test.js:58); out: N entries (expanded). */
int oracle_sum(int N, i64 q, long B, const i64 *e, i64 *out) {
  i64 *acc = (i64 *)calloc((size_t)N + 1, sizeof(i64));
  i64 *tmp = (i64 *)calloc((size_t)N + 1, sizeof(i64));
  int al = 1;
  for (long b = 0; b < B; b++) {
    const i64 *eb = e + (size_t)N * b;
    int l = oracle_add(acc, al, eb, trimmed_len(eb, N), q, tmp);
    memcpy(acc, tmp, sizeof(i64) * (size_t)l);
    al = l;
  }
  expand_to(acc, al, out, N);
  free(acc); free(tmp);
  return 0;
}

