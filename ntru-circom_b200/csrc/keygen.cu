// Batched key generation: loadPrivateKeyF + generatePublicKeyH (index.js:30-79) for B keys.
//
//   fp = f^-1 mod (p, x^N - 1)        polyInv(f, I, p): extended Euclid over GF(3)              (index.js:510-513)
//   fq = f^-1 mod (q, x^N - 1)        polyInv(f, I, q): extended Euclid over GF(2), then the Newton / Hensel lifting
//                                     inverse <- 2 inverse - f inverse^2 (mod q, x^N - 1)        (index.js:494-509)
//   h  = (p fq) * g mod (q, x^N - 1)                                                             (index.js:72-79)
//
// The inverse of f in Z_m[x]/(x^N - 1) is unique, so any correct algorithm returns the reference's coefficients
// (the reference trims; rows here are fixed length).  The split follows the survey: the two extended-Euclid runs are
// sequential, data dependent and cheap (bit-packed over GF(2): ~N^2/64 word operations per key) and stay on the HOST;
// the lifting and h are polynomial products and run on the GPU through the multiply + divide kernel of
// imma_kernels.cu.  One lifting step, written for that kernel (small operand int8, wide operand uint16):
//     t = f * inv                      (f ternary)                    u = (2 - t) mod q = u0 + 128 u1  (u0 < 128, u1 < 64)
//     inv <- inv * u = inv * u0 + 128 (inv * u1)     (mod q, x^N - 1)
// which is the same map inv -> inv (2 - f inv) = 2 inv - f inv^2; precision doubles per step, so ceil(log2 log2 q)
// steps reach q (the reference runs log2 q - 1 steps of the same map: same fixed point).
// Keys whose f is not invertible modulo 2 or modulo p are flagged in valid[] and their outputs zeroed (the reference's
// own checks are vacuous there, SURVEY appendix A: it would hand back garbage; callers redraw f like
// generatePrivateKeyF does).

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <thread>
#include <vector>

#include "ntru_internal.cuh"

namespace ntru {

namespace {

// ---- host: inverse modulo (2, x^N - 1), bit-packed extended Euclid ------------------------------------------------
constexpr int kWords = (2 * kMaxN + 64) / 64 + 1;

struct Bits {
  uint64_t w[kWords];
  void clear() { memset(w, 0, sizeof w); }
  bool zero() const {
    for (int i = 0; i < kWords; ++i)
      if (w[i]) return false;
    return true;
  }
  int degree() const {
    for (int i = kWords - 1; i >= 0; --i)
      if (w[i]) return 64 * i + 63 - __builtin_clzll(w[i]);
    return -1;
  }
  void set(int b) { w[b >> 6] |= 1ull << (b & 63); }
  bool get(int b) const { return (w[b >> 6] >> (b & 63)) & 1; }
  void xor_shifted(const Bits &o, int sh) {      // this ^= o << sh
    const int ws = sh >> 6, bs = sh & 63;
    for (int i = kWords - 1; i >= ws; --i) {
      uint64_t v = o.w[i - ws] << bs;
      if (bs && i - ws - 1 >= 0) v |= o.w[i - ws - 1] >> (64 - bs);
      w[i] ^= v;
    }
  }
};

// inv[i] in {0,1}, i < N; false if gcd(f mod 2, x^N + 1) != 1
bool inverse_mod2(const int8_t *f, int N, uint8_t *inv) {
  Bits u, v, g1, g2;
  u.clear(); v.clear(); g1.clear(); g2.clear();
  for (int i = 0; i < N; ++i)
    if (f[i] & 1) u.set(i);
  v.set(0); v.set(N);
  g1.set(0);
  // invariant: g1 f = u, g2 f = v (mod x^N + 1)
  Bits *pu = &u, *pv = &v, *p1 = &g1, *p2 = &g2;
  while (!pu->zero()) {
    int du = pu->degree(), dv = pv->degree();
    if (du < dv) {
      Bits *t = pu; pu = pv; pv = t;
      t = p1; p1 = p2; p2 = t;
      const int td = du; du = dv; dv = td;
    }
    if (dv < 0) break;
    pu->xor_shifted(*pv, du - dv);
    p1->xor_shifted(*p2, du - dv);
  }
  // the loop ends with *pu == 0 (gcd = *pv, cofactor *p2) or *pv == 0 (gcd = *pu, cofactor *p1)
  const Bits *gc = pu->zero() ? pv : pu, *co = pu->zero() ? p2 : p1;
  if (gc->degree() != 0) return false;
  for (int i = 0; i < N; ++i) inv[i] = 0;
  const int dc = co->degree();
  for (int b = 0; b <= dc; ++b)
    if (co->get(b)) inv[b % N] ^= 1;          // x^N = 1
  return true;
}

// ---- host: inverse modulo (3, x^N - 1), extended Euclid over GF(3) -----------------------------------------------
struct Poly3 {
  std::vector<int8_t> c;       // coefficients in {0,1,2}
  int deg = -1;
  explicit Poly3(int n) : c(n, 0) {}
  void fix() {
    while (deg >= 0 && c[deg] == 0) --deg;
  }
};

// u -= k x^sh v  (mod 3)
void sub_scaled(Poly3 &u, const Poly3 &v, int k, int sh) {
  for (int i = 0; i <= v.deg; ++i) {
    int x = u.c[i + sh] - k * v.c[i];
    x %= 3;
    if (x < 0) x += 3;
    u.c[i + sh] = (int8_t)x;
  }
  if (v.deg + sh > u.deg) u.deg = v.deg + sh;
  u.fix();
}

bool inverse_mod3(const int8_t *f, int N, uint8_t *inv) {
  const int L = 2 * N + 4;
  Poly3 u(L), v(L), g1(L), g2(L);
  for (int i = 0; i < N; ++i) u.c[i] = (int8_t)(((f[i] % 3) + 3) % 3);
  u.deg = N - 1; u.fix();
  v.c[0] = 2; v.c[N] = 1; v.deg = N;           // x^N - 1
  g1.c[0] = 1; g1.deg = 0;
  Poly3 *pu = &u, *pv = &v, *p1 = &g1, *p2 = &g2;
  while (pu->deg >= 0) {
    if (pu->deg < pv->deg) {
      Poly3 *t = pu; pu = pv; pv = t;
      t = p1; p1 = p2; p2 = t;
    }
    if (pv->deg < 0) break;
    // leading coefficients in {1,2}: a / b mod 3 = a * b (1^-1 = 1, 2^-1 = 2)
    const int k = (pu->c[pu->deg] * pv->c[pv->deg]) % 3;
    const int sh = pu->deg - pv->deg;
    sub_scaled(*pu, *pv, k, sh);
    if (p2->deg >= 0) sub_scaled(*p1, *p2, k, sh);
  }
  const Poly3 *gc = pu->deg < 0 ? pv : pu, *co = pu->deg < 0 ? p2 : p1;
  if (gc->deg != 0) return false;
  const int ginv = gc->c[0];                    // 1 or 2, self-inverse
  for (int i = 0; i < N; ++i) inv[i] = 0;
  for (int b = 0; b <= co->deg; ++b) inv[b % N] = (uint8_t)((inv[b % N] + co->c[b] * ginv) % 3);
  return true;
}

// ---- device: the element-wise parts of a lifting step ------------------------------------------------------------
__global__ void k_newton_u(const uint16_t *__restrict__ t, size_t n, int P, int N, uint32_t qmask, int8_t *__restrict__ u0,
                           int8_t *__restrict__ u1) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % P);
    uint32_t u = k < N ? (((k == 0 ? 2u : 0u) - (uint32_t)t[i]) & qmask) : 0u;   // (2 - f inv) mod q, the 2 is the constant term
    u0[i] = (int8_t)(u & 127u);
    u1[i] = (int8_t)(u >> 7);
  }
}

__global__ void k_scale16(const uint16_t *__restrict__ src, size_t n, uint32_t mul, uint16_t *__restrict__ dst) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = (uint16_t)((uint32_t)src[i] * mul);
}

__global__ void k_newton_combine(const uint16_t *__restrict__ r0, const uint16_t *__restrict__ r1, size_t n, uint32_t qmask,
                                 uint16_t *__restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = (uint16_t)(((uint32_t)r0[i] + 128u * (uint32_t)r1[i]) & qmask);
}

}  // namespace

int keygen_batch(ntru_ctx *ctx, size_t B, const int8_t *f, const int8_t *g, uint16_t *fq, uint8_t *fp, uint16_t *h,
                 uint8_t *valid) {
  if (B == 0) return NTRU_OK;
  if (!ctx->tensor_ok || !imma_supported(ctx)) return fail(ctx, NTRU_E_UNSUPPORTED, "batched key generation needs the sm_100 tensor path and N <= 832");
  if (ctx->q < 4) return fail(ctx, NTRU_E_PARAM, "q too small");
  const int N = ctx->N, P = ctx->P;
  const uint32_t qmask = (uint32_t)ctx->q - 1;
  // ---- host: inverses modulo 2 and modulo p ----
  std::vector<uint16_t> inv_h(B * (size_t)P, 0);
  std::vector<int8_t> f_h(B * (size_t)P, 0), g_h(B * (size_t)P, 0);
  // keys are independent: the extended-Euclid runs are spread over the host's cores
  auto work = [&](size_t b0, size_t b1) {
    std::vector<uint8_t> i2(N), i3(N);
    for (size_t b = b0; b < b1; ++b) {
      const int8_t *fb = f + b * (size_t)N;
      bool ok = true;
      for (int i = 0; i < N; ++i) ok = ok && fb[i] >= -1 && fb[i] <= 1 && g[b * (size_t)N + i] >= -1 && g[b * (size_t)N + i] <= 1;
      ok = ok && inverse_mod2(fb, N, i2.data()) && inverse_mod3(fb, N, i3.data());
      valid[b] = ok ? 1 : 0;
      memset(fp + b * (size_t)N, 0, (size_t)N);
      if (!ok) continue;
      memcpy(fp + b * (size_t)N, i3.data(), (size_t)N);
      for (int i = 0; i < N; ++i) {
        inv_h[b * (size_t)P + i] = i2[i];
        f_h[b * (size_t)P + i] = fb[i];
        g_h[b * (size_t)P + i] = g[b * (size_t)N + i];
      }
    }
  };
  {
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > 64) nt = 64;
    if ((size_t)nt > B) nt = (unsigned)B;
    std::vector<std::thread> pool;
    const size_t per = (B + nt - 1) / nt;
    for (unsigned t = 1; t < nt; ++t) pool.emplace_back(work, (size_t)t * per < B ? (size_t)t * per : B, (size_t)(t + 1) * per < B ? (size_t)(t + 1) * per : B);
    work(0, per < B ? per : B);
    for (auto &th : pool) th.join();
  }
  // ---- device: lifting 2 -> q, then h ----
  const size_t n = B * (size_t)P;
  DevBuf &d_f = ctx->slot_bufs[0][0], &d_g = ctx->slot_bufs[0][1], &d_inv = ctx->slot_bufs[0][2], &d_t = ctx->slot_bufs[0][3],
         &d_u0 = ctx->slot_bufs[0][4], &d_u1 = ctx->slot_bufs[0][5], &d_r0 = ctx->slot_bufs[0][6], &d_r1 = ctx->slot_bufs[0][7];
  NTRU_CUDA(ctx, d_f.reserve(n)); NTRU_CUDA(ctx, d_g.reserve(n));
  NTRU_CUDA(ctx, d_inv.reserve(2 * n)); NTRU_CUDA(ctx, d_t.reserve(2 * n));
  NTRU_CUDA(ctx, d_u0.reserve(n)); NTRU_CUDA(ctx, d_u1.reserve(n));
  NTRU_CUDA(ctx, d_r0.reserve(2 * n)); NTRU_CUDA(ctx, d_r1.reserve(2 * n));
  cudaStream_t st = ctx->stream;
  NTRU_CUDA(ctx, cudaMemcpyAsync(d_f.ptr, f_h.data(), n, cudaMemcpyHostToDevice, st));
  NTRU_CUDA(ctx, cudaMemcpyAsync(d_g.ptr, g_h.data(), n, cudaMemcpyHostToDevice, st));
  NTRU_CUDA(ctx, cudaMemcpyAsync(d_inv.ptr, inv_h.data(), 2 * n, cudaMemcpyHostToDevice, st));
  int steps = 0;
  while ((1 << steps) < ctx->logq) ++steps;
  const unsigned grid = (unsigned)ctx->sm_count * 8;
  int rc;
  for (int s = 0; s < steps; ++s) {
    rc = launch_muldiv_imma(ctx, B, (const int8_t *)d_f.ptr, d_inv.ptr, 0, nullptr, d_t.ptr);              // t = f * inv
    if (rc) return rc;
    {
      LaunchTimer timer(ctx, NTRU_K_OTHER);
      k_newton_u<<<grid, 256, 0, st>>>((const uint16_t *)d_t.ptr, n, P, N, qmask, (int8_t *)d_u0.ptr, (int8_t *)d_u1.ptr);
    }
    rc = launch_muldiv_imma(ctx, B, (const int8_t *)d_u0.ptr, d_inv.ptr, 0, nullptr, d_r0.ptr);            // inv * u0
    if (rc) return rc;
    rc = launch_muldiv_imma(ctx, B, (const int8_t *)d_u1.ptr, d_inv.ptr, 0, nullptr, d_r1.ptr);            // inv * u1
    if (rc) return rc;
    {
      LaunchTimer timer(ctx, NTRU_K_OTHER);
      k_newton_combine<<<grid, 256, 0, st>>>((const uint16_t *)d_r0.ptr, (const uint16_t *)d_r1.ptr, n, qmask, (uint16_t *)d_inv.ptr);
    }
  }
  NTRU_CUDA(ctx, cudaGetLastError());
  // h = (p fq) * g mod (q, x^N - 1): y = p fq as uint16, un-reduced (p (q - 1) < 65536; the same product modulo q as the
  // reduced multiplyPolynomialsByScalar(fq, p, q) of index.js:75)
  {
    LaunchTimer timer(ctx, NTRU_K_OTHER);
    k_scale16<<<grid, 256, 0, st>>>((const uint16_t *)d_inv.ptr, n, (uint32_t)ctx->p, (uint16_t *)d_t.ptr);
  }
  NTRU_CUDA(ctx, cudaGetLastError());
  rc = launch_muldiv_imma(ctx, B, (const int8_t *)d_g.ptr, d_t.ptr, 0, nullptr, d_r0.ptr);
  if (rc) return rc;
  NTRU_CUDA(ctx, cudaMemcpy2DAsync(fq, (size_t)N * 2, d_inv.ptr, (size_t)P * 2, (size_t)N * 2, B, cudaMemcpyDeviceToHost, st));
  NTRU_CUDA(ctx, cudaMemcpy2DAsync(h, (size_t)N * 2, d_r0.ptr, (size_t)P * 2, (size_t)N * 2, B, cudaMemcpyDeviceToHost, st));
  NTRU_CUDA(ctx, cudaStreamSynchronize(st));
  for (size_t b = 0; b < B; ++b) {
    if (valid[b]) continue;
    memset(fq + b * (size_t)N, 0, (size_t)N * 2);
    memset(h + b * (size_t)N, 0, (size_t)N * 2);
  }
  return NTRU_OK;
}

}  // namespace ntru
