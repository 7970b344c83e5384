// Per-(atom, channel) arithmetic of the TensorNet hot path, written once as
// __host__ __device__ inline functions so the same source is (a) inlined into the sm_100a
// kernels and (b) compiled by g++ into tests/hostcheck to verify every hand-derived
// backward against torch.autograd on the CPU-only build box.
//
// Irreducible ("compact") layout, same component order as the reference's op boundary
// (torchmdnet/extensions/warp_kernels/compose_tensor.py:54-98, decompose_tensor.py:56-90):
//   c[0]      = I            (trace/3)
//   c[1..3]   = A01, A02, A12           antisymmetric part, upper triangle
//   c[4..8]   = S00, S01, S02, S11, S12 symmetric traceless part (S22 = -(S00+S11))
// Full layout X[9] is row-major 3x3.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define TN_HD __host__ __device__ __forceinline__
#else
#define TN_HD inline
#endif

namespace tn {

// ---------------------------------------------------------------- compose / decompose
// reference: models/tensornet.py:18-39, :125-141
template <typename T> TN_HD void compose(const T* c, T* X) {
  X[0] = c[0] + c[4];  X[1] = c[1] + c[5];   X[2] = c[2] + c[6];
  X[3] = -c[1] + c[5]; X[4] = c[0] + c[7];   X[5] = c[3] + c[8];
  X[6] = -c[2] + c[6]; X[7] = -c[3] + c[8];  X[8] = c[0] - (c[4] + c[7]);
}
// dc = (d compose / d c)^T dX
template <typename T> TN_HD void compose_bwd(const T* dX, T* dc) {
  dc[0] = dX[0] + dX[4] + dX[8];
  dc[1] = dX[1] - dX[3]; dc[2] = dX[2] - dX[6]; dc[3] = dX[5] - dX[7];
  dc[4] = dX[0] - dX[8]; dc[5] = dX[1] + dX[3]; dc[6] = dX[2] + dX[6];
  dc[7] = dX[4] - dX[8]; dc[8] = dX[5] + dX[7];
}
template <typename T> TN_HD void decompose(const T* X, T* c) {
  const T tr3 = (X[0] + X[4] + X[8]) / T(3);
  c[0] = tr3;
  c[1] = (X[1] - X[3]) * T(0.5); c[2] = (X[2] - X[6]) * T(0.5); c[3] = (X[5] - X[7]) * T(0.5);
  c[4] = X[0] - tr3; c[5] = (X[1] + X[3]) * T(0.5); c[6] = (X[2] + X[6]) * T(0.5);
  c[7] = X[4] - tr3; c[8] = (X[5] + X[7]) * T(0.5);
}
template <typename T> TN_HD void decompose_bwd(const T* dc, T* dX) {
  const T t = (dc[0] - dc[4] - dc[7]) / T(3);
  dX[0] = t + dc[4]; dX[4] = t + dc[7]; dX[8] = t;
  dX[1] = (dc[5] + dc[1]) * T(0.5); dX[3] = (dc[5] - dc[1]) * T(0.5);
  dX[2] = (dc[6] + dc[2]) * T(0.5); dX[6] = (dc[6] - dc[2]) * T(0.5);
  dX[5] = (dc[8] + dc[3]) * T(0.5); dX[7] = (dc[8] - dc[3]) * T(0.5);
}

// ---------------------------------------------------------------- 3x3 products
template <typename T> TN_HD void mm(const T* A, const T* B, T* C) {  // C = A B
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
template <typename T> TN_HD void mm_acc(const T* A, const T* B, T* C) {  // C += A B
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] += A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
template <typename T> TN_HD void mm_nt_acc(const T* A, const T* B, T* C) {  // C += A B^T
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] += A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
template <typename T> TN_HD void mm_tn_acc(const T* A, const T* B, T* C) {  // C += A^T B
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] += A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
// reference: tensornet.py:42-51 (_tensor_matmul_o3 = A B + B A, _tensor_matmul_so3 = A B)
template <typename T> TN_HD void mm_o3(const T* A, const T* B, T* C) { mm(A, B, C); mm_acc(B, A, C); }
template <typename T> TN_HD void mm_o3_bwd(const T* A, const T* B, const T* dC, T* dA, T* dB) {
#pragma unroll
  for (int k = 0; k < 9; ++k) { dA[k] = T(0); dB[k] = T(0); }
  mm_nt_acc(dC, B, dA); mm_tn_acc(B, dC, dA);  // dA = dC B^T + B^T dC
  mm_tn_acc(A, dC, dB); mm_nt_acc(dC, A, dB);  // dB = A^T dC + dC A^T
}
template <typename T> TN_HD void mm_so3_bwd(const T* A, const T* B, const T* dC, T* dA, T* dB) {
#pragma unroll
  for (int k = 0; k < 9; ++k) { dA[k] = T(0); dB[k] = T(0); }
  mm_nt_acc(dC, B, dA);  // dA = dC B^T
  mm_tn_acc(A, dC, dB);  // dB = A^T dC
}

// ---------------------------------------------------------------- norms
// squared Frobenius norm of the composed tensor (tensornet.py:144-146), from compact form
template <typename T> TN_HD T norm2_compact(const T* c) {
  const T s22 = c[4] + c[7];
  return T(3) * c[0] * c[0] + T(2) * (c[1] * c[1] + c[2] * c[2] + c[3] * c[3]) +
         (c[4] * c[4] + c[7] * c[7] + s22 * s22 + T(2) * (c[5] * c[5] + c[6] * c[6] + c[8] * c[8]));
}
template <typename T> TN_HD void norm2_compact_grad(const T* c, T* g) {  // d norm2 / d c
  g[0] = T(6) * c[0];
  g[1] = T(4) * c[1]; g[2] = T(4) * c[2]; g[3] = T(4) * c[3];
  g[4] = T(4) * c[4] + T(2) * c[7]; g[5] = T(4) * c[5]; g[6] = T(4) * c[6];
  g[7] = T(4) * c[7] + T(2) * c[4]; g[8] = T(4) * c[8];
}
// the three irrep norms of tensor_norm3 (warp_kernels/tensor_norm3.py:48-81) = (3 I^2, |A|^2, |S|^2)
template <typename T> TN_HD void norm3_compact(const T* c, T* n) {
  const T s22 = c[4] + c[7];
  n[0] = T(3) * c[0] * c[0];
  n[1] = T(2) * (c[1] * c[1] + c[2] * c[2] + c[3] * c[3]);
  n[2] = c[4] * c[4] + c[7] * c[7] + s22 * s22 + T(2) * (c[5] * c[5] + c[6] * c[6] + c[8] * c[8]);
}
template <typename T> TN_HD void norm3_compact_bwd(const T* c, const T* dn, T* dc) {
  dc[0] = T(6) * c[0] * dn[0];
  dc[1] = T(4) * c[1] * dn[1]; dc[2] = T(4) * c[2] * dn[1]; dc[3] = T(4) * c[3] * dn[1];
  dc[4] = (T(4) * c[4] + T(2) * c[7]) * dn[2]; dc[5] = T(4) * c[5] * dn[2]; dc[6] = T(4) * c[6] * dn[2];
  dc[7] = (T(4) * c[7] + T(2) * c[4]) * dn[2]; dc[8] = T(4) * c[8] * dn[2];
}

// ---------------------------------------------------------------- fused per-atom stages of Interaction
// P1: Xn = X / (|X|^2 + 1)                                     (tensornet.py:745)
template <typename T> TN_HD void p1_fwd(const T* c, T* out) {
  const T s = T(1) / (norm2_compact(c) + T(1));
#pragma unroll
  for (int k = 0; k < 9; ++k) out[k] = c[k] * s;
}
template <typename T> TN_HD void p1_bwd(const T* c, const T* dout, T* dc) {
  const T s = T(1) / (norm2_compact(c) + T(1));
  T g[9]; norm2_compact_grad(c, g);
  T dot = T(0);
#pragma unroll
  for (int k = 0; k < 9; ++k) dot += dout[k] * c[k];
  const T f = -s * s * dot;
#pragma unroll
  for (int k = 0; k < 9; ++k) dc[k] = s * dout[k] + f * g[k];
}
// P2: C = qf (Y M + M Y)  [O(3)]   or   C = 2 Y M  [SO(3)];  out = decompose(C) / (|C|^2 + 1)
//                                                              (tensornet.py:786-806)
template <typename T> TN_HD void p2_fwd(const T* Y, const T* M, T qf, bool o3, T* out) {
  T Yf[9], Mf[9], C[9], cc[9];
  compose(Y, Yf); compose(M, Mf);
  if (o3) { mm_o3(Yf, Mf, C); for (int k = 0; k < 9; ++k) C[k] *= qf; }
  else { mm(Yf, Mf, C); for (int k = 0; k < 9; ++k) C[k] *= T(2); }
  decompose(C, cc);
  const T s = T(1) / (norm2_compact(cc) + T(1));
#pragma unroll
  for (int k = 0; k < 9; ++k) out[k] = cc[k] * s;
}
template <typename T> TN_HD void p2_bwd(const T* Y, const T* M, T qf, bool o3, const T* dout, T* dY, T* dM) {
  T Yf[9], Mf[9], C[9], cc[9], dcc[9], dC[9], dYf[9], dMf[9];
  compose(Y, Yf); compose(M, Mf);
  const T sc = o3 ? qf : T(2);
  if (o3) mm_o3(Yf, Mf, C); else mm(Yf, Mf, C);
#pragma unroll
  for (int k = 0; k < 9; ++k) C[k] *= sc;
  decompose(C, cc);
  p1_bwd(cc, dout, dcc);
  decompose_bwd(dcc, dC);
#pragma unroll
  for (int k = 0; k < 9; ++k) dC[k] *= sc;
  if (o3) mm_o3_bwd(Yf, Mf, dC, dYf, dMf); else mm_so3_bwd(Yf, Mf, dC, dYf, dMf);
  compose_bwd(dYf, dY); compose_bwd(dMf, dM);
}
// P3: X' = Xn + dX + qf * (dX dX)                               (tensornet.py:812)
template <typename T> TN_HD void p3_fwd(const T* Xn, const T* dX, T qf, T* out) {
  T D[9], Q[9], qc[9];
  compose(dX, D); mm(D, D, Q); decompose(Q, qc);
#pragma unroll
  for (int k = 0; k < 9; ++k) out[k] = Xn[k] + dX[k] + qf * qc[k];
}
template <typename T> TN_HD void p3_bwd(const T* dX, T qf, const T* dout, T* dXn, T* ddX) {
  T D[9], G[9], dD[9], t[9];
  compose(dX, D);
  decompose_bwd(dout, G);
#pragma unroll
  for (int k = 0; k < 9; ++k) { G[k] *= qf; dD[k] = T(0); }
  mm_nt_acc(G, D, dD); mm_tn_acc(D, G, dD);  // d(D D) -> G D^T + D^T G
  compose_bwd(dD, t);
#pragma unroll
  for (int k = 0; k < 9; ++k) { dXn[k] = dout[k]; ddX[k] = dout[k] + t[k]; }
}

// ---------------------------------------------------------------- activations
template <typename T> TN_HD T sigmoid_(T x) { return T(1) / (T(1) + exp(-x)); }
TN_HD float sigmoid_(float x) { return 1.0f / (1.0f + expf(-x)); }
template <typename T> TN_HD T silu(T x) { return x * sigmoid_(x); }
template <typename T> TN_HD T silu_grad(T x) { const T s = sigmoid_(x); return s * (T(1) + x * (T(1) - s)); }

// ---------------------------------------------------------------- radial functions
#define TN_PI 3.14159265358979323846
// CosineCutoff (models/utils.py:500-528): value and d/dr
template <typename T> TN_HD void cosine_cutoff(T r, T lo, T hi, T* c, T* dc) {
  if (lo > T(0)) {
    const T k = T(2) * T(TN_PI) / (hi - lo);
    const T arg = T(TN_PI) * (T(2) * (r - lo) / (hi - lo) + T(1));
    const bool in = (r < hi) && (r > lo);
    *c = in ? T(0.5) * (cos(arg) + T(1)) : T(0);
    *dc = in ? -T(0.5) * sin(arg) * k : T(0);
  } else {
    const T k = T(TN_PI) / hi;
    const bool in = r < hi;
    *c = in ? T(0.5) * (cos(r * k) + T(1)) : T(0);
    *dc = in ? -T(0.5) * sin(r * k) * k : T(0);
  }
}
// ExpNormalSmearing (models/utils.py:356-407): rbf_k = C0(r) exp(-beta_k (exp(alpha (lo - r)) - mu_k)^2)
// u = exp(alpha (lo - r)), c0/dc0 = CosineCutoff(0, hi) and derivative
template <typename T> TN_HD void expnorm_rbf(T u, T alpha, T c0, T dc0, T mu, T beta, T* v, T* dv) {
  const T t = u - mu;
  const T e = exp(-beta * t * t);
  *v = c0 * e;
  *dv = dc0 * e + T(2) * alpha * beta * u * t * c0 * e;
}

// ---------------------------------------------------------------- embedding edge geometry
// s(rhat) = symmetric-traceless basis of rhat rhat^T in compact order; (tensornet.py:133-141, :405-445)
template <typename T> TN_HD void sym_basis(const T* r, T* s) {
  const T m = (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) / T(3);
  s[0] = r[0] * r[0] - m; s[1] = r[0] * r[1]; s[2] = r[0] * r[2]; s[3] = r[1] * r[1] - m; s[4] = r[1] * r[2];
}
// g_r += w3 * d(sum_k dS_k s_k)/d rhat
template <typename T> TN_HD void sym_basis_bwd_acc(const T* r, const T* dS, T w, T* gr) {
  const T a = T(2) / T(3) * (dS[0] + dS[3]);
  gr[0] += w * (T(2) * dS[0] * r[0] + dS[1] * r[1] + dS[2] * r[2] - a * r[0]);
  gr[1] += w * (T(2) * dS[3] * r[1] + dS[1] * r[0] + dS[4] * r[2] - a * r[1]);
  gr[2] += w * (dS[2] * r[0] + dS[4] * r[1] - a * r[2]);
}
// skew(v) in compact order: A01 = -vz, A02 = +vy, A12 = -vx  (tensornet.py:84-103)
template <typename T> TN_HD void skew_compact(const T* v, T* a) { a[0] = -v[2]; a[1] = v[1]; a[2] = -v[0]; }

}  // namespace tn
