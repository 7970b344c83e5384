// TEST INFRASTRUCTURE ONLY: compiles torchmdnet_b200/csrc/tn_math.h with g++ (double precision) so
// tests/test_tn_math_host.py can verify every hand-derived backward against torch.autograd on the
// CPU-only build box.  Not part of the product library and never loaded by torchmdnet_b200.
#include "../../torchmdnet_b200/csrc/tn_math.h"
#include <string.h>
typedef double T;
extern "C" {
// each op works on `n` independent items; inputs/outputs are packed row-major per item
void hc_compose(const T* c, T* X, int n) { for (int i = 0; i < n; ++i) tn::compose(c + 9 * i, X + 9 * i); }
void hc_compose_bwd(const T* dX, T* dc, int n) { for (int i = 0; i < n; ++i) tn::compose_bwd(dX + 9 * i, dc + 9 * i); }
void hc_decompose(const T* X, T* c, int n) { for (int i = 0; i < n; ++i) tn::decompose(X + 9 * i, c + 9 * i); }
void hc_decompose_bwd(const T* dc, T* dX, int n) { for (int i = 0; i < n; ++i) tn::decompose_bwd(dc + 9 * i, dX + 9 * i); }
void hc_mm_o3(const T* A, const T* B, T* C, int n) { for (int i = 0; i < n; ++i) tn::mm_o3(A + 9 * i, B + 9 * i, C + 9 * i); }
void hc_mm_o3_bwd(const T* A, const T* B, const T* dC, T* dA, T* dB, int n) {
  for (int i = 0; i < n; ++i) tn::mm_o3_bwd(A + 9 * i, B + 9 * i, dC + 9 * i, dA + 9 * i, dB + 9 * i); }
void hc_mm_so3(const T* A, const T* B, T* C, int n) { for (int i = 0; i < n; ++i) tn::mm(A + 9 * i, B + 9 * i, C + 9 * i); }
void hc_mm_so3_bwd(const T* A, const T* B, const T* dC, T* dA, T* dB, int n) {
  for (int i = 0; i < n; ++i) tn::mm_so3_bwd(A + 9 * i, B + 9 * i, dC + 9 * i, dA + 9 * i, dB + 9 * i); }
void hc_norm3(const T* c, T* o, int n) { for (int i = 0; i < n; ++i) tn::norm3_compact(c + 9 * i, o + 3 * i); }
void hc_norm3_bwd(const T* c, const T* dn, T* dc, int n) { for (int i = 0; i < n; ++i) tn::norm3_compact_bwd(c + 9 * i, dn + 3 * i, dc + 9 * i); }
void hc_norm2(const T* c, T* o, int n) { for (int i = 0; i < n; ++i) o[i] = tn::norm2_compact(c + 9 * i); }
void hc_p1(const T* c, T* o, int n) { for (int i = 0; i < n; ++i) tn::p1_fwd(c + 9 * i, o + 9 * i); }
void hc_p1_bwd(const T* c, const T* d, T* dc, int n) { for (int i = 0; i < n; ++i) tn::p1_bwd(c + 9 * i, d + 9 * i, dc + 9 * i); }
void hc_p2(const T* Y, const T* M, const T* qf, int o3, T* o, int n) {
  for (int i = 0; i < n; ++i) tn::p2_fwd(Y + 9 * i, M + 9 * i, qf[i], o3 != 0, o + 9 * i); }
void hc_p2_bwd(const T* Y, const T* M, const T* qf, int o3, const T* d, T* dY, T* dM, int n) {
  for (int i = 0; i < n; ++i) tn::p2_bwd(Y + 9 * i, M + 9 * i, qf[i], o3 != 0, d + 9 * i, dY + 9 * i, dM + 9 * i); }
void hc_p3(const T* Xn, const T* dX, const T* qf, T* o, int n) {
  for (int i = 0; i < n; ++i) tn::p3_fwd(Xn + 9 * i, dX + 9 * i, qf[i], o + 9 * i); }
void hc_p3_bwd(const T* dX, const T* qf, const T* d, T* dXn, T* ddX, int n) {
  for (int i = 0; i < n; ++i) tn::p3_bwd(dX + 9 * i, qf[i], d + 9 * i, dXn + 9 * i, ddX + 9 * i); }
void hc_silu(const T* x, T* v, T* g, int n) { for (int i = 0; i < n; ++i) { v[i] = tn::silu(x[i]); g[i] = tn::silu_grad(x[i]); } }
void hc_cutoff(const T* r, T lo, T hi, T* c, T* dc, int n) { for (int i = 0; i < n; ++i) tn::cosine_cutoff(r[i], lo, hi, c + i, dc + i); }
// rbf[n,R] and d rbf/dr
void hc_expnorm(const T* r, T lo, T hi, const T* mu, const T* beta, int R, T* v, T* dv, int n) {
  const T alpha = 5.0 / (hi - lo);
  for (int i = 0; i < n; ++i) {
    T c0, dc0; tn::cosine_cutoff(r[i], T(0), hi, &c0, &dc0);
    const T u = exp(alpha * (lo - r[i]));
    for (int k = 0; k < R; ++k) tn::expnorm_rbf(u, alpha, c0, dc0, mu[k], beta[k], v + i * R + k, dv + i * R + k);
  }
}
void hc_sym_basis(const T* r, T* s, int n) { for (int i = 0; i < n; ++i) tn::sym_basis(r + 3 * i, s + 5 * i); }
void hc_sym_basis_bwd(const T* r, const T* dS, const T* w, T* gr, int n) {
  for (int i = 0; i < n; ++i) { gr[3 * i] = gr[3 * i + 1] = gr[3 * i + 2] = 0; tn::sym_basis_bwd_acc(r + 3 * i, dS + 5 * i, w[i], gr + 3 * i); } }
}
