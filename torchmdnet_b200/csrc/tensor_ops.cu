// a8/a10: per-(atom, channel) irreducible-tensor algebra.
//
// (1) the reference's granular op set on its own layouts -- X [N,3,3,F]; I [N,1,F], A [N,3,F],
//     S [N,5,F] -- replacing torchmdnet/extensions/warp_kernels/{compose_tensor,decompose_tensor,
//     equivariant_o3_matmul,equivariant_so3_matmul,tensor_norm3}.py (fwd + bwd each);
// (2) the fused stages the B200 model path uses instead (P1/P2/P3/gate/norm), which keep the node
//     state in compact form end-to-end so one Interaction layer is 3 streaming kernels instead of
//     the reference's ~12 (models/tensornet.py:745-812).
// All kernels: one thread per (atom, channel), channel fastest => every load/store of a warp is
// one contiguous 128 B line; pure streaming, no shared memory, no atomics.  Arithmetic lives in
// tn_math.h (host-verified against autograd).
#include "common.cuh"

namespace {

struct CI { const float* I; const float* A; const float* S; };
struct MI { float* I; float* A; float* S; };

__device__ __forceinline__ void ld_c(const CI& t, long long n, int F, int h, float* c) {
  c[0] = t.I[n * F + h];
#pragma unroll
  for (int k = 0; k < 3; ++k) c[1 + k] = t.A[(n * 3 + k) * F + h];
#pragma unroll
  for (int k = 0; k < 5; ++k) c[4 + k] = t.S[(n * 5 + k) * F + h];
}
__device__ __forceinline__ void st_c(const MI& t, long long n, int F, int h, const float* c) {
  t.I[n * F + h] = c[0];
#pragma unroll
  for (int k = 0; k < 3; ++k) t.A[(n * 3 + k) * F + h] = c[1 + k];
#pragma unroll
  for (int k = 0; k < 5; ++k) t.S[(n * 5 + k) * F + h] = c[4 + k];
}
__device__ __forceinline__ void ld_x(const float* X, long long n, int F, int h, float* x) {
#pragma unroll
  for (int k = 0; k < 9; ++k) x[k] = X[(n * 9 + k) * F + h];
}
__device__ __forceinline__ void st_x(float* X, long long n, int F, int h, const float* x) {
#pragma unroll
  for (int k = 0; k < 9; ++k) X[(n * 9 + k) * F + h] = x[k];
}

#define TMD_FOR_NH(n_, h_)                                                                          \
  for (long long idx_ = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx_ < (long long)N * F; \
       idx_ += (long long)gridDim.x * blockDim.x)                                                   \
    if (long long n_ = idx_ / F; true)                                                              \
      if (int h_ = (int)(idx_ - n_ * F); true)

// ---- granular ops -------------------------------------------------------------------------------
__global__ void compose_fwd_k(CI c, float* X, int N, int F) {
  TMD_FOR_NH(n, h) { float cc[9], x[9]; ld_c(c, n, F, h, cc); tn::compose(cc, x); st_x(X, n, F, h, x); }
}
__global__ void compose_bwd_k(const float* dX, MI dc, int N, int F) {
  TMD_FOR_NH(n, h) { float x[9], cc[9]; ld_x(dX, n, F, h, x); tn::compose_bwd(x, cc); st_c(dc, n, F, h, cc); }
}
__global__ void decompose_fwd_k(const float* X, MI c, int N, int F) {
  TMD_FOR_NH(n, h) { float x[9], cc[9]; ld_x(X, n, F, h, x); tn::decompose(x, cc); st_c(c, n, F, h, cc); }
}
__global__ void decompose_bwd_k(CI dc, float* dX, int N, int F) {
  TMD_FOR_NH(n, h) { float cc[9], x[9]; ld_c(dc, n, F, h, cc); tn::decompose_bwd(cc, x); st_x(dX, n, F, h, x); }
}
template <bool O3> __global__ void matmul_fwd_k(const float* A, const float* B, float* C, int N, int F) {
  TMD_FOR_NH(n, h) {
    float a[9], b[9], c[9]; ld_x(A, n, F, h, a); ld_x(B, n, F, h, b);
    if (O3) tn::mm_o3(a, b, c); else tn::mm(a, b, c);
    st_x(C, n, F, h, c);
  }
}
template <bool O3> __global__ void matmul_bwd_k(const float* A, const float* B, const float* dC, float* dA, float* dB, int N, int F) {
  TMD_FOR_NH(n, h) {
    float a[9], b[9], d[9], da[9], db[9]; ld_x(A, n, F, h, a); ld_x(B, n, F, h, b); ld_x(dC, n, F, h, d);
    if (O3) tn::mm_o3_bwd(a, b, d, da, db); else tn::mm_so3_bwd(a, b, d, da, db);
    st_x(dA, n, F, h, da); st_x(dB, n, F, h, db);
  }
}
__global__ void norm3_fwd_k(const float* X, float* out, int N, int F) {
  TMD_FOR_NH(n, h) {
    float x[9], c[9], o[3]; ld_x(X, n, F, h, x); tn::decompose(x, c); tn::norm3_compact(c, o);
    out[n * 3 * F + h] = o[0]; out[n * 3 * F + F + h] = o[1]; out[n * 3 * F + 2 * F + h] = o[2];
  }
}
__global__ void norm3_bwd_k(const float* dout, const float* X, float* dX, int N, int F) {
  TMD_FOR_NH(n, h) {
    float x[9], c[9], dc[9], g[9];
    const float dn[3] = {dout[n * 3 * F + h], dout[n * 3 * F + F + h], dout[n * 3 * F + 2 * F + h]};
    ld_x(X, n, F, h, x); tn::decompose(x, c); tn::norm3_compact_bwd(c, dn, dc); tn::decompose_bwd(dc, g);
    st_x(dX, n, F, h, g);
  }
}

// ---- fused compact stages -----------------------------------------------------------------------
__global__ void norm3c_fwd_k(CI c, float* out, int N, int F) {
  TMD_FOR_NH(n, h) {
    float cc[9], o[3]; ld_c(c, n, F, h, cc); tn::norm3_compact(cc, o);
    out[n * 3 * F + h] = o[0]; out[n * 3 * F + F + h] = o[1]; out[n * 3 * F + 2 * F + h] = o[2];
  }
}
__global__ void norm3c_bwd_k(CI c, const float* dout, MI dc, int N, int F) {
  TMD_FOR_NH(n, h) {
    float cc[9], g[9];
    const float dn[3] = {dout[n * 3 * F + h], dout[n * 3 * F + F + h], dout[n * 3 * F + 2 * F + h]};
    ld_c(c, n, F, h, cc); tn::norm3_compact_bwd(cc, dn, g); st_c(dc, n, F, h, g);
  }
}
__global__ void norm2c_fwd_k(CI c, float* out, int N, int F) {
  TMD_FOR_NH(n, h) { float cc[9]; ld_c(c, n, F, h, cc); out[n * F + h] = tn::norm2_compact(cc); }
}
__global__ void norm2c_bwd_k(CI c, const float* dout, MI dc, int N, int F) {
  TMD_FOR_NH(n, h) {
    float cc[9], g[9]; ld_c(c, n, F, h, cc); tn::norm2_compact_grad(cc, g);
    const float d = dout[n * F + h];
#pragma unroll
    for (int k = 0; k < 9; ++k) g[k] *= d;
    st_c(dc, n, F, h, g);
  }
}
// gate: out[comp] = c[comp] * g[type(comp)],  g [N,3,F]       (tensornet.py:607-615)
__global__ void gate_fwd_k(CI c, const float* g, MI out, int N, int F) {
  TMD_FOR_NH(n, h) {
    float cc[9]; ld_c(c, n, F, h, cc);
    const float g0 = g[n * 3 * F + h], g1 = g[n * 3 * F + F + h], g2 = g[n * 3 * F + 2 * F + h];
    cc[0] *= g0;
#pragma unroll
    for (int k = 1; k < 4; ++k) cc[k] *= g1;
#pragma unroll
    for (int k = 4; k < 9; ++k) cc[k] *= g2;
    st_c(out, n, F, h, cc);
  }
}
__global__ void gate_bwd_k(CI c, const float* g, CI dout, MI dc, float* dg, int N, int F) {
  TMD_FOR_NH(n, h) {
    float cc[9], d[9]; ld_c(c, n, F, h, cc); ld_c(dout, n, F, h, d);
    const float g0 = g[n * 3 * F + h], g1 = g[n * 3 * F + F + h], g2 = g[n * 3 * F + 2 * F + h];
    float s0 = d[0] * cc[0], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 1; k < 4; ++k) s1 += d[k] * cc[k];
#pragma unroll
    for (int k = 4; k < 9; ++k) s2 += d[k] * cc[k];
    dg[n * 3 * F + h] = s0; dg[n * 3 * F + F + h] = s1; dg[n * 3 * F + 2 * F + h] = s2;
    d[0] *= g0;
#pragma unroll
    for (int k = 1; k < 4; ++k) d[k] *= g1;
#pragma unroll
    for (int k = 4; k < 9; ++k) d[k] *= g2;
    st_c(dc, n, F, h, d);
  }
}
// gate backward with the gate given as its pre-activation: g = SiLU(z) (z = z2 of the embedding's scalar MLP); chain != 0: dg is the
// gradient w.r.t. z (x SiLU'(z)), ready to be the A operand of the MLP's backward GEMM (csrc/rows_gemm.cu)
__global__ void gate_bwd_preact_k(CI c, const float* z, CI dout, MI dc, float* dg, int chain, int N, int F) {
  TMD_FOR_NH(n, h) {
    float cc[9], d[9]; ld_c(c, n, F, h, cc); ld_c(dout, n, F, h, d);
    const float z0 = z[n * 3 * F + h], z1 = z[n * 3 * F + F + h], z2 = z[n * 3 * F + 2 * F + h];
    const float g0 = tn::silu(z0), g1 = tn::silu(z1), g2 = tn::silu(z2);
    float s0 = d[0] * cc[0], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 1; k < 4; ++k) s1 += d[k] * cc[k];
#pragma unroll
    for (int k = 4; k < 9; ++k) s2 += d[k] * cc[k];
    if (chain) { s0 *= tn::silu_grad(z0); s1 *= tn::silu_grad(z1); s2 *= tn::silu_grad(z2); }   // gradient w.r.t. z instead of g
    dg[n * 3 * F + h] = s0; dg[n * 3 * F + F + h] = s1; dg[n * 3 * F + 2 * F + h] = s2;
    d[0] *= g0;
#pragma unroll
    for (int k = 1; k < 4; ++k) d[k] *= g1;
#pragma unroll
    for (int k = 4; k < 9; ++k) d[k] *= g2;
    st_c(dc, n, F, h, d);
  }
}
__global__ void p1_fwd_k(CI c, MI out, int N, int F) {
  TMD_FOR_NH(n, h) { float cc[9], o[9]; ld_c(c, n, F, h, cc); tn::p1_fwd(cc, o); st_c(out, n, F, h, o); }
}
__global__ void p1_bwd_k(CI c, CI dout, MI dc, int N, int F) {
  TMD_FOR_NH(n, h) { float cc[9], d[9], g[9]; ld_c(c, n, F, h, cc); ld_c(dout, n, F, h, d); tn::p1_bwd(cc, d, g); st_c(dc, n, F, h, g); }
}
// the same with the incoming gradient given as the sum of two tensors (the layer's own residual path + the node-linear path)
__global__ void p1_bwd2_k(CI c, CI d1, CI d2, MI dc, int N, int F) {
  TMD_FOR_NH(n, h) {
    float cc[9], a[9], b[9], g[9]; ld_c(c, n, F, h, cc); ld_c(d1, n, F, h, a); ld_c(d2, n, F, h, b);
#pragma unroll
    for (int k = 0; k < 9; ++k) a[k] += b[k];
    tn::p1_bwd(cc, a, g); st_c(dc, n, F, h, g);
  }
}
__global__ void p2_fwd_k(CI Y, CI M, const float* qf, int o3, MI out, int N, int F) {
  TMD_FOR_NH(n, h) {
    float y[9], m[9], o[9]; ld_c(Y, n, F, h, y); ld_c(M, n, F, h, m);
    tn::p2_fwd(y, m, qf[n], o3 != 0, o); st_c(out, n, F, h, o);
  }
}
__global__ void p2_bwd_k(CI Y, CI M, const float* qf, int o3, CI dout, MI dY, MI dM, int N, int F) {
  TMD_FOR_NH(n, h) {
    float y[9], m[9], d[9], gy[9], gm[9]; ld_c(Y, n, F, h, y); ld_c(M, n, F, h, m); ld_c(dout, n, F, h, d);
    tn::p2_bwd(y, m, qf[n], o3 != 0, d, gy, gm); st_c(dY, n, F, h, gy); st_c(dM, n, F, h, gm);
  }
}
__global__ void p3_fwd_k(CI Xn, CI dX, const float* qf, MI out, int N, int F) {
  TMD_FOR_NH(n, h) {
    float a[9], b[9], o[9]; ld_c(Xn, n, F, h, a); ld_c(dX, n, F, h, b);
    tn::p3_fwd(a, b, qf[n], o); st_c(out, n, F, h, o);
  }
}
__global__ void p3_bwd_k(CI dX, const float* qf, CI dout, MI gXn, MI gdX, int N, int F) {
  TMD_FOR_NH(n, h) {
    float b[9], d[9], g1[9], g2[9]; ld_c(dX, n, F, h, b); ld_c(dout, n, F, h, d);
    tn::p3_bwd(b, qf[n], d, g1, g2); st_c(gXn, n, F, h, g1); st_c(gdX, n, F, h, g2);
  }
}

// P3 backward, dX branch only (the Xn branch of X' = Xn + dX + qf dX.dX is the identity: the caller passes dout on)
__global__ void p3_bwd_dx_k(CI dX, const float* qf, CI dout, MI gdX, int N, int F) {
  TMD_FOR_NH(n, h) {
    float b[9], d[9], g1[9], g2[9]; ld_c(dX, n, F, h, b); ld_c(dout, n, F, h, d);
    tn::p3_bwd(b, qf[n], d, g1, g2); st_c(gdX, n, F, h, g2);
  }
}

inline int grid_nf(int N, int F) { return tmd_blocks_for((long long)N * F, 256, 148 * 32); }

}  // namespace

#define TMD_LAUNCH_NF(kernel, ...)                                                                \
  do {                                                                                            \
    TMD_REQUIRE(N >= 0 && F > 0, #kernel ": bad sizes");                                          \
    if (N == 0) return TMD_OK;                                                                    \
    kernel<<<grid_nf(N, F), 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__, N, F);                   \
    return tmd_check_launch(#kernel);                                                             \
  } while (0)

extern "C" {
int tmd_compose_fwd(const float* I, const float* A, const float* S, float* X, int N, int F, void* stream) {
  TMD_REQUIRE(I && A && S && X, "compose_fwd: null pointer"); TMD_LAUNCH_NF(compose_fwd_k, CI{I, A, S}, X); }
int tmd_compose_bwd(const float* dX, float* dI, float* dA, float* dS, int N, int F, void* stream) {
  TMD_REQUIRE(dX && dI && dA && dS, "compose_bwd: null pointer"); TMD_LAUNCH_NF(compose_bwd_k, dX, MI{dI, dA, dS}); }
int tmd_decompose_fwd(const float* X, float* I, float* A, float* S, int N, int F, void* stream) {
  TMD_REQUIRE(X && I && A && S, "decompose_fwd: null pointer"); TMD_LAUNCH_NF(decompose_fwd_k, X, MI{I, A, S}); }
int tmd_decompose_bwd(const float* dI, const float* dA, const float* dS, float* dX, int N, int F, void* stream) {
  TMD_REQUIRE(dX && dI && dA && dS, "decompose_bwd: null pointer"); TMD_LAUNCH_NF(decompose_bwd_k, CI{dI, dA, dS}, dX); }
int tmd_matmul_o3_fwd(const float* A, const float* B, float* C, int N, int F, void* stream) {
  TMD_REQUIRE(A && B && C, "matmul_o3_fwd: null pointer"); TMD_LAUNCH_NF(matmul_fwd_k<true>, A, B, C); }
int tmd_matmul_o3_bwd(const float* A, const float* B, const float* dC, float* dA, float* dB, int N, int F, void* stream) {
  TMD_REQUIRE(A && B && dC && dA && dB, "matmul_o3_bwd: null pointer"); TMD_LAUNCH_NF(matmul_bwd_k<true>, A, B, dC, dA, dB); }
int tmd_matmul_so3_fwd(const float* A, const float* B, float* C, int N, int F, void* stream) {
  TMD_REQUIRE(A && B && C, "matmul_so3_fwd: null pointer"); TMD_LAUNCH_NF(matmul_fwd_k<false>, A, B, C); }
int tmd_matmul_so3_bwd(const float* A, const float* B, const float* dC, float* dA, float* dB, int N, int F, void* stream) {
  TMD_REQUIRE(A && B && dC && dA && dB, "matmul_so3_bwd: null pointer"); TMD_LAUNCH_NF(matmul_bwd_k<false>, A, B, dC, dA, dB); }
int tmd_norm3_fwd(const float* X, float* out, int N, int F, void* stream) {
  TMD_REQUIRE(X && out, "norm3_fwd: null pointer"); TMD_LAUNCH_NF(norm3_fwd_k, X, out); }
int tmd_norm3_bwd(const float* dout, const float* X, float* dX, int N, int F, void* stream) {
  TMD_REQUIRE(dout && X && dX, "norm3_bwd: null pointer"); TMD_LAUNCH_NF(norm3_bwd_k, dout, X, dX); }

int tmd_norm3c_fwd(const float* I, const float* A, const float* S, float* out, int N, int F, void* stream) {
  TMD_REQUIRE(I && A && S && out, "norm3c_fwd: null pointer"); TMD_LAUNCH_NF(norm3c_fwd_k, CI{I, A, S}, out); }
int tmd_norm3c_bwd(const float* I, const float* A, const float* S, const float* dout, float* dI, float* dA, float* dS, int N, int F, void* stream) {
  TMD_REQUIRE(I && A && S && dout && dI && dA && dS, "norm3c_bwd: null pointer"); TMD_LAUNCH_NF(norm3c_bwd_k, CI{I, A, S}, dout, MI{dI, dA, dS}); }
int tmd_norm2c_fwd(const float* I, const float* A, const float* S, float* out, int N, int F, void* stream) {
  TMD_REQUIRE(I && A && S && out, "norm2c_fwd: null pointer"); TMD_LAUNCH_NF(norm2c_fwd_k, CI{I, A, S}, out); }
int tmd_norm2c_bwd(const float* I, const float* A, const float* S, const float* dout, float* dI, float* dA, float* dS, int N, int F, void* stream) {
  TMD_REQUIRE(I && A && S && dout && dI && dA && dS, "norm2c_bwd: null pointer"); TMD_LAUNCH_NF(norm2c_bwd_k, CI{I, A, S}, dout, MI{dI, dA, dS}); }
int tmd_gate_fwd(const float* I, const float* A, const float* S, const float* g, float* oI, float* oA, float* oS, int N, int F, void* stream) {
  TMD_REQUIRE(I && A && S && g && oI && oA && oS, "gate_fwd: null pointer"); TMD_LAUNCH_NF(gate_fwd_k, CI{I, A, S}, g, MI{oI, oA, oS}); }
int tmd_gate_bwd(const float* I, const float* A, const float* S, const float* g, const float* dI, const float* dA, const float* dS,
                 float* gI, float* gA, float* gS, float* dg, int N, int F, void* stream) {
  TMD_REQUIRE(I && A && S && g && dI && dA && dS && gI && gA && gS && dg, "gate_bwd: null pointer");
  TMD_LAUNCH_NF(gate_bwd_k, CI{I, A, S}, g, CI{dI, dA, dS}, MI{gI, gA, gS}, dg); }
int tmd_gate_bwd_preact(const float* I, const float* A, const float* S, const float* z, const float* dI, const float* dA, const float* dS,
                        float* gI, float* gA, float* gS, float* dg, int chain, int N, int F, void* stream) {
  TMD_REQUIRE(I && A && S && z && dI && dA && dS && gI && gA && gS && dg, "gate_bwd_preact: null pointer");
  TMD_LAUNCH_NF(gate_bwd_preact_k, CI{I, A, S}, z, CI{dI, dA, dS}, MI{gI, gA, gS}, dg, chain); }
int tmd_p1_bwd2(const float* I, const float* A, const float* S, const float* d1I, const float* d1A, const float* d1S, const float* d2I,
                const float* d2A, const float* d2S, float* gI, float* gA, float* gS, int N, int F, void* stream) {
  TMD_REQUIRE(I && A && S && d1I && d1A && d1S && d2I && d2A && d2S && gI && gA && gS, "p1_bwd2: null pointer");
  TMD_LAUNCH_NF(p1_bwd2_k, CI{I, A, S}, CI{d1I, d1A, d1S}, CI{d2I, d2A, d2S}, MI{gI, gA, gS}); }
int tmd_p3_bwd_dx(const float* DI, const float* DA, const float* DS, const float* qf, const float* dI, const float* dA, const float* dS,
                  float* gDI, float* gDA, float* gDS, int N, int F, void* stream) {
  TMD_REQUIRE(DI && DA && DS && qf && dI && dA && dS && gDI && gDA && gDS, "p3_bwd_dx: null pointer");
  TMD_LAUNCH_NF(p3_bwd_dx_k, CI{DI, DA, DS}, qf, CI{dI, dA, dS}, MI{gDI, gDA, gDS}); }
int tmd_p1_fwd(const float* I, const float* A, const float* S, float* oI, float* oA, float* oS, int N, int F, void* stream) {
  TMD_REQUIRE(I && A && S && oI && oA && oS, "p1_fwd: null pointer"); TMD_LAUNCH_NF(p1_fwd_k, CI{I, A, S}, MI{oI, oA, oS}); }
int tmd_p1_bwd(const float* I, const float* A, const float* S, const float* dI, const float* dA, const float* dS,
               float* gI, float* gA, float* gS, int N, int F, void* stream) {
  TMD_REQUIRE(I && A && S && dI && dA && dS && gI && gA && gS, "p1_bwd: null pointer");
  TMD_LAUNCH_NF(p1_bwd_k, CI{I, A, S}, CI{dI, dA, dS}, MI{gI, gA, gS}); }
int tmd_p2_fwd(const float* YI, const float* YA, const float* YS, const float* MI_, const float* MA, const float* MS,
               const float* qf, int o3, float* oI, float* oA, float* oS, int N, int F, void* stream) {
  TMD_REQUIRE(YI && YA && YS && MI_ && MA && MS && qf && oI && oA && oS, "p2_fwd: null pointer");
  TMD_LAUNCH_NF(p2_fwd_k, CI{YI, YA, YS}, CI{MI_, MA, MS}, qf, o3, MI{oI, oA, oS}); }
int tmd_p2_bwd(const float* YI, const float* YA, const float* YS, const float* MI_, const float* MA, const float* MS,
               const float* qf, int o3, const float* dI, const float* dA, const float* dS, float* gYI, float* gYA,
               float* gYS, float* gMI, float* gMA, float* gMS, int N, int F, void* stream) {
  TMD_REQUIRE(YI && YA && YS && MI_ && MA && MS && qf && dI && dA && dS && gYI && gYA && gYS && gMI && gMA && gMS, "p2_bwd: null pointer");
  TMD_LAUNCH_NF(p2_bwd_k, CI{YI, YA, YS}, CI{MI_, MA, MS}, qf, o3, CI{dI, dA, dS}, MI{gYI, gYA, gYS}, MI{gMI, gMA, gMS}); }
int tmd_p3_fwd(const float* XI, const float* XA, const float* XS, const float* DI, const float* DA, const float* DS,
               const float* qf, float* oI, float* oA, float* oS, int N, int F, void* stream) {
  TMD_REQUIRE(XI && XA && XS && DI && DA && DS && qf && oI && oA && oS, "p3_fwd: null pointer");
  TMD_LAUNCH_NF(p3_fwd_k, CI{XI, XA, XS}, CI{DI, DA, DS}, qf, MI{oI, oA, oS}); }
int tmd_p3_bwd(const float* DI, const float* DA, const float* DS, const float* qf, const float* dI, const float* dA,
               const float* dS, float* gXI, float* gXA, float* gXS, float* gDI, float* gDA, float* gDS, int N, int F, void* stream) {
  TMD_REQUIRE(DI && DA && DS && qf && dI && dA && dS && gXI && gXA && gXS && gDI && gDA && gDS, "p3_bwd: null pointer");
  TMD_LAUNCH_NF(p3_bwd_k, CI{DI, DA, DS}, qf, CI{dI, dA, dS}, MI{gXI, gXA, gXS}, MI{gDI, gDA, gDS}); }
}  // extern "C"
