// a5, per-atom part of TensorEmbedding (models/tensornet.py:588-619): the scalar gate of the initial tensors,
//   g = SiLU(W2 SiLU(W1 LayerNorm_F(|X|^2) + b1) + b2)  in  R^{3F},      X'_c = (W_t X)_c * g[type(c)]
// with |X|^2 the squared Frobenius norm per channel of the un-normalised embedding X = (I, A, S), W_t the three node linears
// (irreps_linear_kernel) and type(c) = 0 / 1 / 2 for the I / A / S components.  The reference runs tensor_norm, LayerNorm, two
// cuBLAS GEMMs, two SiLU launches, three slices/multiplies and their autograd mirrors; in round 1 of this repo it was
// norm2c + torch LayerNorm + 2 cuBLAS SIMT SGEMMs + 2 SiLU + gate (and 10 launches in the backward).
// Here: ONE forward kernel per 32-atom tile (norm -> LN -> two tile GEMMs, 3xTF32 tensor-core MMAs with fp32 parity, mlp_tile.cuh
// -> gate applied to the node-linear output in the epilogue) and ONE backward kernel (SiLU', two transposed tile GEMMs, LayerNorm
// backward, norm backward, plus the gradient that reached X through the node linears added in the same pass).  The element-wise
// gate backward (tmd_gate_bwd_preact, tensor_ops.cu) runs in between because the node-linear backward needs its output.
// Geometry gradients only (frozen weights); training keeps the autograd-composed torch path.
#include "common.cuh"
#include "mlp_tile.cuh"

namespace {

using mlp::TA; using mlp::HT; using mlp::MAXNT;

struct CI { const float* I; const float* A; const float* S; };
struct MI { float* I; float* A; float* S; };

__device__ __forceinline__ void ld_c(const CI& t, long long n, int F, int h, float* c) {
  c[0] = t.I[n * F + h];
#pragma unroll
  for (int k = 0; k < 3; ++k) c[1 + k] = t.A[(n * 3 + k) * F + h];
#pragma unroll
  for (int k = 0; k < 5; ++k) c[4 + k] = t.S[(n * 5 + k) * F + h];
}

struct MlpW {
  const float* ln_w; const float* ln_b;   // [F]      init_norm
  const float* W1T; const float* b1;      // [F][2F], [2F]    linears_scalar[0] (transposed)
  const float* W2T; const float* b2;      // [2F][3F], [3F]   linears_scalar[1] (transposed)
  const float* W1; const float* W2;       // [2F][F], [3F][2F] untransposed (backward)
  float eps;
};

// shared memory: xs [TA][F+4] | h1s [TA][2F+4] | wbuf [3][KC][F+8]
__global__ void __launch_bounds__(HT, 2) embed_mlp_fwd_kernel(CI X, CI C, int N, int F, MlpW P, MI out, float* __restrict__ mu_rstd,
                                                           float* __restrict__ z1, float* __restrict__ z2) {
  extern __shared__ __align__(16) float sm[];
  const int ldx = F + 4, ld1 = 2 * F + 4;
  float* xs = sm; float* h1s = xs + TA * ldx; float* wbuf = h1s + TA * ld1;
  const int tid = threadIdx.x, lane = tid & 31, wy = tid >> 5;
  const long long a0 = (long long)blockIdx.x * TA;
  // 1. |X|^2 per channel
  for (int idx = tid; idx < TA * F; idx += HT) {
    const int a = idx / F, h = idx - a * F;
    float o = 0.f;
    if (a0 + a < N) { float c[9]; ld_c(X, a0 + a, F, h, c); o = tn::norm2_compact(c); }
    xs[a * ldx + h] = o;
  }
  __syncthreads();
  // 2. LayerNorm over F (two-pass, biased variance), in place; warp w owns atoms 4w .. 4w+3
  for (int q = 0; q < 4; ++q) {
    const int a = 4 * wy + q;
    float* row = xs + a * ldx;
    float s = 0.f;
    for (int k = lane; k < F; k += 32) s += row[k];
    const float mu = warp_sum(s) / (float)F;
    float v = 0.f;
    for (int k = lane; k < F; k += 32) { const float d = row[k] - mu; v = fmaf(d, d, v); }
    const float rstd = rsqrtf(warp_sum(v) / (float)F + P.eps);
    for (int k = lane; k < F; k += 32) row[k] = (row[k] - mu) * rstd * __ldg(P.ln_w + k) + __ldg(P.ln_b + k);
    if (lane == 0 && a0 + a < N) { mu_rstd[2 * (a0 + a)] = mu; mu_rstd[2 * (a0 + a) + 1] = rstd; }
  }
  __syncthreads();
  float acc[2][MAXNT][4];
  // 3. h1 = SiLU(W1 xn + b1): 2F columns in two blocks of F
  for (int blk = 0; blk < 2; ++blk) {
    mlp::zero_acc(acc);
    mlp::tile_gemm_tc(xs, ldx, F, P.W1T + blk * F, 2 * F, F, acc, wbuf, wy, lane, tid);
    mlp::for_each_acc(acc, F, wy, lane, [&](int row, int col, float& v) {
      const int cc = blk * F + col;
      const float zz = v + __ldg(P.b1 + cc);
      if (a0 + row < N) z1[(a0 + row) * 2 * F + cc] = zz;
      h1s[row * ld1 + cc] = tn::silu(zz);
    });
  }
  __syncthreads();
  // 4. g = SiLU(W2 h1 + b2): three blocks of F = the gates of I, A, S; applied to the node-linear output on the way out
  for (int blk = 0; blk < 3; ++blk) {
    mlp::zero_acc(acc);
    mlp::tile_gemm_tc(h1s, ld1, 2 * F, P.W2T + blk * F, 3 * F, F, acc, wbuf, wy, lane, tid);
    mlp::for_each_acc(acc, F, wy, lane, [&](int row, int col, float& v) {
      const long long n = a0 + row;
      if (n < N) {
        const float zz = v + __ldg(P.b2 + blk * F + col);
        z2[n * 3 * F + blk * F + col] = zz;
        const float gg = tn::silu(zz);
        if (blk == 0) {
          out.I[n * F + col] = C.I[n * F + col] * gg;
        } else if (blk == 1) {
#pragma unroll
          for (int k = 0; k < 3; ++k) out.A[(n * 3 + k) * F + col] = C.A[(n * 3 + k) * F + col] * gg;
        } else {
#pragma unroll
          for (int k = 0; k < 5; ++k) out.S[(n * 5 + k) * F + col] = C.S[(n * 5 + k) * F + col] * gg;
        }
      }
    });
  }
}

// dg [N][3F] = gradient w.r.t. the gates g (from tmd_gate_bwd_preact).  shared memory: d2s [TA][3F+4] | d1s [TA][2F+4] | wbuf;
// dxs [TA][F+4] aliases d2s (dead after the first GEMM).  add (may be null): gradient that reached X through the node linears.
__global__ void __launch_bounds__(HT, 2) embed_mlp_bwd_kernel(CI X, int N, int F, MlpW P, const float* __restrict__ dg,
                                                           const float* __restrict__ mu_rstd, const float* __restrict__ z1,
                                                           const float* __restrict__ z2, CI add, MI gX) {
  extern __shared__ __align__(16) float sm[];
  const int ld2 = 3 * F + 4, ld1 = 2 * F + 4, ldx = F + 4;
  float* d2s = sm; float* d1s = d2s + TA * ld2; float* wbuf = d1s + TA * ld1; float* dxs = d2s;
  const int tid = threadIdx.x, lane = tid & 31, wy = tid >> 5;
  const long long a0 = (long long)blockIdx.x * TA;
  // dz2 = dg * SiLU'(z2)
  for (int idx = tid; idx < TA * 3 * F; idx += HT) {
    const int a = idx / (3 * F), c = idx - a * 3 * F;
    float v = 0.f;
    if (a0 + a < N) v = dg[(a0 + a) * 3 * F + c] * tn::silu_grad(z2[(a0 + a) * 3 * F + c]);
    d2s[a * ld2 + c] = v;
  }
  __syncthreads();
  float acc[2][MAXNT][4];
  // dz1 = (dz2 . W2) * SiLU'(z1): 2F columns in two blocks
  for (int blk = 0; blk < 2; ++blk) {
    mlp::zero_acc(acc);
    mlp::tile_gemm_tc(d2s, ld2, 3 * F, P.W2 + blk * F, 2 * F, F, acc, wbuf, wy, lane, tid);
    mlp::for_each_acc(acc, F, wy, lane, [&](int row, int col, float& v) {
      const int cc = blk * F + col;
      const float zz = (a0 + row < N) ? z1[(a0 + row) * 2 * F + cc] : 0.f;
      d1s[row * ld1 + cc] = v * tn::silu_grad(zz);
    });
  }
  __syncthreads();   // d2s is dead from here on: dxs takes its place
  // dxn = dz1 . W1; g = dxn * gamma
  mlp::zero_acc(acc);
  mlp::tile_gemm_tc(d1s, ld1, 2 * F, P.W1, F, F, acc, wbuf, wy, lane, tid);
  mlp::for_each_acc(acc, F, wy, lane, [&](int row, int col, float& v) { dxs[row * ldx + col] = v * __ldg(P.ln_w + col); });
  __syncthreads();
  // LayerNorm backward per atom, then the norm backward: gX_c = add_c + d|X|^2 * d norm2 / d X_c
  for (int q = 0; q < 4; ++q) {
    const int a = 4 * wy + q;
    if (a0 + a >= N) continue;   // warp-uniform
    const long long n = a0 + a;
    const float mu = mu_rstd[2 * n], rstd = mu_rstd[2 * n + 1];
    const float* g = dxs + a * ldx;
    float s1 = 0.f, s2 = 0.f;
    for (int h = lane; h < F; h += 32) {
      float c[9]; ld_c(X, n, F, h, c);
      const float xh = (tn::norm2_compact(c) - mu) * rstd;
      s1 += g[h]; s2 = fmaf(g[h], xh, s2);
    }
    s1 = warp_sum(s1) / (float)F; s2 = warp_sum(s2) / (float)F;
    for (int h = lane; h < F; h += 32) {
      float c[9], gc[9];
      ld_c(X, n, F, h, c);
      const float xh = (tn::norm2_compact(c) - mu) * rstd;
      const float dn = rstd * (g[h] - s1 - xh * s2);
      tn::norm2_compact_grad(c, gc);
      if (add.I) {
        float ad[9]; ld_c(add, n, F, h, ad);
#pragma unroll
        for (int k = 0; k < 9; ++k) gc[k] = fmaf(gc[k], dn, ad[k]);
      } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) gc[k] *= dn;
      }
      gX.I[n * F + h] = gc[0];
#pragma unroll
      for (int k = 0; k < 3; ++k) gX.A[(n * 3 + k) * F + h] = gc[1 + k];
#pragma unroll
      for (int k = 0; k < 5; ++k) gX.S[(n * 5 + k) * F + h] = gc[4 + k];
    }
  }
}

}  // namespace

extern "C" {

// X = (I, A, S): un-normalised embedding (input of the norm); C = (cI, cA, cS): node-linear output that is gated.
// W1T [F][2F], W2T [2F][3F] transposed nn.Linear weights.  Saves (mu, rstd) [N,2], z1 [N,2F], z2 [N,3F].  F % 32 == 0, F <= 256.
int tmd_embed_mlp_fwd(const float* I, const float* A, const float* S, const float* cI, const float* cA, const float* cS, int N, int F,
                      const float* ln_w, const float* ln_b, double eps, const float* W1T, const float* b1, const float* W2T,
                      const float* b2, float* oI, float* oA, float* oS, float* mu_rstd, float* z1, float* z2, void* stream) {
  TMD_REQUIRE(F > 0 && F % 32 == 0 && F <= 256, "embed_mlp_fwd: F must be a multiple of 32, <= 256");
  TMD_REQUIRE(I && A && S && cI && cA && cS && ln_w && ln_b && W1T && b1 && W2T && b2 && oI && oA && oS && mu_rstd && z1 && z2,
              "embed_mlp_fwd: null pointer");
  if (N == 0) return TMD_OK;
  MlpW P{ln_w, ln_b, W1T, b1, W2T, b2, nullptr, nullptr, (float)eps};
  const size_t smem = ((size_t)TA * ((F + 4) + (2 * F + 4)) + mlp::wbuf_floats(F)) * sizeof(float);
  cudaFuncSetAttribute(embed_mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  embed_mlp_fwd_kernel<<<(N + TA - 1) / TA, HT, smem, (cudaStream_t)stream>>>(CI{I, A, S}, CI{cI, cA, cS}, N, F, P, MI{oI, oA, oS}, mu_rstd, z1, z2);
  return tmd_check_launch("embed_mlp_fwd");
}

// gX = add + (gradient of the gates dg [N,3F] taken back through SiLU, W2, SiLU, W1, LayerNorm and |X|^2).  W1 [2F][F], W2 [3F][2F]
// untransposed nn.Linear weights; add triple may be NULL (all three).
int tmd_embed_mlp_bwd(const float* I, const float* A, const float* S, int N, int F, const float* ln_w, double eps, const float* W1,
                      const float* W2, const float* dg, const float* mu_rstd, const float* z1, const float* z2, const float* addI,
                      const float* addA, const float* addS, float* gI, float* gA, float* gS, void* stream) {
  TMD_REQUIRE(F > 0 && F % 32 == 0 && F <= 256, "embed_mlp_bwd: F must be a multiple of 32, <= 256");
  TMD_REQUIRE(I && A && S && ln_w && W1 && W2 && dg && mu_rstd && z1 && z2 && gI && gA && gS, "embed_mlp_bwd: null pointer");
  TMD_REQUIRE((addI && addA && addS) || (!addI && !addA && !addS), "embed_mlp_bwd: the add triple is all-or-nothing");
  if (N == 0) return TMD_OK;
  MlpW P{ln_w, nullptr, nullptr, nullptr, nullptr, nullptr, W1, W2, (float)eps};
  const size_t smem = ((size_t)TA * ((3 * F + 4) + (2 * F + 4)) + mlp::wbuf_floats(F)) * sizeof(float);
  TMD_REQUIRE(smem <= 227 * 1024, "embed_mlp_bwd: F too large for the tile buffers");
  cudaFuncSetAttribute(embed_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  embed_mlp_bwd_kernel<<<(N + TA - 1) / TA, HT, smem, (cudaStream_t)stream>>>(CI{I, A, S}, N, F, P, dg, mu_rstd, z1, z2, CI{addI, addA, addS},
                                                                             MI{gI, gA, gS});
  return tmd_check_launch("embed_mlp_bwd");
}

}  // extern "C"
