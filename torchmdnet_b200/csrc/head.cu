// a10 + a11: TensorNet tail + Scalar head + per-molecule energy reduction, and their backward (the seed of the force pass).
//
// Reference chain (all separate launches there): tensor_norm3 -> LayerNorm(3F) -> Linear(3F->F) -> SiLU
// (models/tensornet.py:394-402) -> MLP Linear(F->F/2) -> SiLU -> Linear(F/2->1) (models/output_modules.py:79-117,
// models/utils.py:531-580) -> * std -> + atomref[z] (priors/atomref.py:66) -> scatter-sum over `batch`
// (output_modules.py:43-73, models/utils.py:671-701) -> + mean (models/model.py:590-610); in round 1 this was one own kernel
// (norm3c) + LayerNorm + 3 cuBLAS SIMT SGEMMs + SiLU/fill/add/scatter launches of torch and their autograd mirrors.
//
// Here: ONE forward kernel per 32-atom tile (norm3 -> LN -> two GEMM layers on the tile, weights streamed k-major through shared
// memory, 3xTF32 warp-level tensor-core MMAs with fp32 parity: mlp_tile.cuh -> per-atom energy), one deterministic segmented sum,
// ONE backward kernel (dE/de_i = gy[batch] * std seeded by the caller's grad, back through the MLP with the saved
// pre-activations, LayerNorm backward, norm3 backward -> dI, dA, dS).  The work is ~2.4 GFLOP per C4 frame.
// Weight gradients are not produced (inference / frozen weights); training keeps the autograd-composed torch path.
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

struct HeadW {
  const float* ln_w; const float* ln_b;   // [3F]   representation_model.out_norm
  const float* W1T; const float* b1;      // [3F][F], [F]    representation_model.linear (transposed)
  const float* W2T; const float* b2;      // [F][H], [H]     output_network.layers.0 (transposed)
  const float* w3; const float* b3;       // [H], [1]        output_network.layers.2
  const float* W1; const float* W2;       // [F][3F], [H][F] untransposed copies (backward)
  const float* std_;                      // [1]  model.std
  float eps;
};

// shared memory: xs [TA][3F+4] | h1s [TA][F+4] | h2s [TA][H+4] | wbuf [3][KC][max(F,H)+8] (weight staging, mlp_tile.cuh)
__global__ void __launch_bounds__(HT, 2) head_fwd_kernel(CI X, int N, int F, int H, HeadW P, const int64_t* __restrict__ z,
                                                      const float* __restrict__ atomref, int n_ref, float* __restrict__ e_out,
                                                      float* __restrict__ mu_rstd, float* __restrict__ z1, float* __restrict__ z2) {
  extern __shared__ __align__(16) float sm[];
  const int ldx = 3 * F + 4, ld1 = F + 4, ld2 = H + 4;
  float* xs = sm; float* h1s = xs + TA * ldx; float* h2s = h1s + TA * ld1; float* wbuf = h2s + TA * ld2;
  const int tid = threadIdx.x, lane = tid & 31, wy = tid >> 5;
  const long long a0 = (long long)blockIdx.x * TA;
  // 1. norm3 of the node tensors
  for (int idx = tid; idx < TA * F; idx += HT) {
    const int a = idx / F, h = idx - a * F;
    float o[3] = {0.f, 0.f, 0.f};
    if (a0 + a < N) { float c[9]; ld_c(X, a0 + a, F, h, c); tn::norm3_compact(c, o); }
    xs[a * ldx + h] = o[0]; xs[a * ldx + F + h] = o[1]; xs[a * ldx + 2 * F + h] = o[2];
  }
  __syncthreads();
  // 2. LayerNorm over 3F (two-pass, biased variance), in place; warp w owns atoms 4w .. 4w+3
  for (int q = 0; q < 4; ++q) {
    const int a = 4 * wy + q;
    float* row = xs + a * ldx;
    float s = 0.f;
    for (int k = lane; k < 3 * F; k += 32) s += row[k];
    const float mu = warp_sum(s) / (float)(3 * F);
    float v = 0.f;
    for (int k = lane; k < 3 * F; k += 32) { const float d = row[k] - mu; v = fmaf(d, d, v); }
    const float rstd = rsqrtf(warp_sum(v) / (float)(3 * F) + P.eps);
    for (int k = lane; k < 3 * F; k += 32) row[k] = (row[k] - mu) * rstd * __ldg(P.ln_w + k) + __ldg(P.ln_b + k);
    if (lane == 0 && a0 + a < N) { mu_rstd[2 * (a0 + a)] = mu; mu_rstd[2 * (a0 + a) + 1] = rstd; }
  }
  __syncthreads();
  float acc[2][MAXNT][4];
  // 3. h1 = SiLU(W1 xn + b1)
  mlp::zero_acc(acc);
  mlp::tile_gemm_tc(xs, ldx, 3 * F, P.W1T, F, F, acc, wbuf, wy, lane, tid);
  mlp::for_each_acc(acc, F, wy, lane, [&](int row, int col, float& v) {
    const float zz = v + __ldg(P.b1 + col);
    if (a0 + row < N) z1[(a0 + row) * F + col] = zz;
    h1s[row * ld1 + col] = tn::silu(zz);
  });
  __syncthreads();
  // 4. h2 = SiLU(W2 h1 + b2)
  mlp::zero_acc(acc);
  mlp::tile_gemm_tc(h1s, ld1, F, P.W2T, H, H, acc, wbuf, wy, lane, tid);
  mlp::for_each_acc(acc, H, wy, lane, [&](int row, int col, float& v) {
    const float zz = v + __ldg(P.b2 + col);
    if (a0 + row < N) z2[(a0 + row) * H + col] = zz;
    h2s[row * ld2 + col] = tn::silu(zz);
  });
  __syncthreads();
  // 5. e_i = (w3 . h2 + b3) * std + atomref[z_i]
  for (int q = 0; q < 4; ++q) {
    const int a = 4 * wy + q;
    float s = 0.f;
    for (int k = lane; k < H; k += 32) s = fmaf(h2s[a * ld2 + k], __ldg(P.w3 + k), s);
    s = warp_sum(s);
    if (lane == 0 && a0 + a < N) {
      float e = (s + __ldg(P.b3)) * __ldg(P.std_);
      if (atomref) { const int64_t zz = z[a0 + a]; if (zz >= 0 && zz < n_ref) e += __ldg(atomref + zz); }
      e_out[a0 + a] = e;
    }
  }
}

// shared memory: dxs [TA][3F+4] | d1s [TA][F+4] | d2s [TA][H+4] | wbuf.  The pre-LayerNorm norms (xhat of the LayerNorm backward)
// are recomputed from the node tensors instead of being kept in a second [TA][3F] tile: two CTAs share an SM.
__global__ void __launch_bounds__(HT, 2) head_bwd_kernel(CI X, int N, int F, int H, HeadW P, const float* __restrict__ gy,
                                                      const int64_t* __restrict__ batch, int M, const float* __restrict__ mu_rstd,
                                                      const float* __restrict__ z1, const float* __restrict__ z2, MI dX) {
  extern __shared__ __align__(16) float sm[];
  const int ldx = 3 * F + 4, ld1 = F + 4, ld2 = H + 4;
  float* dxs = sm; float* d1s = dxs + TA * ldx; float* d2s = d1s + TA * ld1; float* wbuf = d2s + TA * ld2;
  const int tid = threadIdx.x, lane = tid & 31, wy = tid >> 5;
  const long long a0 = (long long)blockIdx.x * TA;
  // dz2 = de * w3 * SiLU'(z2)
  const float stdv = __ldg(P.std_);
  for (int idx = tid; idx < TA * H; idx += HT) {
    const int a = idx / H, c = idx - a * H;
    float v = 0.f;
    if (a0 + a < N) {
      const int64_t m = batch[a0 + a];
      const float de = (m >= 0 && m < M) ? __ldg(gy + m) * stdv : 0.f;
      v = de * __ldg(P.w3 + c) * tn::silu_grad(z2[(a0 + a) * H + c]);
    }
    d2s[a * ld2 + c] = v;
  }
  __syncthreads();
  float acc[2][MAXNT][4];
  // dz1 = (dz2 . W2) * SiLU'(z1)
  mlp::zero_acc(acc);
  mlp::tile_gemm_tc(d2s, ld2, H, P.W2, F, F, acc, wbuf, wy, lane, tid);
  mlp::for_each_acc(acc, F, wy, lane, [&](int row, int col, float& v) {
    const float zz = (a0 + row < N) ? z1[(a0 + row) * F + col] : 0.f;
    d1s[row * ld1 + col] = v * tn::silu_grad(zz);
  });
  __syncthreads();
  // dxn = dz1 . W1, in three column blocks of F; g = dxn * gamma
  for (int blk = 0; blk < 3; ++blk) {
    mlp::zero_acc(acc);
    mlp::tile_gemm_tc(d1s, ld1, F, P.W1 + blk * F, 3 * F, F, acc, wbuf, wy, lane, tid);
    mlp::for_each_acc(acc, F, wy, lane, [&](int row, int col, float& v) {
      dxs[row * ldx + blk * F + col] = v * __ldg(P.ln_w + blk * F + col);
    });
  }
  __syncthreads();
  // LayerNorm backward per atom, dx = rstd (g - mean(g) - xhat mean(g xhat)), fused with the norm3 backward -> dI, dA, dS.
  // The node row is read twice (sums, then gradients); the second read hits L1.
  for (int q = 0; q < 4; ++q) {
    const int a = 4 * wy + q;
    if (a0 + a >= N) continue;   // warp-uniform
    const long long n = a0 + a;
    const float mu = mu_rstd[2 * n], rstd = mu_rstd[2 * n + 1];
    const float* g = dxs + a * ldx;
    float s1 = 0.f, s2 = 0.f;
    for (int h = lane; h < F; h += 32) {
      float c[9], o[3];
      ld_c(X, n, F, h, c); tn::norm3_compact(c, o);
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const float gv = g[t * F + h];
        s1 += gv; s2 = fmaf(gv, (o[t] - mu) * rstd, s2);
      }
    }
    s1 = warp_sum(s1) / (float)(3 * F); s2 = warp_sum(s2) / (float)(3 * F);
    for (int h = lane; h < F; h += 32) {
      float c[9], o[3], dn[3], dc[9];
      ld_c(X, n, F, h, c); tn::norm3_compact(c, o);
#pragma unroll
      for (int t = 0; t < 3; ++t) dn[t] = rstd * (g[t * F + h] - s1 - (o[t] - mu) * rstd * s2);
      tn::norm3_compact_bwd(c, dn, dc);
      dX.I[n * F + h] = dc[0];
#pragma unroll
      for (int k = 0; k < 3; ++k) dX.A[(n * 3 + k) * F + h] = dc[1 + k];
#pragma unroll
      for (int k = 0; k < 5; ++k) dX.S[(n * 5 + k) * F + h] = dc[4 + k];
    }
  }
}

// y[m] = mean + sum_{i: batch[i] == m} e[i], deterministic: one CTA per molecule, fixed summation tree.  `batch` sorted
// (how every caller builds it): the segment is found by binary search; otherwise (flag raised by the check kernel) the
// CTA scans all atoms.
__global__ void batch_sorted_kernel(const int64_t* __restrict__ batch, int N, int* __restrict__ unsorted) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i + 1 < N; i += gridDim.x * blockDim.x)
    if (batch[i] > batch[i + 1]) atomicOr(unsorted, 1);
}
__global__ void __launch_bounds__(256) segment_sum_kernel(const float* __restrict__ e, const int64_t* __restrict__ batch, int N, int M,
                                                          const float* __restrict__ mean, const int* __restrict__ unsorted,
                                                          float* __restrict__ y) {
  __shared__ float part[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int m = blockIdx.x; m < M; m += gridDim.x) {
    int lo = 0, hi = N;
    const bool scan_all = *unsorted != 0;
    if (!scan_all) {
      int a = 0, b = N;
      while (a < b) { const int mid = (a + b) >> 1; if (batch[mid] < m) a = mid + 1; else b = mid; }
      lo = a; b = N;
      while (a < b) { const int mid = (a + b) >> 1; if (batch[mid] <= m) a = mid + 1; else b = mid; }
      hi = a;
    }
    float s = 0.f;
    for (int i = lo + threadIdx.x; i < hi; i += 256) if (!scan_all || batch[i] == m) s += e[i];
    s = warp_sum(s);
    __syncthreads();
    if (lane == 0) part[w] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int k = 0; k < 8; ++k) t += part[k];
      y[m] = t + (mean ? *mean : 0.f);
    }
  }
}

}  // namespace

extern "C" {

// e[N] = per-atom energies; saves (mu, rstd) [N,2], z1 [N,F], z2 [N,H] for tmd_head_bwd.  F % 32 == 0, F <= 256, H % 32 == 0,
// H <= 256.  atomref may be NULL (no Atomref prior); z int64 atomic numbers (only read with atomref).
int tmd_head_fwd(const float* I, const float* A, const float* S, int N, int F, int H, const float* ln_w, const float* ln_b,
                 double eps, const float* W1T, const float* b1, const float* W2T, const float* b2, const float* w3,
                 const float* b3, const float* std_, const int64_t* z, const float* atomref, int n_ref, float* e,
                 float* mu_rstd, float* z1, float* z2, void* stream) {
  TMD_REQUIRE(F > 0 && F % 32 == 0 && F <= 256 && H > 0 && H % 32 == 0 && H <= 256, "head_fwd: F, H must be multiples of 32, <= 256");
  TMD_REQUIRE(I && A && S && ln_w && ln_b && W1T && b1 && W2T && b2 && w3 && b3 && std_ && e && mu_rstd && z1 && z2, "head_fwd: null pointer");
  TMD_REQUIRE(!atomref || z, "head_fwd: atomref without z");
  if (N == 0) return TMD_OK;
  HeadW P{ln_w, ln_b, W1T, b1, W2T, b2, w3, b3, nullptr, nullptr, std_, (float)eps};
  const size_t smem = ((size_t)TA * ((3 * F + 4) + (F + 4) + (H + 4)) + mlp::wbuf_floats(F > H ? F : H)) * sizeof(float);
  cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  head_fwd_kernel<<<(N + TA - 1) / TA, HT, smem, (cudaStream_t)stream>>>(CI{I, A, S}, N, F, H, P, z, atomref, n_ref, e, mu_rstd, z1, z2);
  return tmd_check_launch("head_fwd");
}

// y[M] = mean + per-molecule sums of e (deterministic).  flag: one int of scratch.
int tmd_segment_sum(const float* e, const int64_t* batch, int N, int M, const float* mean, int* flag, float* y, void* stream) {
  TMD_REQUIRE(e && batch && flag && y && M >= 0, "segment_sum: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(flag, 0, sizeof(int), st);
  if (M == 0) return tmd_check_launch("segment_sum(empty)");
  if (N > 1) batch_sorted_kernel<<<tmd_blocks_for(N, 256, 148 * 4), 256, 0, st>>>(batch, N, flag);
  segment_sum_kernel<<<M < 148 * 8 ? M : 148 * 8, 256, 0, st>>>(e, batch, N, M, mean, flag, y);
  return tmd_check_launch("segment_sum");
}

// dI, dA, dS of y = segment_sum(head(I, A, S)) for the upstream gradient gy [M] (forces: ones).
int tmd_head_bwd(const float* I, const float* A, const float* S, int N, int F, int H, const float* ln_w, double eps,
                 const float* W1, const float* W2, const float* w3, const float* std_, const float* gy, const int64_t* batch,
                 int M, const float* mu_rstd, const float* z1, const float* z2, float* dI, float* dA, float* dS, void* stream) {
  TMD_REQUIRE(F > 0 && F % 32 == 0 && F <= 256 && H > 0 && H % 32 == 0 && H <= 256, "head_bwd: F, H must be multiples of 32, <= 256");
  TMD_REQUIRE(I && A && S && ln_w && W1 && W2 && w3 && std_ && gy && batch && mu_rstd && z1 && z2 && dI && dA && dS, "head_bwd: null pointer");
  if (N == 0) return TMD_OK;
  HeadW P{ln_w, nullptr, nullptr, nullptr, nullptr, nullptr, w3, nullptr, W1, W2, std_, (float)eps};
  const size_t smem = ((size_t)TA * ((3 * F + 4) + (F + 4) + (H + 4)) + mlp::wbuf_floats(F > H ? F : H)) * sizeof(float);
  TMD_REQUIRE(smem <= 227 * 1024, "head_bwd: F too large for the tile buffers");
  cudaFuncSetAttribute(head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  head_bwd_kernel<<<(N + TA - 1) / TA, HT, smem, (cudaStream_t)stream>>>(CI{I, A, S}, N, F, H, P, gy, batch, M, mu_rstd, z1, z2, MI{dI, dA, dS});
  return tmd_check_launch("head_bwd");
}

}  // extern "C"
