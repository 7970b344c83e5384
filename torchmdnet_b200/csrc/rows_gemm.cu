// Per-atom dense layers of the hot path as weight-tiled GEMMs, and the row-wise stages around them.
//
// Rows a10 / a11 (TensorNet tail norm3 -> LayerNorm -> Linear -> SiLU and the Scalar head MLP; reference models/tensornet.py:394-402,
// models/output_modules.py:79-117) and the scalar gate of the embedding (row a5: LayerNorm(|X|^2) -> Linear -> SiLU -> Linear ->
// SiLU, models/tensornet.py:588-606) are chains of [N, K] x [K, Nout] products with N = atoms, K, Nout <= 768.  The reference runs
// them as cuBLAS SGEMMs with separate bias / activation / LayerNorm launches and their autograd mirrors.
//
// tmd_rows_gemm: C = epilogue(A . W), A [M, K] row-major fp32, W [K, N] k-major fp32, fp32-parity through 3xTF32 warp-level
// tensor-core MMAs (operands split on the fly into two TF32 terms, lo.hi + hi.lo + hi.hi accumulated per k-step into a zeroed
// temporary that is added with an fp32 FADD: the tensor core adds with truncation, and chaining every MMA through one accumulator
// biased the result by ~1e-5, see mlp_tile.cuh).  64 x 64 (or 64 x 32) output tiles, both operands streamed through a four-stage
// cp.async pipeline; unlike the per-atom-tile kernels this form re-uses every staged weight chunk for 64 rows and keeps enough
// independent CTAs per SM to cover the L2 latency of the next chunk (the per-tile kernels in head.cu / atom_mlp.cu re-stream the
// whole weight matrix for every 32 atoms and were latency-bound: 334 us for the embedding MLP, whose arithmetic takes ~10 us).
// Epilogues: bias + SiLU (optionally saving the pre-activation), x SiLU'(aux) and x per-column scale: what the forward and the
// hand-written backward of the chains need, so no element-wise launch sits between two GEMMs.
//
// tmd_norm_ln_fwd / tmd_norm_ln_bwd: per-atom norms of the node tensors (|X|^2 per channel, or the three irrep norms) followed by
// LayerNorm, and their backward fused with the gradient that reached the node tensors along another path.
// tmd_head_energy / tmd_head_seed: the last Linear(H -> 1) * std + atomref of the Scalar head, and the seed of its backward.
#include <stdlib.h>
#include "common.cuh"
#include "mlp_tile.cuh"

namespace {

constexpr int BK = 16, GST = 4, GT = 256;               // k per stage, stages, threads; rows per tile BM = 64 or 128 (template)
constexpr int ALD = BK + 4;                            // A tile row stride (20: fragment reads hit 32 banks)

enum Epi { EPI_BIAS_SILU = 0, EPI_MUL_SILU_GRAD = 1, EPI_COLSCALE = 2, EPI_LINEAR = 3, EPI_ADD = 4 };

template <int BM, int BN>
__device__ __forceinline__ void stage_tiles(float* As, float* Bs, const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                                            int m0, int n0, int k0, int M, int tid) {
#pragma unroll
  for (int i = tid; i < BM * 4; i += GT) {   // A: BM rows x 16 columns = BM * 4 float4
    const int r = i >> 2, c4 = i & 3;
    const int row = min(m0 + r, M - 1);                // rows past M read the last row (never stored)
    mlp::cp_async16(As + (size_t)r * ALD + 4 * c4, A + (size_t)row * lda + k0 + 4 * c4);
  }
  {   // B: 16 rows x BN columns
    constexpr int C4 = BN / 4;
    for (int i = tid; i < BK * C4; i += GT) {
      const int r = i / C4, c4 = i - r * C4;
      mlp::cp_async16(Bs + (size_t)r * (BN + 8) + 4 * c4, W + (size_t)(k0 + r) * ldw + n0 + 4 * c4);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// BM = 64: 8 warps = 4 (rows) x 2 (columns), warp tile 16 x BN/2.  BM = 128: warp tile 32 x BN/2 (two m16 tiles per warp share
// every B fragment and its split: fewer shared-memory reads and conversions per MMA); used when the grid still fills the SMs.
template <int BM, int BN, int EPI>
__global__ void __launch_bounds__(GT) rows_gemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw, int M, int K,
                                                       int N, const float* __restrict__ vec, const float* __restrict__ aux, int ldaux,
                                                       float* __restrict__ zout, float* __restrict__ out, int ldo) {
  extern __shared__ __align__(16) float sm[];
  constexpr int NT = BN / 16;                          // n8 tiles per warp
  constexpr int MT = BM / 64;                          // m16 tiles per warp
  constexpr int ASZ = BM * ALD, BSZ = BK * (BN + 8);
  float* As = sm; float* Bs = sm + GST * ASZ;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp >> 1, wn = warp & 1;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int nchunks = K / BK;
  float acc[MT][NT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[mt][nt][r] = 0.f;
#pragma unroll
  for (int s = 0; s < GST - 1; ++s) {
    if (s < nchunks) stage_tiles<BM, BN>(As + s * ASZ, Bs + s * BSZ, A, lda, W, ldw, m0, n0, s * BK, M, tid);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int c = 0; c < nchunks; ++c) {
    asm volatile("cp.async.wait_group %0;" ::"n"(GST - 2) : "memory");
    __syncthreads();
    if (c + GST - 1 < nchunks) {
      const int s = (c + GST - 1) % GST;
      stage_tiles<BM, BN>(As + s * ASZ, Bs + s * BSZ, A, lda, W, ldw, m0, n0, (c + GST - 1) * BK, M, tid);
    } else {
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const float* as = As + (c % GST) * ASZ;
    const float* bs = Bs + (c % GST) * BSZ;
#pragma unroll
    for (int ks = 0; ks < BK; ks += 8) {
      uint32_t ah[MT][4], al[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const float* a = as + (size_t)(16 * (MT * wm + mt) + g) * ALD + ks + t;
        mlp::split_tf32(a[0], ah[mt][0], al[mt][0]);
        mlp::split_tf32(a[8 * ALD], ah[mt][1], al[mt][1]);
        mlp::split_tf32(a[4], ah[mt][2], al[mt][2]);
        mlp::split_tf32(a[8 * ALD + 4], ah[mt][3], al[mt][3]);
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float* b = bs + (size_t)(ks + t) * (BN + 8) + wn * (BN / 2) + 8 * nt + g;
        uint32_t bh0, bl0, bh1, bl1;
        mlp::split_tf32(b[0], bh0, bl0);
        mlp::split_tf32(b[4 * (BN + 8)], bh1, bl1);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          float tmp[4] = {0.f, 0.f, 0.f, 0.f};
          mlp::mma_tf32(tmp, al[mt], bh0, bh1);
          mlp::mma_tf32(tmp, ah[mt], bl0, bl1);
          mlp::mma_tf32(tmp, ah[mt], bh0, bh1);
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[mt][nt][r] += tmp[r];
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // epilogue: rows 16 (MT wm + mt) + g (+ 8), columns n0 + wn BN/2 + 8 nt + 2 t (+ 1): one 8-byte store per (row, tile)
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int col = n0 + wn * (BN / 2) + 8 * nt + 2 * t;
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int row = m0 + 16 * (MT * wm + mt) + g + 8 * hr;
      if (row >= M) continue;
      float v0 = acc[mt][nt][2 * hr], v1 = acc[mt][nt][2 * hr + 1];
      if (EPI == EPI_BIAS_SILU) {
        v0 += __ldg(vec + col); v1 += __ldg(vec + col + 1);
        if (zout) *reinterpret_cast<float2*>(zout + (size_t)row * ldo + col) = make_float2(v0, v1);
        v0 = tn::silu(v0); v1 = tn::silu(v1);
      } else if (EPI == EPI_MUL_SILU_GRAD) {
        const float2 z = *reinterpret_cast<const float2*>(aux + (size_t)row * ldaux + col);
        v0 *= tn::silu_grad(z.x); v1 *= tn::silu_grad(z.y);
      } else if (EPI == EPI_COLSCALE) {
        v0 *= __ldg(vec + col); v1 *= __ldg(vec + col + 1);
      } else if (EPI == EPI_LINEAR) {
        if (vec) { v0 += __ldg(vec + col); v1 += __ldg(vec + col + 1); }
      } else {
        const float2 z = *reinterpret_cast<const float2*>(aux + (size_t)row * ldaux + col);
        v0 += z.x; v1 += z.y;
      }
      *reinterpret_cast<float2*>(out + (size_t)row * ldo + col) = make_float2(v0, v1);
    }
  }
}

struct CI { const float* I; const float* A; const float* S; };
struct MI { float* I; float* A; float* S; };
__device__ __forceinline__ void ld_c(const CI& t, long long n, int F, int h, float* c) {
  c[0] = t.I[n * F + h];
#pragma unroll
  for (int k = 0; k < 3; ++k) c[1 + k] = t.A[(n * 3 + k) * F + h];
#pragma unroll
  for (int k = 0; k < 5; ++k) c[4 + k] = t.S[(n * 5 + k) * F + h];
}

// one warp per atom.  MODE 3: x = (3 I^2, |A|^2, |S|^2) in R^{3F} (tensor_norm3);  MODE 2: x = |X|^2 per channel in R^F.
// xn = LayerNorm(x) (two-pass, biased variance); mu_rstd [N, 2] saved for the backward.  F <= 256.
template <int MODE>
__global__ void __launch_bounds__(256) norm_ln_fwd_kernel(CI X, int N, int F, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                          float eps, float* __restrict__ xn, float* __restrict__ mu_rstd) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  constexpr int NV = MODE == 3 ? 3 : 1;
  const int K = NV * F;
  for (long long n = blockIdx.x * wpb + (threadIdx.x >> 5); n < N; n += (long long)gridDim.x * wpb) {
    float x[8][NV];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int h = lane + 32 * j;
      if (h < F) {
        float c[9]; ld_c(X, n, F, h, c);
        if (MODE == 3) { float o[3]; tn::norm3_compact(c, o); for (int q = 0; q < NV; ++q) x[j][q] = o[q]; }
        else x[j][0] = tn::norm2_compact(c);
#pragma unroll
        for (int q = 0; q < NV; ++q) s += x[j][q];
      }
    }
    const float mu = warp_sum(s) / (float)K;
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (lane + 32 * j < F)
#pragma unroll
        for (int q = 0; q < NV; ++q) { const float d = x[j][q] - mu; v = fmaf(d, d, v); }
    const float rstd = rsqrtf(warp_sum(v) / (float)K + eps);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int h = lane + 32 * j;
      if (h < F)
#pragma unroll
        for (int q = 0; q < NV; ++q)
          xn[n * K + q * F + h] = (x[j][q] - mu) * rstd * __ldg(ln_w + q * F + h) + __ldg(ln_b + q * F + h);
    }
    if (lane == 0) { mu_rstd[2 * n] = mu; mu_rstd[2 * n + 1] = rstd; }
  }
}

// g [N, K] = d xn * gamma.  dx = rstd (g - mean(g) - xhat mean(g xhat)); then the norm backward; gX = add + ...
template <int MODE>
__global__ void __launch_bounds__(256) norm_ln_bwd_kernel(CI X, int N, int F, const float* __restrict__ g, const float* __restrict__ mu_rstd,
                                                          CI add, MI gX) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  constexpr int NV = MODE == 3 ? 3 : 1;
  const int K = NV * F;
  for (long long n = blockIdx.x * wpb + (threadIdx.x >> 5); n < N; n += (long long)gridDim.x * wpb) {
    const float mu = mu_rstd[2 * n], rstd = mu_rstd[2 * n + 1];
    float s1 = 0.f, s2 = 0.f;
    for (int h = lane; h < F; h += 32) {
      float c[9], o[3]; ld_c(X, n, F, h, c);
      if (MODE == 3) tn::norm3_compact(c, o); else o[0] = tn::norm2_compact(c);
#pragma unroll
      for (int q = 0; q < NV; ++q) { const float gv = g[n * K + q * F + h]; s1 += gv; s2 = fmaf(gv, (o[q] - mu) * rstd, s2); }
    }
    s1 = warp_sum(s1) / (float)K; s2 = warp_sum(s2) / (float)K;
    for (int h = lane; h < F; h += 32) {     // second pass: the node row comes from L1
      float c[9], o[3], dn[3], dc[9]; ld_c(X, n, F, h, c);
      if (MODE == 3) tn::norm3_compact(c, o); else o[0] = tn::norm2_compact(c);
#pragma unroll
      for (int q = 0; q < NV; ++q) dn[q] = rstd * (g[n * K + q * F + h] - s1 - (o[q] - mu) * rstd * s2);
      if (MODE == 3) tn::norm3_compact_bwd(c, dn, dc);
      else { tn::norm2_compact_grad(c, dc); for (int k = 0; k < 9; ++k) dc[k] *= dn[0]; }
      if (add.I) { float ad[9]; ld_c(add, n, F, h, ad); for (int k = 0; k < 9; ++k) dc[k] += ad[k]; }
      gX.I[n * F + h] = dc[0];
#pragma unroll
      for (int k = 0; k < 3; ++k) gX.A[(n * 3 + k) * F + h] = dc[1 + k];
#pragma unroll
      for (int k = 0; k < 5; ++k) gX.S[(n * 5 + k) * F + h] = dc[4 + k];
    }
  }
}

// e_i = (w3 . h2_i + b3) * std + atomref[z_i]      (warp per atom)
__global__ void __launch_bounds__(256) head_energy_kernel(const float* __restrict__ h2, int N, int H, const float* __restrict__ w3,
                                                          const float* __restrict__ b3, const float* __restrict__ std_, const int64_t* __restrict__ z,
                                                          const float* __restrict__ atomref, int n_ref, float* __restrict__ e) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (long long n = blockIdx.x * wpb + (threadIdx.x >> 5); n < N; n += (long long)gridDim.x * wpb) {
    float s = 0.f;
    for (int k = lane; k < H; k += 32) s = fmaf(h2[n * H + k], __ldg(w3 + k), s);
    s = warp_sum(s);
    if (lane == 0) {
      float v = (s + __ldg(b3)) * __ldg(std_);
      if (atomref) { const int64_t zz = z[n]; if (zz >= 0 && zz < n_ref) v += __ldg(atomref + zz); }
      e[n] = v;
    }
  }
}
// dz2[i, c] = gy[batch_i] * std * w3[c] * SiLU'(z2[i, c])
__global__ void head_seed_kernel(const float* __restrict__ gy, const int64_t* __restrict__ batch, int M, const float* __restrict__ std_,
                                 const float* __restrict__ w3, const float* __restrict__ z2, int N, int H, float* __restrict__ dz2) {
  const float stdv = __ldg(std_);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)N * H; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / H; const int c = (int)(i - n * H);
    const int64_t m = batch[n];
    const float de = (m >= 0 && m < M) ? __ldg(gy + m) * stdv : 0.f;
    dz2[i] = de * __ldg(w3 + c) * tn::silu_grad(z2[i]);
  }
}

// ---------------------------------------------------------------------------------------------- a15: GatedEquivariantBlock pieces
// (reference models/utils.py:583-655).  cat [N, 2H] = [x | vec1], vec1[n,h] = |P1[n,:,h]| over the 3 vector components (0 for an
// all-zero row; the reference masks those rows out of torch.norm so that the gradient stays finite).
__global__ void geb_cat_fwd_kernel(const float* __restrict__ x, const float* __restrict__ P1, int N, int H, float* __restrict__ cat) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)N * H; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / H; const int h = (int)(i - n * H);
    const float a = P1[(n * 3) * H + h], b = P1[(n * 3 + 1) * H + h], c = P1[(n * 3 + 2) * H + h];
    const float sq = a * a + b * b + c * c;
    cat[n * 2 * H + h] = x[i];
    cat[n * 2 * H + H + h] = sq > 0.f ? sqrtf(sq) : 0.f;
  }
}
// g_P1[n,c,h] = g_vec1[n,h] * P1[n,c,h] / vec1[n,h]  (0 where vec1 = 0); g_cat [N, 2H] holds g_vec1 in its second half
__global__ void geb_cat_bwd_kernel(const float* __restrict__ P1, const float* __restrict__ cat, const float* __restrict__ g_cat, int N, int H,
                                   float* __restrict__ g_P1) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)N * H; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / H; const int h = (int)(i - n * H);
    const float v1 = cat[n * 2 * H + H + h];
    const float s = v1 > 0.f ? g_cat[n * 2 * H + H + h] / v1 : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) g_P1[(n * 3 + c) * H + h] = s * P1[(n * 3 + c) * H + h];
  }
}
// hh [N, 2O] = [xo | gate]; V2 [N,3,O]: xo_out = act ? SiLU(xo) : xo, vo = gate * V2
__global__ void geb_out_fwd_kernel(const float* __restrict__ hh, const float* __restrict__ V2, int N, int O, int act, float* __restrict__ xo,
                                   float* __restrict__ vo) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)N * O; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / O; const int o = (int)(i - n * O);
    const float a = hh[n * 2 * O + o], gate = hh[n * 2 * O + O + o];
    xo[i] = act ? tn::silu(a) : a;
#pragma unroll
    for (int c = 0; c < 3; ++c) vo[(n * 3 + c) * O + o] = gate * V2[(n * 3 + c) * O + o];
  }
}
// g_hh [N, 2O] = [g_xo * (act ? SiLU'(xo) : 1) | sum_c g_vo * V2],  g_V2 = gate * g_vo   (g_xo / g_vo may be NULL = zero)
__global__ void geb_out_bwd_kernel(const float* __restrict__ hh, const float* __restrict__ V2, const float* __restrict__ g_xo, int ldgx,
                                   const float* __restrict__ g_vo, int N, int O, int act, float* __restrict__ g_hh, float* __restrict__ g_V2) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)N * O; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / O; const int o = (int)(i - n * O);
    const float a = hh[n * 2 * O + o], gate = hh[n * 2 * O + O + o];
    const float gx = g_xo ? g_xo[n * ldgx + o] : 0.f;
    g_hh[n * 2 * O + o] = act ? gx * tn::silu_grad(a) : gx;
    float gg = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float gv = g_vo ? g_vo[(n * 3 + c) * O + o] : 0.f;
      gg = fmaf(gv, V2[(n * 3 + c) * O + o], gg);
      g_V2[(n * 3 + c) * O + o] = gate * gv;
    }
    g_hh[n * 2 * O + O + o] = gg;
  }
}
// narrow layers (Nn <= 8 outputs, e.g. the last block's Linear(F/2 -> 2) and vec2_proj(F/2 -> 1)): warp per row.
// fwd: out[r][c] = sum_k A[r][k] W[k][c] (+ b[c]);   bwd: dA[r][k] (+)= sum_c dOut[r][c] Wn[c][k]   (Wn = the nn.Linear weight)
__global__ void __launch_bounds__(256) rows_narrow_fwd_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int M, int K,
                                                              int Nn, const float* __restrict__ b, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (long long r = blockIdx.x * wpb + (threadIdx.x >> 5); r < M; r += (long long)gridDim.x * wpb) {
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float a = A[r * lda + k];
#pragma unroll
      for (int c = 0; c < 8; ++c) if (c < Nn) acc[c] = fmaf(a, __ldg(W + (size_t)k * Nn + c), acc[c]);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) if (c < Nn) acc[c] = warp_sum(acc[c]);
    if (lane == 0)
#pragma unroll
      for (int c = 0; c < 8; ++c) if (c < Nn) out[r * Nn + c] = acc[c] + (b ? __ldg(b + c) : 0.f);
  }
}
__global__ void rows_narrow_bwd_kernel(const float* __restrict__ dOut, const float* __restrict__ Wn, int M, int K, int Nn, int accumulate,
                                       const float* __restrict__ aux, float* __restrict__ dA) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)M * K; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / K; const int k = (int)(i - r * K);
    float s = 0.f;
    for (int c = 0; c < Nn; ++c) s = fmaf(dOut[r * Nn + c], __ldg(Wn + (size_t)c * K + k), s);
    if (aux) s *= tn::silu_grad(aux[i]);          // backward through the activation that produced the layer's input
    dA[i] = accumulate ? dA[i] + s : s;
  }
}

template <int BM, int BN, int EPI>
int launch_gemm_bm(const float* A, int lda, const float* W, int ldw, int M, int K, int N, const float* vec, const float* aux, int ldaux,
                   float* zout, float* out, int ldo, cudaStream_t st) {
  const size_t smem = (size_t)GST * (BM * ALD + BK * (BN + 8)) * sizeof(float);
  cudaFuncSetAttribute(rows_gemm_kernel<BM, BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((M + BM - 1) / BM, N / BN);
  rows_gemm_kernel<BM, BN, EPI><<<grid, GT, smem, st>>>(A, lda, W, ldw, M, K, N, vec, aux, ldaux, zout, out, ldo);
  return tmd_check_launch("rows_gemm");
}
template <int BN, int EPI>
int launch_gemm(const float* A, int lda, const float* W, int ldw, int M, int K, int N, const float* vec, const float* aux, int ldaux,
                float* zout, float* out, int ldo, cudaStream_t st) {
  // 64-row tiles by default.  The 128-row form (two m16 tiles per warp share each B fragment: ~20 % fewer instructions per MMA) was
  // measured SLOWER on both bench workloads -- C4 4.251 vs 4.220 ms per frame, C3 7.16 vs 7.09 ms per batch: half as many CTAs to
  // cover the L2 latency of the operand stream -- and is kept behind TMD_GEMM_BM=128 as a tuning aid.
  static const int forced = getenv("TMD_GEMM_BM") ? atoi(getenv("TMD_GEMM_BM")) : 0;
  const bool big = forced == 128;
  return big ? launch_gemm_bm<128, BN, EPI>(A, lda, W, ldw, M, K, N, vec, aux, ldaux, zout, out, ldo, st)
             : launch_gemm_bm<64, BN, EPI>(A, lda, W, ldw, M, K, N, vec, aux, ldaux, zout, out, ldo, st);
}

}  // namespace

extern "C" {

// out[M, N] = epilogue(A[M, K] . W[K, N]); lda / ldw / ldo / ldaux: row strides in floats (multiples of 4).  K % 16 == 0, N % 32 == 0.
// epi 0: z = acc + vec[col]; zout (may be NULL, row stride ldo) = z; out = SiLU(z)
// epi 1: out = acc * SiLU'(aux[row, col])
// epi 2: out = acc * vec[col]
// epi 3: out = acc (+ vec[col] when vec != NULL)          plain Linear / its transposed backward
// epi 4: out = acc + aux[row, col]                         backward of a tensor consumed by two Linears
int tmd_rows_gemm(const float* A, int lda, const float* W, int ldw, int M, int K, int N, int epi, const float* vec, const float* aux,
                  int ldaux, float* zout, float* out, int ldo, void* stream) {
  TMD_REQUIRE(A && W && out, "rows_gemm: null pointer");
  TMD_REQUIRE(M >= 0 && K > 0 && K % BK == 0 && N > 0 && N % 32 == 0, "rows_gemm: need K % 16 == 0 and N % 32 == 0");
  TMD_REQUIRE(lda % 4 == 0 && ldw % 4 == 0 && ldo % 2 == 0 && lda >= K && ldw >= N && ldo >= N, "rows_gemm: bad leading dimension");
  TMD_REQUIRE(epi >= 0 && epi <= 4, "rows_gemm: unknown epilogue");
  TMD_REQUIRE((epi == EPI_MUL_SILU_GRAD || epi == EPI_ADD) ? (aux && ldaux >= N && ldaux % 2 == 0) : (epi == EPI_LINEAR || vec != nullptr),
              "rows_gemm: epilogue operand missing");
  if (M == 0) return TMD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool wide = N % 64 == 0;
#define TMD_GEMM_CASE(E)                                                                                              \
  return wide ? launch_gemm<64, E>(A, lda, W, ldw, M, K, N, vec, aux, ldaux, zout, out, ldo, st)                      \
              : launch_gemm<32, E>(A, lda, W, ldw, M, K, N, vec, aux, ldaux, zout, out, ldo, st)
  if (epi == EPI_BIAS_SILU) { TMD_GEMM_CASE(EPI_BIAS_SILU); }
  if (epi == EPI_MUL_SILU_GRAD) { TMD_GEMM_CASE(EPI_MUL_SILU_GRAD); }
  if (epi == EPI_COLSCALE) { TMD_GEMM_CASE(EPI_COLSCALE); }
  if (epi == EPI_LINEAR) { TMD_GEMM_CASE(EPI_LINEAR); }
  TMD_GEMM_CASE(EPI_ADD);
#undef TMD_GEMM_CASE
}

// mode 3: xn [N, 3F] = LayerNorm_{3F}(3 I^2, |A|^2, |S|^2)   (tensornet.py:394-400);  mode 2: xn [N, F] = LayerNorm_F(|X|^2)
// (tensornet.py:588-590).  mu_rstd [N, 2].  F <= 256.
int tmd_norm_ln_fwd(const float* I, const float* A, const float* S, int N, int F, int mode, const float* ln_w, const float* ln_b, double eps,
                    float* xn, float* mu_rstd, void* stream) {
  TMD_REQUIRE(I && A && S && ln_w && ln_b && xn && mu_rstd, "norm_ln_fwd: null pointer");
  TMD_REQUIRE(F > 0 && F <= 256 && (mode == 2 || mode == 3), "norm_ln_fwd: F <= 256, mode 2 or 3");
  if (N == 0) return TMD_OK;
  const int blocks = tmd_blocks_for((long long)N * 32, 256);
  if (mode == 3) norm_ln_fwd_kernel<3><<<blocks, 256, 0, (cudaStream_t)stream>>>(CI{I, A, S}, N, F, ln_w, ln_b, (float)eps, xn, mu_rstd);
  else norm_ln_fwd_kernel<2><<<blocks, 256, 0, (cudaStream_t)stream>>>(CI{I, A, S}, N, F, ln_w, ln_b, (float)eps, xn, mu_rstd);
  return tmd_check_launch("norm_ln_fwd");
}

// g [N, K] = gradient w.r.t. xn already multiplied by the LayerNorm weight; gX = add (all NULL or all given) + d/dX.
int tmd_norm_ln_bwd(const float* I, const float* A, const float* S, int N, int F, int mode, const float* g, const float* mu_rstd,
                    const float* addI, const float* addA, const float* addS, float* gI, float* gA, float* gS, void* stream) {
  TMD_REQUIRE(I && A && S && g && mu_rstd && gI && gA && gS, "norm_ln_bwd: null pointer");
  TMD_REQUIRE((addI && addA && addS) || (!addI && !addA && !addS), "norm_ln_bwd: the add triple is all-or-nothing");
  TMD_REQUIRE(F > 0 && F <= 256 && (mode == 2 || mode == 3), "norm_ln_bwd: F <= 256, mode 2 or 3");
  if (N == 0) return TMD_OK;
  const int blocks = tmd_blocks_for((long long)N * 32, 256);
  if (mode == 3) norm_ln_bwd_kernel<3><<<blocks, 256, 0, (cudaStream_t)stream>>>(CI{I, A, S}, N, F, g, mu_rstd, CI{addI, addA, addS}, MI{gI, gA, gS});
  else norm_ln_bwd_kernel<2><<<blocks, 256, 0, (cudaStream_t)stream>>>(CI{I, A, S}, N, F, g, mu_rstd, CI{addI, addA, addS}, MI{gI, gA, gS});
  return tmd_check_launch("norm_ln_bwd");
}

// e [N] = (w3 . h2 + b3) * std + atomref[z]   (output_modules.py:79-117, priors/atomref.py:66); atomref may be NULL.
int tmd_head_energy(const float* h2, int N, int H, const float* w3, const float* b3, const float* std_, const int64_t* z, const float* atomref,
                    int n_ref, float* e, void* stream) {
  TMD_REQUIRE(h2 && w3 && b3 && std_ && e && H > 0, "head_energy: null pointer");
  TMD_REQUIRE(!atomref || z, "head_energy: atomref without z");
  if (N == 0) return TMD_OK;
  head_energy_kernel<<<tmd_blocks_for((long long)N * 32, 256), 256, 0, (cudaStream_t)stream>>>(h2, N, H, w3, b3, std_, z, atomref, n_ref, e);
  return tmd_check_launch("head_energy");
}

// dz2 [N, H] = gy[batch] * std * w3 * SiLU'(z2): the seed of the head's backward for the upstream gradient gy [M] (forces: ones).
int tmd_head_seed(const float* gy, const int64_t* batch, int M, const float* std_, const float* w3, const float* z2, int N, int H, float* dz2,
                  void* stream) {
  TMD_REQUIRE(gy && batch && std_ && w3 && z2 && dz2 && H > 0, "head_seed: null pointer");
  if (N == 0) return TMD_OK;
  head_seed_kernel<<<tmd_blocks_for((long long)N * H, 256), 256, 0, (cudaStream_t)stream>>>(gy, batch, M, std_, w3, z2, N, H, dz2);
  return tmd_check_launch("head_seed");
}

// ---- a15 (ET output head): pieces of GatedEquivariantBlock around the GEMM launches (reference models/utils.py:583-655)
int tmd_geb_cat_fwd(const float* x, const float* P1, int N, int H, float* cat, void* stream) {
  TMD_REQUIRE(x && P1 && cat && H > 0, "geb_cat_fwd: null pointer");
  if (N == 0) return TMD_OK;
  geb_cat_fwd_kernel<<<tmd_blocks_for((long long)N * H, 256), 256, 0, (cudaStream_t)stream>>>(x, P1, N, H, cat);
  return tmd_check_launch("geb_cat_fwd");
}
int tmd_geb_cat_bwd(const float* P1, const float* cat, const float* g_cat, int N, int H, float* g_P1, void* stream) {
  TMD_REQUIRE(P1 && cat && g_cat && g_P1 && H > 0, "geb_cat_bwd: null pointer");
  if (N == 0) return TMD_OK;
  geb_cat_bwd_kernel<<<tmd_blocks_for((long long)N * H, 256), 256, 0, (cudaStream_t)stream>>>(P1, cat, g_cat, N, H, g_P1);
  return tmd_check_launch("geb_cat_bwd");
}
int tmd_geb_out_fwd(const float* hh, const float* V2, int N, int O, int act, float* xo, float* vo, void* stream) {
  TMD_REQUIRE(hh && V2 && xo && vo && O > 0, "geb_out_fwd: null pointer");
  if (N == 0) return TMD_OK;
  geb_out_fwd_kernel<<<tmd_blocks_for((long long)N * O, 256), 256, 0, (cudaStream_t)stream>>>(hh, V2, N, O, act, xo, vo);
  return tmd_check_launch("geb_out_fwd");
}
int tmd_geb_out_bwd(const float* hh, const float* V2, const float* g_xo, int ldgx, const float* g_vo, int N, int O, int act, float* g_hh,
                    float* g_V2, void* stream) {
  TMD_REQUIRE(hh && V2 && g_hh && g_V2 && O > 0 && (!g_xo || ldgx >= O), "geb_out_bwd: null pointer / bad stride");
  if (N == 0) return TMD_OK;
  geb_out_bwd_kernel<<<tmd_blocks_for((long long)N * O, 256), 256, 0, (cudaStream_t)stream>>>(hh, V2, g_xo, ldgx, g_vo, N, O, act, g_hh, g_V2);
  return tmd_check_launch("geb_out_bwd");
}
// out[M, Nn] = A[M, K] . W[K, Nn] (+ b), Nn <= 8
int tmd_rows_narrow_fwd(const float* A, int lda, const float* W, int M, int K, int Nn, const float* b, float* out, void* stream) {
  TMD_REQUIRE(A && W && out && K > 0 && Nn > 0 && Nn <= 8 && lda >= K, "rows_narrow_fwd: need 1 <= Nn <= 8");
  if (M == 0) return TMD_OK;
  rows_narrow_fwd_kernel<<<tmd_blocks_for((long long)M * 32, 256), 256, 0, (cudaStream_t)stream>>>(A, lda, W, M, K, Nn, b, out);
  return tmd_check_launch("rows_narrow_fwd");
}
// dA[M, K] (+)= (dOut[M, Nn] . Wn[Nn, K]) (* SiLU'(aux[M, K]) when aux != NULL), Nn <= 8
int tmd_rows_narrow_bwd(const float* dOut, const float* Wn, int M, int K, int Nn, int accumulate, const float* aux, float* dA, void* stream) {
  TMD_REQUIRE(dOut && Wn && dA && K > 0 && Nn > 0 && Nn <= 8, "rows_narrow_bwd: need 1 <= Nn <= 8");
  if (M == 0) return TMD_OK;
  rows_narrow_bwd_kernel<<<tmd_blocks_for((long long)M * K, 256), 256, 0, (cudaStream_t)stream>>>(dOut, Wn, M, K, Nn, accumulate, aux, dA);
  return tmd_check_launch("rows_narrow_bwd");
}

}  // extern "C"
