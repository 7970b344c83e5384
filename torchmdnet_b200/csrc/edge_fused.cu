// a4/a5/a6/a7 fused: the per-edge radial pipeline of TensorNet on unique atom PAIRS.
//
// The reference evaluates, per interaction layer, rbf(r) -> 3 x (Linear + SiLU) -> * C(r) on every
// directed edge with cuBLAS SGEMMs and ~10 elementwise launches, materialising [E,R], [E,F], [E,2F],
// [E,3F] (+ the same again for autograd) -- models/tensornet.py:738-743; the embedding does the same for
// its three radial projections (:558-567).  All of it is a function of ONE scalar per edge, r_ij = r_ji.
// B200-first restructuring:
//   * evaluate it once per unordered pair (half the FLOPs), directly from r (rbf and cutoff computed in
//     registers, never written to HBM);
//   * evaluate the tangent d/dr of the whole chain in the same pass (forward-mode: the weights are read
//     once for value and tangent), so the force backward of an interaction layer needs no GEMM at all:
//     dE/dr_p = <dE/df_p, f'_p>, fused into the transposed message-passing gather;
//   * no intermediate activation ever leaves the SM.
// This file holds the pair index builder, the (CUDA-core, fp32-exact) chain kernel, and the fused
// message-passing kernels that consume pair-indexed weights.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------- pair index
// canonical directed edge of a pair = the one with j <= i; rows are sorted by j so they are the first
// c_i slots of row i.  pair id = (exclusive scan of c)[i] + offset in row.
__global__ void pair_count_kernel(const int* __restrict__ indptr, const int* __restrict__ nbr_j,
                                  const int* __restrict__ num_edges, int max_edges, int n, int* __restrict__ cnt) {
  const int E = min(*num_edges, max_edges);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int lo = min(indptr[i], E), hi = min(indptr[i + 1], E);
    const int s = lo;
    while (lo < hi) { int m = (lo + hi) >> 1; if (nbr_j[m] <= i) lo = m + 1; else hi = m; }
    cnt[i] = lo - s;
  }
}
__global__ void pair_fill_kernel(const int* __restrict__ indptr, const int* __restrict__ nbr_i,
                                 const int* __restrict__ nbr_j, const int* __restrict__ rev,
                                 const int* __restrict__ num_edges, int max_edges, const int* __restrict__ pair_ptr,
                                 const float* __restrict__ dist, int* __restrict__ pid, int* __restrict__ pair_edge,
                                 float* __restrict__ pair_r) {
  const int E = min(*num_edges, max_edges);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    const int i = nbr_i[e], j = nbr_j[e];
    if (j <= i) {
      const int p = pair_ptr[i] + (e - indptr[i]);
      pid[e] = p; pair_edge[p] = e; pair_r[p] = dist[e];
    } else {
      const int e2 = rev[e];  // (j, i): canonical, lives in row j
      pid[e] = pair_ptr[j] + (e2 - indptr[j]);
    }
  }
}
__global__ void __launch_bounds__(1024) scan_small_kernel(const int* __restrict__ in, int n, int* __restrict__ out,
                                                           int* __restrict__ total_out) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) { carry_s = 0; out[0] = 0; }
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    int x = (i < n) ? in[i] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    if (w == 0) {
      int t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
      warp_tot[lane] = t;
    }
    __syncthreads();
    const int incl = x + (w > 0 ? warp_tot[w - 1] : 0) + carry_s;
    if (i < n) out[i + 1] = incl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

// ---------------------------------------------------------------------------------------------- radial chain (JVP)
struct ChainParams {
  const float* WT[3];   // transposed weights, layer l: [K_l][N_l] row-major (K_0 = R)
  const float* bias[3]; // [N_l]
  int dims[4];          // R, N_0, N_1, N_2
  int n_layers;         // 1..3
  int act;              // SiLU after every layer
  int mul_cutoff;       // multiply the output by CosineCutoff(lo, hi)(r)
  const float* means;   // [R]  ExpNormalSmearing
  const float* betas;   // [R]
  float alpha, lo, hi;
};

constexpr int TM = 32;       // pairs per tile (value rows 0..31, tangent rows 32..63)
constexpr int CH_THREADS = 256;
constexpr int MAX_COLS = 12; // N_l / 32 <= 12  (N <= 384)

// one CTA = one tile of TM pairs; thread (wy = warp 0..7 -> pairs 4wy..4wy+3, lane -> columns lane + 32c)
// smem ping-pong: bufA [2*TM][ldA] holds layer inputs of even layers (R, then N_1), bufB [2*TM][ldB] odd (N_0)
template <bool JVP>
__global__ void __launch_bounds__(CH_THREADS) pair_chain_kernel(const float* __restrict__ pair_r,
                                                                const int* __restrict__ num_unique, int max_pairs,
                                                                ChainParams P, float* __restrict__ out_f,
                                                                float* __restrict__ out_fp, int ldA, int ldB) {
  extern __shared__ float smem[];
  float* bufA = smem;
  float* bufB = smem + 2 * TM * ldA;
  __shared__ float s_r[TM], s_c[TM], s_dc[TM];
  const int n_pairs = min(*num_unique, max_pairs);
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int R = P.dims[0];
  for (int tile = blockIdx.x; tile * TM < n_pairs; tile += gridDim.x) {
    const int p0 = tile * TM;
    __syncthreads();
    // ---- layer-0 input: rbf(r) (value rows) and d rbf/dr (tangent rows)
    if (threadIdx.x < TM) {
      const int p = p0 + threadIdx.x;
      const float r = (p < n_pairs) ? pair_r[p] : 0.f;
      float c, dc;
      tn::cosine_cutoff(r, P.lo, P.hi, &c, &dc);
      s_r[threadIdx.x] = r; s_c[threadIdx.x] = c; s_dc[threadIdx.x] = dc;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < TM * R; idx += CH_THREADS) {
      const int m = idx / R, k = idx - m * R;
      const float r = s_r[m];
      float c0, dc0, v, dv;
      tn::cosine_cutoff(r, 0.f, P.hi, &c0, &dc0);  // the rbf envelope is CosineCutoff(0, hi) (utils.py:371)
      const float u = expf(P.alpha * (P.lo - r));
      tn::expnorm_rbf(u, P.alpha, c0, dc0, P.means[k], P.betas[k], &v, &dv);
      bufA[m * ldA + k] = v;
      if (JVP) bufA[(TM + m) * ldA + k] = dv;
    }
    __syncthreads();
    // ---- layers
    for (int l = 0; l < P.n_layers; ++l) {
      const int K = P.dims[l], N = P.dims[l + 1];
      const float* in = (l & 1) ? bufB : bufA;
      float* outb = (l & 1) ? bufA : bufB;
      const int ldi = (l & 1) ? ldB : ldA, ldo = (l & 1) ? ldA : ldB;
      const float* __restrict__ WT = P.WT[l];
      const int ncol = N >> 5;  // columns per thread
      float acc_v[4][MAX_COLS], acc_t[4][MAX_COLS];
#pragma unroll
      for (int c = 0; c < MAX_COLS; ++c) {
        const float b = (c < ncol) ? __ldg(P.bias[l] + lane + 32 * c) : 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) { acc_v[q][c] = b; acc_t[q][c] = 0.f; }
      }
      const float* av = in + (4 * wy) * ldi;
      const float* at = in + (TM + 4 * wy) * ldi;
      for (int k = 0; k < K; k += 4) {
        float4 a_v[4], a_t[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          a_v[q] = *reinterpret_cast<const float4*>(av + q * ldi + k);
          if (JVP) a_t[q] = *reinterpret_cast<const float4*>(at + q * ldi + k);
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float* wrow = WT + (size_t)(k + kk) * N + lane;
#pragma unroll
          for (int c = 0; c < MAX_COLS; ++c) {
            if (c < ncol) {
              const float b = __ldg(wrow + 32 * c);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float x = kk == 0 ? a_v[q].x : kk == 1 ? a_v[q].y : kk == 2 ? a_v[q].z : a_v[q].w;
                acc_v[q][c] = fmaf(x, b, acc_v[q][c]);
                if (JVP) {
                  const float y = kk == 0 ? a_t[q].x : kk == 1 ? a_t[q].y : kk == 2 ? a_t[q].z : a_t[q].w;
                  acc_t[q][c] = fmaf(y, b, acc_t[q][c]);
                }
              }
            }
          }
        }
      }
      // ---- epilogue: activation (+ tangent), last layer -> HBM, else -> smem for the next layer
      const bool last = (l == P.n_layers - 1);
#pragma unroll
      for (int c = 0; c < MAX_COLS; ++c) {
        if (c < ncol) {
          const int col = lane + 32 * c;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int m = 4 * wy + q;
            float v = acc_v[q][c], t = acc_t[q][c];
            if (P.act) { const float g = tn::silu_grad(v); v = tn::silu(v); t *= g; }
            if (last) {
              if (P.mul_cutoff) { t = t * s_c[m] + v * s_dc[m]; v *= s_c[m]; }
              const int p = p0 + m;
              if (p < n_pairs) {
                out_f[(size_t)p * N + col] = v;
                if (JVP) out_fp[(size_t)p * N + col] = t;
              }
            } else {
              outb[m * ldo + col] = v;
              if (JVP) outb[(TM + m) * ldo + col] = t;
            }
          }
        }
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------- fused message passing
struct CIrr { const float* I; const float* A; const float* S; };
struct MIrr { float* I; float* A; float* S; };

__device__ __forceinline__ void ld9(const CIrr& T, long long n, int F, int h, float4* c) {
  c[0] = ldg4(T.I + n * F + h);
#pragma unroll
  for (int k = 0; k < 3; ++k) c[1 + k] = ldg4(T.A + (n * 3 + k) * F + h);
#pragma unroll
  for (int k = 0; k < 5; ++k) c[4 + k] = ldg4(T.S + (n * 5 + k) * F + h);
}
__device__ __forceinline__ void st9(const MIrr& T, long long n, int F, int h, const float4* c) {
  st4(T.I + n * F + h, c[0]);
#pragma unroll
  for (int k = 0; k < 3; ++k) st4(T.A + (n * 3 + k) * F + h, c[1 + k]);
#pragma unroll
  for (int k = 0; k < 5; ++k) st4(T.S + (n * 5 + k) * F + h, c[4 + k]);
}
__device__ __forceinline__ float4 fma4(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// backward of  M_b = sum_{e=(b,j)} f_p(e) * Y_j  with symmetric pair weights:
//   dY_b      = sum_{e=(b,j)} f_p * dM_j                         (same gather shape as the forward)
//   dE/dr_p   = < f'_p , dM_b*Y_j + dM_j*Y_b >  (reduced over components of each irrep and all channels),
// the latter evaluated once per pair by the warp that owns the canonical edge (j <= b).  Requires F <= 128*? any
// multiple of 4; the per-pair scalar is accumulated over channel chunks in registers and reduced with shuffles.
__global__ void __launch_bounds__(256, 2) mp_pair_bwd_kernel(CIrr Y, CIrr dM, const float* __restrict__ f,
                                                             const float* __restrict__ fp, const int* __restrict__ pid,
                                                             const int* __restrict__ nbr_j, const int* __restrict__ indptr,
                                                             int n, int F, MIrr dY, float* __restrict__ d_pair_r) {
  // per-warp staging of the row atom's own Y_b and dM_b (18 x F floats) keeps them out of the register file
  extern __shared__ float4 sm4[];
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, wid = threadIdx.x >> 5;
  const int chunks = (F + 127) / 128;            // channel chunks of 128 (one float4 per lane)
  float4* myb = sm4 + (size_t)wid * 18 * chunks * 32;  // [chunk][18][32 lanes]
  for (int b = blockIdx.x * wpb + wid; b < n; b += gridDim.x * wpb) {
    const int k0 = indptr[b], k1 = indptr[b + 1];
    __syncwarp();
    for (int ch = 0; ch < chunks; ++ch) {
      const int h = ch * 128 + lane * 4;
      float4 t[9];
      if (h < F) {
        ld9(Y, b, F, h, t);
#pragma unroll
        for (int c = 0; c < 9; ++c) myb[(ch * 18 + c) * 32 + lane] = t[c];
        ld9(dM, b, F, h, t);
#pragma unroll
        for (int c = 0; c < 9; ++c) myb[(ch * 18 + 9 + c) * 32 + lane] = t[c];
      }
    }
    __syncwarp();
    // ---- pass 1: dY_b = sum_k f_p * dM_j  (chunk outer so the accumulators stay in registers)
    for (int ch = 0; ch < chunks; ++ch) {
      const int h = ch * 128 + lane * 4;
      if (h >= F) continue;
      float4 acc[9];
#pragma unroll
      for (int c = 0; c < 9; ++c) acc[c] = zero4();
#pragma unroll 2
      for (int k = k0; k < k1; ++k) {
        const long long j = nbr_j[k], p = pid[k];
        const float* fe = f + p * 3 * (long long)F + h;
        const float4 w1 = ldg4(fe), w2 = ldg4(fe + F), w3 = ldg4(fe + 2 * F);
        float4 dj[9];
        ld9(dM, j, F, h, dj);
        acc[0] = fma4(w1, dj[0], acc[0]);
#pragma unroll
        for (int c = 1; c < 4; ++c) acc[c] = fma4(w2, dj[c], acc[c]);
#pragma unroll
        for (int c = 4; c < 9; ++c) acc[c] = fma4(w3, dj[c], acc[c]);
      }
      st9(dY, b, F, h, acc);
    }
    // ---- pass 2: canonical pairs (j <= b; rows are sorted so they come first): dE/dr_p
    if (fp) {
      for (int k = k0; k < k1; ++k) {
        const long long j = nbr_j[k];
        if (j > b) break;
        const long long p = pid[k];
        const float sym = (j == b) ? 0.f : 1.f;  // a self loop is a single directed edge
        const float4 s4 = make_float4(sym, sym, sym, sym);
        float s = 0.f;
        for (int ch = 0; ch < chunks; ++ch) {
          const int h = ch * 128 + lane * 4;
          if (h >= F) continue;
          const float4* yb = myb + (ch * 18) * 32 + lane;
          const float4* db = myb + (ch * 18 + 9) * 32 + lane;
          const float* fpe = fp + p * 3 * (long long)F + h;
          float4 g = fma4(s4, mul4(ldg4(dM.I + j * F + h), yb[0]), mul4(db[0], ldg4(Y.I + j * F + h)));
          s += dot4(ldg4(fpe), g);
          g = zero4();
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            g = fma4(db[(1 + c) * 32], ldg4(Y.A + (j * 3 + c) * F + h), g);
            g = fma4(s4, mul4(ldg4(dM.A + (j * 3 + c) * F + h), yb[(1 + c) * 32]), g);
          }
          s += dot4(ldg4(fpe + F), g);
          g = zero4();
#pragma unroll
          for (int c = 0; c < 5; ++c) {
            g = fma4(db[(4 + c) * 32], ldg4(Y.S + (j * 5 + c) * F + h), g);
            g = fma4(s4, mul4(ldg4(dM.S + (j * 5 + c) * F + h), yb[(4 + c) * 32]), g);
          }
          s += dot4(ldg4(fpe + 2 * F), g);
        }
        s = warp_sum(s);
        if (lane == 0) d_pair_r[p] = s;
      }
    }
  }
}

// Single-pass variant of the backward above, emitting the radial gradient per DIRECTED slot instead of per pair:
//   dY_b = sum_k f_p dM_j     and     g_dist[k] = < f'_p , dM_j * Y_b >     for slot k = (b, j)
// g_dist[k] is the contribution of the transposed edge (j <- b) to dE/dr_p; the position-gradient kernel combines
// g_dist[k] and g_dist[rev k] (g_e - g_rev(e) with opposite unit vectors), which sums the two halves of the pair.
// Only dM_j is gathered (Y_j is not needed at all): half the gather traffic of mp_pair_bwd_kernel.
__global__ void __launch_bounds__(256, 2) mp_edge_bwd_kernel(CIrr Y, CIrr dM, const float* __restrict__ f,
                                                             const float* __restrict__ fp, const int* __restrict__ pid,
                                                             const int* __restrict__ nbr_j, const int* __restrict__ indptr,
                                                             int n, int F, MIrr dY, float* __restrict__ g_dist) {
  extern __shared__ float4 sm4[];
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, wid = threadIdx.x >> 5;
  const int chunks = (F + 127) / 128;
  float4* myb = sm4 + (size_t)wid * 9 * chunks * 32;   // own Y_b: [chunk][9][32 lanes]
  for (int b = blockIdx.x * wpb + wid; b < n; b += gridDim.x * wpb) {
    const int k0 = indptr[b], k1 = indptr[b + 1];
    __syncwarp();
    for (int ch = 0; ch < chunks; ++ch) {
      const int h = ch * 128 + lane * 4;
      if (h < F) {
        float4 t[9];
        ld9(Y, b, F, h, t);
#pragma unroll
        for (int c = 0; c < 9; ++c) myb[(ch * 9 + c) * 32 + lane] = t[c];
      }
    }
    __syncwarp();
    if (chunks == 1) {
      const int h = lane * 4;
      const bool on = h < F;
      const float4* yb = myb + lane;
      float4 acc[9];
#pragma unroll
      for (int c = 0; c < 9; ++c) acc[c] = zero4();
#pragma unroll 2
      for (int k = k0; k < k1; ++k) {
        const long long j = nbr_j[k], p = pid[k];
        float t = 0.f;
        if (on) {
          const float* fe = f + p * 3 * (long long)F + h;
          const float4 w1 = ldg4(fe), w2 = ldg4(fe + F), w3 = ldg4(fe + 2 * F);
          float4 dj[9];
          ld9(dM, j, F, h, dj);
          acc[0] = fma4(w1, dj[0], acc[0]);
          float4 g2 = zero4(), g3 = zero4();
#pragma unroll
          for (int c = 1; c < 4; ++c) { acc[c] = fma4(w2, dj[c], acc[c]); g2 = fma4(dj[c], yb[c * 32], g2); }
#pragma unroll
          for (int c = 4; c < 9; ++c) { acc[c] = fma4(w3, dj[c], acc[c]); g3 = fma4(dj[c], yb[c * 32], g3); }
          if (fp) {
            const float* fpe = fp + p * 3 * (long long)F + h;
            t = dot4(ldg4(fpe), mul4(dj[0], yb[0])) + dot4(ldg4(fpe + F), g2) + dot4(ldg4(fpe + 2 * F), g3);
          }
        }
        if (fp) {
          t = warp_sum(t);
          if (lane == 0) g_dist[k] = (j == b) ? 0.f : t;
        }
      }
      if (on) st9(dY, b, F, h, acc);
    } else {
      for (int ch = 0; ch < chunks; ++ch) {
        const int h = ch * 128 + lane * 4;
        if (h >= F) continue;
        float4 acc[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) acc[c] = zero4();
        for (int k = k0; k < k1; ++k) {
          const long long j = nbr_j[k], p = pid[k];
          const float* fe = f + p * 3 * (long long)F + h;
          const float4 w1 = ldg4(fe), w2 = ldg4(fe + F), w3 = ldg4(fe + 2 * F);
          float4 dj[9];
          ld9(dM, j, F, h, dj);
          acc[0] = fma4(w1, dj[0], acc[0]);
#pragma unroll
          for (int c = 1; c < 4; ++c) acc[c] = fma4(w2, dj[c], acc[c]);
#pragma unroll
          for (int c = 4; c < 9; ++c) acc[c] = fma4(w3, dj[c], acc[c]);
        }
        st9(dY, b, F, h, acc);
      }
      if (fp) {
        for (int k = k0; k < k1; ++k) {
          const long long j = nbr_j[k], p = pid[k];
          float t = 0.f;
          for (int ch = 0; ch < chunks; ++ch) {
            const int h = ch * 128 + lane * 4;
            if (h >= F) continue;
            const float4* yb = myb + (ch * 9) * 32 + lane;
            float4 dj[9];
            ld9(dM, j, F, h, dj);
            float4 g2 = zero4(), g3 = zero4();
#pragma unroll
            for (int c = 1; c < 4; ++c) g2 = fma4(dj[c], yb[c * 32], g2);
#pragma unroll
            for (int c = 4; c < 9; ++c) g3 = fma4(dj[c], yb[c * 32], g3);
            const float* fpe = fp + p * 3 * (long long)F + h;
            t += dot4(ldg4(fpe), mul4(dj[0], yb[0])) + dot4(ldg4(fpe + F), g2) + dot4(ldg4(fpe + 2 * F), g3);
          }
          t = warp_sum(t);
          if (lane == 0) g_dist[k] = (j == b) ? 0.f : t;
        }
      }
    }
  }
}

// embedding: w_e[k,:] = C(r_e) (U_i + V_j) dp_p[k,:]  accumulated into I/A/S of atom i  (tensornet.py:554-585)
// The unit vector rhat = vec / |vec| (0 for the self loop, tensornet.py:363-366) is formed in the kernel from the neighbour
// list's (vec, dist); U / V are either per-atom rows (zidx == NULL) or per-element tables indexed through the atomic numbers
// zidx (U_i = U[z_i], V_j = V[z_j]): emb2([Z_i ; Z_j]) = W_u emb(z_i) + W_v emb(z_j) + b depends on the elements only, so
// for frozen weights it is two [max_z, F] tables built once per model (the analogue of the reference's Zij table,
// TensorEmbedding.setup_for_inference, tensornet.py:509-533) -- 64 KB, L1-resident.
// Per-edge scalars are prepared LANE-PARALLEL for 32 edges at a time (pair id, element row of the neighbour, unit vector,
// cutoff and its derivative -- one sincos per EDGE instead of one per lane) and broadcast with shuffles, so no dependent
// index -> gather chain sits on the per-edge critical path; the radial rows of the next edge are prefetched into registers.
struct EdgeMeta { int p; int zj; float r0, r1, r2, cut, dcut, inv; };
__device__ __forceinline__ EdgeMeta load_edge_meta(const float* __restrict__ vec, const float* __restrict__ dist,
                                                   const int* __restrict__ pid, const int* __restrict__ nbr_j,
                                                   const int64_t* __restrict__ zidx, int k, bool valid, float lo, float hi) {
  EdgeMeta m{0, 0, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f};
  if (valid) {
    m.p = pid[k];
    const int j = nbr_j[k];
    m.zj = zidx ? (int)zidx[j] : j;
    const float d = __ldg(dist + k);
    m.inv = d == 0.f ? 1.f : 1.f / d;     // self loop: vec = 0 -> rhat = 0 (tensornet.py:363-366)
    m.r0 = __ldg(vec + 3 * (size_t)k) * m.inv; m.r1 = __ldg(vec + 3 * (size_t)k + 1) * m.inv; m.r2 = __ldg(vec + 3 * (size_t)k + 2) * m.inv;
    tn::cosine_cutoff(d, lo, hi, &m.cut, &m.dcut);
  }
  return m;
}
__device__ __forceinline__ float4 splat4(float x) { return make_float4(x, x, x, x); }

__global__ void __launch_bounds__(256) embed_fwd_kernel(const float* __restrict__ vec, const float* __restrict__ dist,
                                                        const float* __restrict__ dp, const int* __restrict__ pid,
                                                        const int* __restrict__ nbr_j, const int* __restrict__ indptr,
                                                        const float* __restrict__ U, const float* __restrict__ V,
                                                        const int64_t* __restrict__ zidx, float lo, float hi, int n, int F, MIrr out) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < n; b += gridDim.x * wpb) {
    const int k0 = indptr[b], k1 = indptr[b + 1];
    const long long ub_row = zidx ? zidx[b] : b;
    for (int hh = 0; hh < F; hh += 128) {      // every lane runs the loop (the shuffles below need the whole warp); lanes past F idle
      const int h = hh + lane * 4;
      const bool on = h < F;
      float4 acc[9];
#pragma unroll
      for (int c = 0; c < 9; ++c) acc[c] = zero4();
      const float4 ub = on ? ldg4(U + ub_row * F + h) : zero4();
      for (int e0 = k0; e0 < k1; e0 += 32) {
        const EdgeMeta me = load_edge_meta(vec, dist, pid, nbr_j, zidx, e0 + lane, e0 + lane < k1, lo, hi);
        const int m = min(32, k1 - e0);
        const long long p0 = __shfl_sync(0xffffffffu, me.p, 0);
        float4 q1 = zero4(), q2 = zero4(), q3 = zero4();
        if (on) { const float* de = dp + p0 * 3 * F + h; q1 = ldg4(de); q2 = ldg4(de + F); q3 = ldg4(de + 2 * F); }
        for (int i = 0; i < m; ++i) {
          const int zj = __shfl_sync(0xffffffffu, me.zj, i);
          const float cut = __shfl_sync(0xffffffffu, me.cut, i);
          const float r[3] = {__shfl_sync(0xffffffffu, me.r0, i), __shfl_sync(0xffffffffu, me.r1, i), __shfl_sync(0xffffffffu, me.r2, i)};
          const int pn = __shfl_sync(0xffffffffu, me.p, (i + 1) & 31);
          const float4 c1 = q1, c2 = q2, c3 = q3;
          if (!on) continue;
          if (i + 1 < m) {   // prefetch the next edge's radial row
            const float* dn = dp + (long long)pn * 3 * F + h;
            q1 = ldg4(dn); q2 = ldg4(dn + F); q3 = ldg4(dn + 2 * F);
          }
          float4 zc = add4(ub, ldg4(V + (long long)zj * F + h));
          zc = mul4(zc, splat4(cut));
          const float4 w1 = mul4(zc, c1), w2 = mul4(zc, c2), w3 = mul4(zc, c3);
          float a[3], sb[5];
          tn::skew_compact(r, a); tn::sym_basis(r, sb);
          acc[0] = add4(acc[0], w1);
#pragma unroll
          for (int c = 0; c < 3; ++c) acc[1 + c] = fma4(splat4(a[c]), w2, acc[1 + c]);
#pragma unroll
          for (int c = 0; c < 5; ++c) acc[4 + c] = fma4(splat4(sb[c]), w3, acc[4 + c]);
        }
      }
      if (on) st9(out, b, F, h, acc);
    }
  }
}
// backward wrt geometry only (U, V, weights are inference constants): g_dist[e] (through the cutoff, the radial projections
// and the normalisation of rhat) and g_vec[e,3] = g_rhat / |vec|
template <bool ONE_CHUNK>
__global__ void __launch_bounds__(256, 2) embed_bwd_kernel(const float* __restrict__ vec, const float* __restrict__ dist,
                                                        const float* __restrict__ dp, const float* __restrict__ dpp,
                                                        const int* __restrict__ pid, const int* __restrict__ nbr_j,
                                                        const int* __restrict__ indptr, const float* __restrict__ U,
                                                        const float* __restrict__ V, const int64_t* __restrict__ zidx, float lo,
                                                        float hi, int n, int F, CIrr dout, float* __restrict__ g_dist,
                                                        float* __restrict__ g_vec) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < n; b += gridDim.x * wpb) {
    const int k0 = indptr[b], k1 = indptr[b + 1];
    const long long ub_row = zidx ? zidx[b] : b;
    // ONE_CHUNK (F <= 128): the row's own dOut and U are loaded once, not once per edge
    float4 d[9], ub = zero4();
    if (ONE_CHUNK && lane * 4 < F) { ld9(dout, b, F, lane * 4, d); ub = ldg4(U + ub_row * F + lane * 4); }
    for (int e0 = k0; e0 < k1; e0 += 32) {
      const EdgeMeta me = load_edge_meta(vec, dist, pid, nbr_j, zidx, e0 + lane, e0 + lane < k1, lo, hi);
      const int m = min(32, k1 - e0);
      float keep_gd = 0.f, keep_g0 = 0.f, keep_g1 = 0.f, keep_g2 = 0.f;   // lane i keeps the results of edge e0 + i
      float4 q1 = zero4(), q2 = zero4(), q3 = zero4(), t1 = zero4(), t2 = zero4(), t3 = zero4();
      const long long p0 = __shfl_sync(0xffffffffu, me.p, 0);
      if (ONE_CHUNK && lane * 4 < F) {
        const float* de = dp + p0 * 3 * F + lane * 4; const float* dde = dpp + p0 * 3 * F + lane * 4;
        q1 = ldg4(de); q2 = ldg4(de + F); q3 = ldg4(de + 2 * F); t1 = ldg4(dde); t2 = ldg4(dde + F); t3 = ldg4(dde + 2 * F);
      }
      for (int i = 0; i < m; ++i) {
        const int zj = __shfl_sync(0xffffffffu, me.zj, i);
        const float cut = __shfl_sync(0xffffffffu, me.cut, i), dcut = __shfl_sync(0xffffffffu, me.dcut, i);
        const float inv = __shfl_sync(0xffffffffu, me.inv, i);
        const float r[3] = {__shfl_sync(0xffffffffu, me.r0, i), __shfl_sync(0xffffffffu, me.r1, i), __shfl_sync(0xffffffffu, me.r2, i)};
        const long long pc = __shfl_sync(0xffffffffu, me.p, i);
        const long long pn = __shfl_sync(0xffffffffu, me.p, (i + 1) & 31);
        float a[3], sb[5];
        tn::skew_compact(r, a); tn::sym_basis(r, sb);
        float gr[3] = {0.f, 0.f, 0.f};
        float gd = 0.f;
        for (int h = lane * 4; h < F; h += 128) {
          float4 c1, c2, c3, u1, u2, u3;
          if (ONE_CHUNK) {
            c1 = q1; c2 = q2; c3 = q3; u1 = t1; u2 = t2; u3 = t3;
            if (i + 1 < m) {   // prefetch the next edge's radial rows (value and d/dr)
              const float* de = dp + pn * 3 * F + h; const float* dde = dpp + pn * 3 * F + h;
              q1 = ldg4(de); q2 = ldg4(de + F); q3 = ldg4(de + 2 * F); t1 = ldg4(dde); t2 = ldg4(dde + F); t3 = ldg4(dde + 2 * F);
            }
          } else {
            ld9(dout, b, F, h, d); ub = ldg4(U + ub_row * F + h);
            const float* de = dp + pc * 3 * F + h; const float* dde = dpp + pc * 3 * F + h;
            c1 = ldg4(de); c2 = ldg4(de + F); c3 = ldg4(de + 2 * F); u1 = ldg4(dde); u2 = ldg4(dde + F); u3 = ldg4(dde + 2 * F);
          }
          const float4 z = add4(ub, ldg4(V + (long long)zj * F + h));
          // dL/dw_k
          float4 g2 = zero4(), g3 = zero4();
#pragma unroll
          for (int c = 0; c < 3; ++c) g2 = fma4(splat4(a[c]), d[1 + c], g2);
#pragma unroll
          for (int c = 0; c < 5; ++c) g3 = fma4(splat4(sb[c]), d[4 + c], g3);
          // w_k = cut * z * q_k  ->  dw_k/dr = dcut * z * q_k + cut * z * q'_k
          const float4 s1 = fma4(splat4(dcut), c1, mul4(splat4(cut), u1));
          const float4 s2 = fma4(splat4(dcut), c2, mul4(splat4(cut), u2));
          const float4 s3 = fma4(splat4(dcut), c3, mul4(splat4(cut), u3));
          gd += dot4(mul4(z, s1), d[0]) + dot4(mul4(z, s2), g2) + dot4(mul4(z, s3), g3);
          // d/d rhat through skew and the symmetric-traceless basis
          const float4 w2 = mul4(splat4(cut), mul4(z, c2));
          const float4 w3 = mul4(splat4(cut), mul4(z, c3));
          gr[0] -= dot4(w2, d[3]); gr[1] += dot4(w2, d[2]); gr[2] -= dot4(w2, d[1]);
          const float w3v[4] = {w3.x, w3.y, w3.z, w3.w};
          const float dSv[5][4] = {{d[4].x, d[4].y, d[4].z, d[4].w}, {d[5].x, d[5].y, d[5].z, d[5].w},
                                   {d[6].x, d[6].y, d[6].z, d[6].w}, {d[7].x, d[7].y, d[7].z, d[7].w},
                                   {d[8].x, d[8].y, d[8].z, d[8].w}};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float dS[5] = {dSv[0][q], dSv[1][q], dSv[2][q], dSv[3][q], dSv[4][q]};
            tn::sym_basis_bwd_acc(r, dS, w3v[q], gr);
          }
        }
        gd = warp_sum(gd); gr[0] = warp_sum(gr[0]); gr[1] = warp_sum(gr[1]); gr[2] = warp_sum(gr[2]);
        if (lane == i) {
          // rhat = vec * inv with inv = 1/dist (1 for the self loop, whose vec is 0): d/dvec = g_rhat * inv,
          // d/ddist = -(g_rhat . rhat) * inv (0 for the self loop: the masked denominator is a constant there)
          const bool self = (r[0] == 0.f && r[1] == 0.f && r[2] == 0.f) && inv == 1.f;
          keep_gd = gd - (self ? 0.f : (gr[0] * r[0] + gr[1] * r[1] + gr[2] * r[2]) * inv);
          keep_g0 = gr[0] * inv; keep_g1 = gr[1] * inv; keep_g2 = gr[2] * inv;
        }
      }
      if (lane < m) {
        const size_t k = (size_t)e0 + lane;
        g_dist[k] = keep_gd;
        g_vec[3 * k] = keep_g0; g_vec[3 * k + 1] = keep_g1; g_vec[3 * k + 2] = keep_g2;
      }
    }
  }
}

}  // namespace

extern "C" {

// pair index over the symmetric sorted CSR.  workspace: n ints.  pair_ptr[n+1].
int tmd_pair_index(const int* indptr, const int* nbr_i, const int* nbr_j, const int* rev, const int* num_edges,
                   int max_edges, int n, const float* dist, int* pair_ptr, int* pid, int* pair_edge, float* pair_r,
                   int* num_unique, int* workspace, void* stream) {
  TMD_REQUIRE(indptr && nbr_i && nbr_j && rev && num_edges && dist && pair_ptr && pid && pair_edge && pair_r && num_unique && workspace,
              "pair_index: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) { cudaMemsetAsync(num_unique, 0, sizeof(int), st); return tmd_check_launch("pair_index(empty)"); }
  pair_count_kernel<<<tmd_blocks_for(n, 128), 128, 0, st>>>(indptr, nbr_j, num_edges, max_edges, n, workspace);
  scan_small_kernel<<<1, 1024, 0, st>>>(workspace, n, pair_ptr, num_unique);
  pair_fill_kernel<<<tmd_blocks_for(max_edges, 256), 256, 0, st>>>(indptr, nbr_i, nbr_j, rev, num_edges, max_edges, pair_ptr, dist, pid, pair_edge, pair_r);
  return tmd_check_launch("pair_index");
}

// radial chain on pairs: out_f[p,:] = chain(rbf(r_p)) (* cutoff), out_fp = d/dr (may be NULL: value only).
// WT_l: [K_l][N_l] fp32 (transposed nn.Linear weight); dims = {R, N_0, N_1, N_2}; N_l % 32 == 0, N_l <= 384,
// R % 4 == 0, hidden widths % 4 == 0.
int tmd_pair_chain(const float* pair_r, const int* num_unique, int max_pairs, const float* WT0, const float* b0,
                   const float* WT1, const float* b1, const float* WT2, const float* b2, const int* dims, int n_layers,
                   int act, int mul_cutoff, const float* means, const float* betas, double alpha, double lo, double hi,
                   float* out_f, float* out_fp, void* stream) {
  TMD_REQUIRE(pair_r && num_unique && WT0 && b0 && dims && means && betas && out_f, "pair_chain: null pointer");
  TMD_REQUIRE(n_layers >= 1 && n_layers <= 3, "pair_chain: n_layers must be 1..3");
  ChainParams P;
  const float* W[3] = {WT0, WT1, WT2};
  const float* B[3] = {b0, b1, b2};
  int ldA = dims[0], ldB = 4;
  for (int l = 0; l < 3; ++l) { P.WT[l] = W[l]; P.bias[l] = B[l]; }
  for (int l = 0; l <= n_layers; ++l) {
    P.dims[l] = dims[l];
    TMD_REQUIRE(dims[l] > 0 && dims[l] % 4 == 0, "pair_chain: widths must be positive multiples of 4");
    if (l > 0) TMD_REQUIRE(dims[l] % 32 == 0 && dims[l] <= 32 * MAX_COLS, "pair_chain: layer widths must be multiples of 32, <= 384");
    if (l > 0) TMD_REQUIRE(W[l - 1] && B[l - 1], "pair_chain: missing layer weights");
    if (l < n_layers) { if (l % 2 == 0) ldA = max(ldA, dims[l]); else ldB = max(ldB, dims[l]); }
  }
  for (int l = n_layers + 1; l < 4; ++l) P.dims[l] = 0;
  P.n_layers = n_layers; P.act = act; P.mul_cutoff = mul_cutoff; P.means = means; P.betas = betas;
  P.alpha = (float)alpha; P.lo = (float)lo; P.hi = (float)hi;
  ldA += 4; ldB += 4;  // pad rows: conflict-free float4 row reads
  const size_t smem = (size_t)2 * TM * (ldA + ldB) * sizeof(float);
  TMD_REQUIRE(smem <= 200 * 1024, "pair_chain: layer widths exceed shared memory");
  if (max_pairs == 0) return TMD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles = (max_pairs + TM - 1) / TM;
  const int grid = tiles < 148 * 2 ? tiles : 148 * 2;
  if (out_fp) {
    cudaFuncSetAttribute(pair_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pair_chain_kernel<true><<<grid, CH_THREADS, smem, st>>>(pair_r, num_unique, max_pairs, P, out_f, out_fp, ldA, ldB);
  } else {
    cudaFuncSetAttribute(pair_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pair_chain_kernel<false><<<grid, CH_THREADS, smem, st>>>(pair_r, num_unique, max_pairs, P, out_f, nullptr, ldA, ldB);
  }
  return tmd_check_launch("pair_chain");
}

int tmd_mp_pair_bwd(const float* YI, const float* YA, const float* YS, const float* dMI, const float* dMA,
                    const float* dMS, const float* f, const float* fp, const int* pid, const int* nbr_j,
                    const int* indptr, int n, int F, float* dYI, float* dYA, float* dYS, float* d_pair_r, void* stream) {
  TMD_REQUIRE(F > 0 && F % 4 == 0, "mp_pair_bwd: F must be a positive multiple of 4");
  TMD_REQUIRE(YI && YA && YS && dMI && dMA && dMS && f && pid && nbr_j && indptr && dYI && dYA && dYS, "mp_pair_bwd: null pointer");
  TMD_REQUIRE(!fp || d_pair_r, "mp_pair_bwd: fp given without d_pair_r");
  if (n == 0) return TMD_OK;
  const int chunks = (F + 127) / 128;
  const size_t smem = (size_t)8 * 18 * chunks * 32 * sizeof(float4);
  TMD_REQUIRE(smem <= 200 * 1024, "mp_pair_bwd: F too large for the per-warp staging buffer (F <= 256)");
  cudaFuncSetAttribute(mp_pair_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mp_pair_bwd_kernel<<<tmd_blocks_for((long long)n * 32, 256), 256, smem, (cudaStream_t)stream>>>(
      CIrr{YI, YA, YS}, CIrr{dMI, dMA, dMS}, f, fp, pid, nbr_j, indptr, n, F, MIrr{dYI, dYA, dYS}, d_pair_r);
  return tmd_check_launch("mp_pair_bwd");
}

// Single-pass backward of tmd_mp_fwd with pair-indexed weights: dY and, if fp != NULL, g_dist[k] = <f'_p, dM_j * Y_b> per
// directed slot k = (b, j) (the two slots of a pair together carry dE/dr_p; self-loop slots get 0).
int tmd_mp_edge_bwd(const float* YI, const float* YA, const float* YS, const float* dMI, const float* dMA,
                    const float* dMS, const float* f, const float* fp, const int* pid, const int* nbr_j,
                    const int* indptr, int n, int F, float* dYI, float* dYA, float* dYS, float* g_dist, void* stream) {
  TMD_REQUIRE(F > 0 && F % 4 == 0, "mp_edge_bwd: F must be a positive multiple of 4");
  TMD_REQUIRE(YI && YA && YS && dMI && dMA && dMS && f && pid && nbr_j && indptr && dYI && dYA && dYS, "mp_edge_bwd: null pointer");
  TMD_REQUIRE(!fp || g_dist, "mp_edge_bwd: fp given without g_dist");
  if (n == 0) return TMD_OK;
  const int chunks = (F + 127) / 128;
  const size_t smem = (size_t)8 * 9 * chunks * 32 * sizeof(float4);
  TMD_REQUIRE(smem <= 200 * 1024, "mp_edge_bwd: F too large for the per-warp staging buffer");
  cudaFuncSetAttribute(mp_edge_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mp_edge_bwd_kernel<<<tmd_blocks_for((long long)n * 32, 256), 256, smem, (cudaStream_t)stream>>>(
      CIrr{YI, YA, YS}, CIrr{dMI, dMA, dMS}, f, fp, pid, nbr_j, indptr, n, F, MIrr{dYI, dYA, dYS}, g_dist);
  return tmd_check_launch("mp_edge_bwd");
}

int tmd_embed_fwd(const float* vec, const float* dist, const float* dp, const int* pid, const int* nbr_j,
                  const int* indptr, const float* U, const float* V, const int64_t* zidx, double lo, double hi, int n, int F,
                  float* oI, float* oA, float* oS, void* stream) {
  TMD_REQUIRE(F > 0 && F % 4 == 0, "embed_fwd: F must be a positive multiple of 4");
  TMD_REQUIRE(vec && dist && dp && pid && nbr_j && indptr && U && V && oI && oA && oS, "embed_fwd: null pointer");
  if (n == 0) return TMD_OK;
  embed_fwd_kernel<<<tmd_blocks_for((long long)n * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      vec, dist, dp, pid, nbr_j, indptr, U, V, zidx, (float)lo, (float)hi, n, F, MIrr{oI, oA, oS});
  return tmd_check_launch("embed_fwd");
}
int tmd_embed_bwd(const float* vec, const float* dist, const float* dp, const float* dpp, const int* pid,
                  const int* nbr_j, const int* indptr, const float* U, const float* V, const int64_t* zidx, double lo, double hi,
                  int n, int F, const float* dI, const float* dA, const float* dS, float* g_dist, float* g_vec, void* stream) {
  TMD_REQUIRE(F > 0 && F % 4 == 0, "embed_bwd: F must be a positive multiple of 4");
  TMD_REQUIRE(vec && dist && dp && dpp && pid && nbr_j && indptr && U && V && dI && dA && dS && g_dist && g_vec, "embed_bwd: null pointer");
  if (n == 0) return TMD_OK;
  if (F <= 128)
    embed_bwd_kernel<true><<<tmd_blocks_for((long long)n * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        vec, dist, dp, dpp, pid, nbr_j, indptr, U, V, zidx, (float)lo, (float)hi, n, F, CIrr{dI, dA, dS}, g_dist, g_vec);
  else
    embed_bwd_kernel<false><<<tmd_blocks_for((long long)n * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        vec, dist, dp, dpp, pid, nbr_j, indptr, U, V, zidx, (float)lo, (float)hi, n, F, CIrr{dI, dA, dS}, g_dist, g_vec);
  return tmd_check_launch("embed_bwd");
}

}  // extern "C"
