// a5/a7: edge -> atom segmented reductions of TensorNet.
//
// Replaces torchmdnet/extensions/warp_kernels/tensornet_radial_mp.py:43-189 and
// tensornet_mp.py:42-131 (one thread per (atom, channel), scalar loads, fp atomics in the
// radial backward).  Here: one warp per atom, lanes stride the channel axis with 128-bit
// loads (F-last layouts make every edge row one coalesced burst), the per-edge scalar
// gradient is reduced with warp shuffles and written once -- no atomics anywhere, results
// are bit-reproducible.
//
// Layouts (fp32): edge_attr [E,3,F]; I [N,1,F], A [N,3,F], S [N,5,F] with the compact
// component order of tn_math.h; CSR: indptr[N+1], optional edge-id indirection `data`
// (nullptr = slot k is edge k, the layout tmd_neighbor_csr emits) and neighbour index `indices`.
#include "common.cuh"

namespace {

struct Irreps {  // compact node tensor: component k of atom n at p[k-group] + n*stride + k*F + h
  float* I; float* A; float* S; long long sI, sA, sS;
};
struct CIrreps {
  const float* I; const float* A; const float* S; long long sI, sA, sS;
};

__device__ __forceinline__ float4 f4_fma(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4_axpy(float s, float4 a, float4 c) {
  return make_float4(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z), fmaf(s, a.w, c.w));
}
__device__ __forceinline__ float f4_dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

__device__ __forceinline__ void load9(const CIrreps& T, long long n, int F, int h, float4* c) {
  c[0] = ldg4(T.I + n * T.sI + h);
#pragma unroll
  for (int k = 0; k < 3; ++k) c[1 + k] = ldg4(T.A + n * T.sA + (long long)k * F + h);
#pragma unroll
  for (int k = 0; k < 5; ++k) c[4 + k] = ldg4(T.S + n * T.sS + (long long)k * F + h);
}
__device__ __forceinline__ void store9(const Irreps& T, long long n, int F, int h, const float4* c) {
  st4(T.I + n * T.sI + h, c[0]);
#pragma unroll
  for (int k = 0; k < 3; ++k) st4(T.A + n * T.sA + (long long)k * F + h, c[1 + k]);
#pragma unroll
  for (int k = 0; k < 5; ++k) st4(T.S + n * T.sS + (long long)k * F + h, c[4 + k]);
}

// ---------------------------------------------------------------------------------------------
// embedding ("radial") message passing:  I_b = sum w1,  A_b = skew(sum w2 rhat),  S_b = sum w3 s(rhat)
// reference: models/tensornet.py:405-445 (PyTorch convention for the sign of A)
__global__ void __launch_bounds__(256) radial_mp_fwd_kernel(const float* __restrict__ rhat, const float* __restrict__ w,
                                                            const int* __restrict__ data, const int* __restrict__ indptr,
                                                            int n, int F, Irreps out) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < n; b += gridDim.x * wpb) {
    const int k0 = indptr[b], k1 = indptr[b + 1];
    for (int h = lane * 4; h < F; h += 128) {
      float4 acc[9];
#pragma unroll
      for (int c = 0; c < 9; ++c) acc[c] = f4_zero();
      for (int k = k0; k < k1; ++k) {
        const long long e = data ? data[k] : k;
        const float r[3] = {__ldg(rhat + 3 * e), __ldg(rhat + 3 * e + 1), __ldg(rhat + 3 * e + 2)};
        const float* we = w + e * 3 * (long long)F + h;
        const float4 w1 = ldg4(we), w2 = ldg4(we + F), w3 = ldg4(we + 2 * F);
        float a[3], s[5];
        tn::skew_compact(r, a); tn::sym_basis(r, s);
        acc[0] = make_float4(acc[0].x + w1.x, acc[0].y + w1.y, acc[0].z + w1.z, acc[0].w + w1.w);
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[1 + c] = f4_axpy(a[c], w2, acc[1 + c]);
#pragma unroll
        for (int c = 0; c < 5; ++c) acc[4 + c] = f4_axpy(s[c], w3, acc[4 + c]);
      }
      CIrreps dummy; (void)dummy;
      store9(out, b, F, h, acc);
    }
  }
}

// backward: dw[e,:,h] and d rhat[e] (reduced over all channels with shuffles; one writer per edge)
__global__ void __launch_bounds__(256) radial_mp_bwd_kernel(const float* __restrict__ rhat, const float* __restrict__ w,
                                                            const int* __restrict__ data, const int* __restrict__ indptr,
                                                            int n, int F, CIrreps dout, float* __restrict__ dw,
                                                            float* __restrict__ drhat) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < n; b += gridDim.x * wpb) {
    const int k0 = indptr[b], k1 = indptr[b + 1];
    for (int k = k0; k < k1; ++k) {
      const long long e = data ? data[k] : k;
      const float r[3] = {__ldg(rhat + 3 * e), __ldg(rhat + 3 * e + 1), __ldg(rhat + 3 * e + 2)};
      float a[3], s[5];
      tn::skew_compact(r, a); tn::sym_basis(r, s);
      float gr[3] = {0.f, 0.f, 0.f};
      for (int h = lane * 4; h < F; h += 128) {
        float4 d[9];
        load9(dout, b, F, h, d);
        const float* we = w + e * 3 * (long long)F + h;
        const float4 w2 = ldg4(we + F), w3 = ldg4(we + 2 * F);
        float4 g2 = f4_zero(), g3 = f4_zero();
#pragma unroll
        for (int c = 0; c < 3; ++c) g2 = f4_axpy(a[c], d[1 + c], g2);
#pragma unroll
        for (int c = 0; c < 5; ++c) g3 = f4_axpy(s[c], d[4 + c], g3);
        float* de = dw + e * 3 * (long long)F + h;
        st4(de, d[0]); st4(de + F, g2); st4(de + 2 * F, g3);
        // d/d rhat: A part  (A01=-vz, A02=+vy, A12=-vx with v = sum w2 rhat)
        gr[0] -= f4_dot(w2, d[3]); gr[1] += f4_dot(w2, d[2]); gr[2] -= f4_dot(w2, d[1]);
        // S part, channel by channel
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
      gr[0] = warp_sum(gr[0]); gr[1] = warp_sum(gr[1]); gr[2] = warp_sum(gr[2]);
      if (lane == 0) { drhat[3 * e] = gr[0]; drhat[3 * e + 1] = gr[1]; drhat[3 * e + 2] = gr[2]; }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// interaction message passing: out[b,comp,h] = sum_{k in row b} w[e_k, type(comp), h] * T[j_k, comp, h]
// reference: models/tensornet.py:622-679
__global__ void __launch_bounds__(256) mp_fwd_kernel(CIrreps T, const float* __restrict__ w, const int* __restrict__ data,
                                                     const int* __restrict__ indices, const int* __restrict__ indptr,
                                                     int n, int F, Irreps out) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < n; b += gridDim.x * wpb) {
    const int k0 = indptr[b], k1 = indptr[b + 1];
    for (int h = lane * 4; h < F; h += 128) {
      float4 acc[9];
#pragma unroll
      for (int c = 0; c < 9; ++c) acc[c] = f4_zero();
#pragma unroll 2
      for (int k = k0; k < k1; ++k) {
        const long long e = data ? data[k] : k;
        const long long j = indices[k];
        const float* we = w + e * 3 * (long long)F + h;
        const float4 w1 = ldg4(we), w2 = ldg4(we + F), w3 = ldg4(we + 2 * F);
        float4 t[9];
        load9(T, j, F, h, t);
        acc[0] = f4_fma(w1, t[0], acc[0]);
#pragma unroll
        for (int c = 1; c < 4; ++c) acc[c] = f4_fma(w2, t[c], acc[c]);
#pragma unroll
        for (int c = 4; c < 9; ++c) acc[c] = f4_fma(w3, t[c], acc[c]);
      }
      store9(out, b, F, h, acc);
    }
  }
}

// backward over the transposed structure (CSC of the forward's rows):
//   dT[b] = sum_{k in col b} w[e_k] * dOut[i_k],   dw[e_k, type, h] = sum_{comp in type} dOut[i_k,comp,h] * T[b,comp,h]
// every edge is written by exactly one warp -> no atomics (same property as tensornet_mp.py:80-131)
__global__ void __launch_bounds__(256) mp_bwd_kernel(CIrreps T, const float* __restrict__ w, CIrreps dout,
                                                     const int* __restrict__ data, const int* __restrict__ indices,
                                                     const int* __restrict__ indptr, int n, int F, Irreps dT,
                                                     float* __restrict__ dw) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < n; b += gridDim.x * wpb) {
    const int k0 = indptr[b], k1 = indptr[b + 1];
    for (int h = lane * 4; h < F; h += 128) {
      float4 acc[9], tb[9];
#pragma unroll
      for (int c = 0; c < 9; ++c) acc[c] = f4_zero();
      load9(T, b, F, h, tb);
#pragma unroll 2
      for (int k = k0; k < k1; ++k) {
        const long long e = data ? data[k] : k;
        const long long i = indices[k];
        const float* we = w + e * 3 * (long long)F + h;
        const float4 w1 = ldg4(we), w2 = ldg4(we + F), w3 = ldg4(we + 2 * F);
        float4 d[9];
        load9(dout, i, F, h, d);
        acc[0] = f4_fma(w1, d[0], acc[0]);
        float4 g2 = f4_zero(), g3 = f4_zero();
#pragma unroll
        for (int c = 1; c < 4; ++c) { acc[c] = f4_fma(w2, d[c], acc[c]); g2 = f4_fma(d[c], tb[c], g2); }
#pragma unroll
        for (int c = 4; c < 9; ++c) { acc[c] = f4_fma(w3, d[c], acc[c]); g3 = f4_fma(d[c], tb[c], g3); }
        if (dw) {
          float* de = dw + e * 3 * (long long)F + h;
          st4(de, f4_mul(d[0], tb[0])); st4(de + F, g2); st4(de + 2 * F, g3);
        }
      }
      store9(dT, b, F, h, acc);
    }
  }
}

}  // namespace

static inline bool f_ok(int F) { return F > 0 && (F % 4) == 0; }

extern "C" {

int tmd_radial_mp_fwd(const float* rhat, const float* edge_attr, const int* row_data, const int* row_indptr, int n,
                      int F, float* out_I, float* out_A, float* out_S, void* stream) {
  TMD_REQUIRE(f_ok(F), "radial_mp_fwd: F must be a positive multiple of 4");
  TMD_REQUIRE(row_indptr && out_I && out_A && out_S, "radial_mp_fwd: null pointer");
  if (n == 0) return TMD_OK;
  Irreps o{out_I, out_A, out_S, F, 3LL * F, 5LL * F};
  radial_mp_fwd_kernel<<<tmd_blocks_for((long long)n * 32, 256), 256, 0, (cudaStream_t)stream>>>(rhat, edge_attr, row_data, row_indptr, n, F, o);
  return tmd_check_launch("radial_mp_fwd");
}
int tmd_radial_mp_bwd(const float* rhat, const float* edge_attr, const int* row_data, const int* row_indptr, int n,
                      int F, const float* dI, const float* dA, const float* dS, float* d_edge_attr, float* d_rhat,
                      void* stream) {
  TMD_REQUIRE(f_ok(F), "radial_mp_bwd: F must be a positive multiple of 4");
  TMD_REQUIRE(row_indptr && dI && dA && dS && d_edge_attr && d_rhat, "radial_mp_bwd: null pointer");
  if (n == 0) return TMD_OK;
  CIrreps d{dI, dA, dS, F, 3LL * F, 5LL * F};
  radial_mp_bwd_kernel<<<tmd_blocks_for((long long)n * 32, 256), 256, 0, (cudaStream_t)stream>>>(rhat, edge_attr, row_data, row_indptr, n, F, d, d_edge_attr, d_rhat);
  return tmd_check_launch("radial_mp_bwd");
}
int tmd_mp_fwd(const float* I, const float* A, const float* S, const float* edge_attr, const int* row_data,
               const int* row_indices, const int* row_indptr, int n, int F, float* out_I, float* out_A, float* out_S,
               void* stream) {
  TMD_REQUIRE(f_ok(F), "mp_fwd: F must be a positive multiple of 4");
  TMD_REQUIRE(I && A && S && row_indices && row_indptr && out_I && out_A && out_S, "mp_fwd: null pointer");
  if (n == 0) return TMD_OK;
  CIrreps t{I, A, S, F, 3LL * F, 5LL * F};
  Irreps o{out_I, out_A, out_S, F, 3LL * F, 5LL * F};
  mp_fwd_kernel<<<tmd_blocks_for((long long)n * 32, 256), 256, 0, (cudaStream_t)stream>>>(t, edge_attr, row_data, row_indices, row_indptr, n, F, o);
  return tmd_check_launch("mp_fwd");
}
int tmd_mp_bwd(const float* I, const float* A, const float* S, const float* edge_attr, const float* dout_I,
               const float* dout_A, const float* dout_S, const int* col_data, const int* col_indices,
               const int* col_indptr, int n, int F, float* dI, float* dA, float* dS, float* d_edge_attr,
               void* stream) {
  TMD_REQUIRE(f_ok(F), "mp_bwd: F must be a positive multiple of 4");
  TMD_REQUIRE(I && A && S && dout_I && dout_A && dout_S && col_indices && col_indptr && dI && dA && dS, "mp_bwd: null pointer");
  if (n == 0) return TMD_OK;
  CIrreps t{I, A, S, F, 3LL * F, 5LL * F}, d{dout_I, dout_A, dout_S, F, 3LL * F, 5LL * F};
  Irreps o{dI, dA, dS, F, 3LL * F, 5LL * F};
  mp_bwd_kernel<<<tmd_blocks_for((long long)n * 32, 256), 256, 0, (cudaStream_t)stream>>>(t, edge_attr, d, col_data, col_indices, col_indptr, n, F, o, d_edge_attr);
  return tmd_check_launch("mp_bwd");
}

}  // extern "C"
