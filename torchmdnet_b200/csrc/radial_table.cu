// the radial function of a frozen model, r -> f(r) = chain(rbf(r)) (x cutoff) and f'(r), read from a per-model
// table of cubic Hermite pieces instead of being re-evaluated for every pair of every frame.  The input of the chain is ONE scalar
// per pair (reference models/tensornet.py:738-743), so for fixed weights the 3F outputs are smooth 1-D functions of r.
//
// Table layout [n_int][4][N] fp32, interval i = [i h, (i+1) h) with h = 2^-k:
//   (f_i,  D_i = f_{i+1} - f_i,  m_i = h f'(r_i),  m_{i+1} = h f'(r_{i+1}))       -- built on the host in fp64, rounded once
// h is a power of two, so the node positions i h and the local coordinate t = r / h - i are exact in fp32, and storing the
// difference D_i (not f_{i+1}) keeps the derivative free of cancellation noise:
//   f (r) = f_i + t^2 (3 - 2t) D_i + t (1 - t)^2 m_i + t^2 (t - 1) m_{i+1}
//   f'(r) = [6 t (1 - t) D_i + (1 - t)(1 - 3t) m_i + t (3t - 2) m_{i+1}] / h
// One thread per (pair, 4 columns): 64 B of table in, 32 B out; the table (16-63 MB) lives in L2, the kernel is bound by
// writing f, f' [P_u, N] to HBM.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) radial_table_kernel(const float* __restrict__ pair_r, const int* __restrict__ num_unique,
                                                           int max_pairs, const float* __restrict__ table, int n_int, float inv_h,
                                                           int N, float* __restrict__ out_f, float* __restrict__ out_fp) {
  const int n_pairs = min(*num_unique, max_pairs);
  const int q = N >> 2;
  const long long total = (long long)n_pairs * q;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(idx / q);
    const int c = (int)(idx - (long long)p * q) << 2;
    const float x = fmaxf(__ldg(pair_r + p), 0.f) * inv_h;   // exact: inv_h is a power of two
    const int i = min((int)x, n_int - 1);
    const float t = fminf(x - (float)i, 1.f);
    const float* row = table + (size_t)i * 4 * N + c;
    const float4 f0 = ldg4(row), D = ldg4(row + N), m0 = ldg4(row + 2 * N), m1 = ldg4(row + 3 * N);
    const float u = 1.f - t;
    const float a = t * t * (3.f - 2.f * t), b = t * u * u, g = -t * t * u;
    float4 o;
    o.x = fmaf(a, D.x, fmaf(b, m0.x, fmaf(g, m1.x, f0.x)));
    o.y = fmaf(a, D.y, fmaf(b, m0.y, fmaf(g, m1.y, f0.y)));
    o.z = fmaf(a, D.z, fmaf(b, m0.z, fmaf(g, m1.z, f0.z)));
    o.w = fmaf(a, D.w, fmaf(b, m0.w, fmaf(g, m1.w, f0.w)));
    st4(out_f + (size_t)p * N + c, o);
    if (out_fp) {
      const float da = 6.f * t * u * inv_h, db = u * (1.f - 3.f * t) * inv_h, dg = t * (3.f * t - 2.f) * inv_h;
      float4 d;
      d.x = fmaf(da, D.x, fmaf(db, m0.x, dg * m1.x));
      d.y = fmaf(da, D.y, fmaf(db, m0.y, dg * m1.y));
      d.z = fmaf(da, D.z, fmaf(db, m0.z, dg * m1.z));
      d.w = fmaf(da, D.w, fmaf(db, m0.w, dg * m1.w));
      st4(out_fp + (size_t)p * N + c, d);
    }
  }
}

}  // namespace

extern "C" {

int tmd_radial_table(const float* pair_r, const int* num_unique, int max_pairs, const float* table, int n_int, double h, int N,
                     float* out_f, float* out_fp, void* stream) {
  TMD_REQUIRE(pair_r && num_unique && table && out_f, "radial_table: null pointer");
  TMD_REQUIRE(n_int > 0 && h > 0 && N > 0 && N % 4 == 0, "radial_table: need n_int > 0, h > 0, N a positive multiple of 4");
  if (max_pairs == 0) return TMD_OK;
  radial_table_kernel<<<tmd_blocks_for((long long)max_pairs * (N / 4), 256), 256, 0, (cudaStream_t)stream>>>(
      pair_r, num_unique, max_pairs, table, n_int, (float)(1.0 / h), N, out_f, out_fp);
  return tmd_check_launch("radial_table");
}

}  // extern "C"
