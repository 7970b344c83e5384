// a5 (TensorEmbedding) by radial moments: the embedding's edge weights are LINEAR in the radial basis,
//     w_g[e, h] = C(r_e) (U[z_i] + V[z_j])[h] (sum_k W_g[h, k] rbf_k(r_e) + b_g[h]),      g = 1..3   (tensornet.py:554-585)
// so the sum over the neighbours of atom i can be taken BEFORE the linear map: with the per-edge angular factors
// phi_0 = 1 (I), phi_1..3 = skew(rhat) (A), phi_4..8 = sym-traceless(rhat rhat^T) (S), the per-atom, per-neighbour-element moments
//     m_i[v][k][c] = sum_{e in row i, z_j = v} C(r_e) phi_c(rhat_e) rho_k(r_e),        rho_k = rbf_k (k < R), rho_R = 1 (bias)
// give   X_i[c, h] = sum_v (U[z_i] + V[v])[h] sum_k W~_g(c)[h, k] m_i[v][k][c].
// Per edge this costs R exponentials and 9 (R+1) FMAs on 4-byte scalars instead of streaming a 3F-wide radial row per pair
// (1.5 KB, 3 KB in the backward) that a separate chain/table launch had to write first: the embedding leaves the per-pair data
// path altogether (C4: radial_table 166 us + embed_fwd 177 us + embed_bwd 413 us -> two launches of tens of microseconds).
// The backward runs the same factorisation in reverse: G_i[v][k][c] = sum_h dX_i[c, h] (U + V)[h] W~[h, k] per atom, then per
// edge dE/dr and dE/d rhat from <G, rho>, <G, rho'> and the derivatives of phi -- geometry gradients only (frozen weights).
// One warp per atom; everything that is indexed dynamically lives in a per-warp shared-memory tile.
#include "common.cuh"

namespace {

constexpr int CP = 12;   // 9 components padded to three float4

struct Geo { int v; float d, inv, r0, r1, r2, cut, dcut, u, c0, dc0; };

__device__ __forceinline__ Geo load_geo(const float* __restrict__ vec, const float* __restrict__ dist, const int* __restrict__ nbr_j,
                                        const int* __restrict__ zslot, int k, bool valid, float lo, float hi, float alpha) {
  Geo m{0, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (valid) {
    m.v = zslot[nbr_j[k]];
    const float d = __ldg(dist + k);
    m.d = d;
    m.inv = d == 0.f ? 1.f : 1.f / d;   // self loop: vec = 0 -> rhat = 0 (tensornet.py:363-366)
    m.r0 = __ldg(vec + 3 * (size_t)k) * m.inv; m.r1 = __ldg(vec + 3 * (size_t)k + 1) * m.inv; m.r2 = __ldg(vec + 3 * (size_t)k + 2) * m.inv;
    tn::cosine_cutoff(d, lo, hi, &m.cut, &m.dcut);
    tn::cosine_cutoff(d, 0.f, hi, &m.c0, &m.dc0);
    m.u = expf(alpha * (lo - d));
  }
  return m;
}
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 fma4s(float4 a, float s, float4 c) {
  return make_float4(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y), fmaf(a.z, s, c.z), fmaf(a.w, s, c.w));
}
__device__ __forceinline__ float4 fma4v(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 add4v(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 mul4v(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }

struct Out3 { float* I; float* A; float* S; };
struct In3 { const float* I; const float* A; const float* S; };

// ------------------------------------------------------------------------------------------------------------ forward
// Wk [3][R+1][F]: Wk[g][k][h] = W_g[h][k] (k < R), Wk[g][R][h] = b_g[h].  Uz, Vz [Z][F]: rows of the element tables for the Z
// species present; zslot[n]: species slot of every atom.  dynamic smem per warp: Z x (R+1) x 12 floats (moments) + 32 x 12 floats
// (per-edge factors of the current 32-edge chunk).
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <bool BIG>   // BIG: 32 < R <= 64 (two radial indices per lane)
__global__ void __launch_bounds__(256) embed_moments_fwd_kernel(const float* __restrict__ vec, const float* __restrict__ dist,
                                                                const int* __restrict__ nbr_j, const int* __restrict__ indptr,
                                                                const int* __restrict__ zslot, const float* __restrict__ Uz,
                                                                const float* __restrict__ Vz, const float* __restrict__ Wk,
                                                                const float* __restrict__ means, const float* __restrict__ betas,
                                                                float alpha, float lo, float hi, int R, int n, int F, int Z, Out3 out) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int RP = R + 1;
  const size_t per_warp = (size_t)Z * RP * CP + 32 * CP;
  float* mw = smem + (size_t)wid * per_warp;
  float* st = mw + (size_t)Z * RP * CP;          // [32 edges][12]: cut * phi_0..8, u, c0, -
  constexpr int NQ = BIG ? 2 : 1;
  float mu[NQ], be2[NQ];                          // be2 = -beta log2(e): exp(-beta t^2) = ex2(be2 t^2)
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int k = lane + 32 * q;
    mu[q] = k < R ? means[k] : 0.f;
    be2[q] = k < R ? -betas[k] * 1.4426950408889634f : 0.f;
  }
  for (int b = blockIdx.x * wpb + wid; b < n; b += gridDim.x * wpb) {
    const int k0 = indptr[b], k1 = indptr[b + 1];
    unsigned present = 0u;
    // ---- phase 1: moments.  Per 32-edge chunk the per-edge factors are prepared lane-parallel (lane <-> edge) and staged in
    // shared memory; the accumulation runs lane <-> radial index with register accumulators, one species after the other,
    // and touches the moment tile once per (chunk, species) ----
    for (int e0 = k0; e0 < k1; e0 += 32) {
      const bool valid = e0 + lane < k1;
      const Geo ge = load_geo(vec, dist, nbr_j, zslot, e0 + lane, valid, lo, hi, alpha);
      float c9[9];
      {
        const float r[3] = {ge.r0, ge.r1, ge.r2};
        float a[3], sb[5];
        tn::skew_compact(r, a); tn::sym_basis(r, sb);
        c9[0] = ge.cut;
#pragma unroll
        for (int c = 0; c < 3; ++c) c9[1 + c] = ge.cut * a[c];
#pragma unroll
        for (int c = 0; c < 5; ++c) c9[4 + c] = ge.cut * sb[c];
        float4* sp = reinterpret_cast<float4*>(st + (size_t)lane * CP);
        sp[0] = make_float4(c9[0], c9[1], c9[2], c9[3]); sp[1] = make_float4(c9[4], c9[5], c9[6], c9[7]);
        sp[2] = make_float4(c9[8], ge.u, ge.c0, 0.f);
      }
      __syncwarp();
      unsigned todo = __ballot_sync(0xffffffffu, valid);
      while (todo) {
        const int v = __shfl_sync(0xffffffffu, ge.v, __ffs(todo) - 1);
        const unsigned mask = __ballot_sync(0xffffffffu, valid && ge.v == v);
        todo &= ~mask;
        float acc[NQ][9];
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
          for (int c = 0; c < 9; ++c) acc[q][c] = 0.f;
        for (unsigned mm = mask; mm; mm &= mm - 1) {
          const int i = __ffs(mm) - 1;
          const float4 s0 = *reinterpret_cast<const float4*>(st + (size_t)i * CP), s1 = *reinterpret_cast<const float4*>(st + (size_t)i * CP + 4),
                       s2 = *reinterpret_cast<const float4*>(st + (size_t)i * CP + 8);
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const float t = s2.y - mu[q];
            const float val = s2.z * ex2_approx(be2[q] * t * t);
            acc[q][0] = fmaf(val, s0.x, acc[q][0]); acc[q][1] = fmaf(val, s0.y, acc[q][1]); acc[q][2] = fmaf(val, s0.z, acc[q][2]);
            acc[q][3] = fmaf(val, s0.w, acc[q][3]); acc[q][4] = fmaf(val, s1.x, acc[q][4]); acc[q][5] = fmaf(val, s1.y, acc[q][5]);
            acc[q][6] = fmaf(val, s1.z, acc[q][6]); acc[q][7] = fmaf(val, s1.w, acc[q][7]); acc[q][8] = fmaf(val, s2.x, acc[q][8]);
          }
        }
        // bias row (rho_R = 1): totals of cut * phi over the species' edges; lane c < 9 ends up with component c
        float bt[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) bt[c] = (c < 9 && ((mask >> lane) & 1u)) ? c9[c < 9 ? c : 0] : 0.f;
        const float btot = warp_reduce_transpose<16>(bt, lane);
        float* mv = mw + (size_t)v * RP * CP;
        const bool first = !((present >> v) & 1u);
        present |= 1u << v;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int k = lane + 32 * q;
          if (k < R) {
            float4* mp = reinterpret_cast<float4*>(mv + (size_t)k * CP);
            float4 x = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]), y = make_float4(acc[q][4], acc[q][5], acc[q][6], acc[q][7]);
            float z = acc[q][8];
            if (!first) {
              const float4 ox = mp[0], oy = mp[1];
              x = add4v(x, ox); y = add4v(y, oy); z += mv[(size_t)k * CP + 8];
            }
            mp[0] = x; mp[1] = y; mp[2] = make_float4(z, 0.f, 0.f, 0.f);
          }
        }
        if (lane < 9) mv[(size_t)R * CP + lane] = (first ? 0.f : mv[(size_t)R * CP + lane]) + btot;
      }
      __syncwarp();   // the staging rows are rewritten by the next chunk
    }
    __syncwarp();
    // ---- phase 2: contraction with the radial weights (lane <-> 4 channels) ----
    const int ub = zslot[b];
    for (int hh = 0; hh < F; hh += 128) {
      const int h = hh + lane * 4;
      const bool on = h < F;
      float4 X[9];
#pragma unroll
      for (int c = 0; c < 9; ++c) X[c] = zero4();
      const float4 uvec = on ? ldg4(Uz + (size_t)ub * F + h) : zero4();
      for (unsigned pm = present; pm; pm &= pm - 1) {
        const int v = __ffs(pm) - 1;
        const float* mv = mw + (size_t)v * RP * CP;
        float4 acc[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) acc[c] = zero4();
        if (on) {
#pragma unroll 2
          for (int k = 0; k < RP; ++k) {
            const float4 ma = *reinterpret_cast<const float4*>(mv + (size_t)k * CP);
            const float4 mb = *reinterpret_cast<const float4*>(mv + (size_t)k * CP + 4);
            const float mc = mv[(size_t)k * CP + 8];
            const float4 w1 = ldg4(Wk + ((size_t)0 * RP + k) * F + h);
            const float4 w2 = ldg4(Wk + ((size_t)1 * RP + k) * F + h);
            const float4 w3 = ldg4(Wk + ((size_t)2 * RP + k) * F + h);
            acc[0] = fma4s(w1, ma.x, acc[0]);
            acc[1] = fma4s(w2, ma.y, acc[1]); acc[2] = fma4s(w2, ma.z, acc[2]); acc[3] = fma4s(w2, ma.w, acc[3]);
            acc[4] = fma4s(w3, mb.x, acc[4]); acc[5] = fma4s(w3, mb.y, acc[5]); acc[6] = fma4s(w3, mb.z, acc[6]);
            acc[7] = fma4s(w3, mb.w, acc[7]); acc[8] = fma4s(w3, mc, acc[8]);
          }
          const float4 zc = add4v(uvec, ldg4(Vz + (size_t)v * F + h));
#pragma unroll
          for (int c = 0; c < 9; ++c) X[c] = fma4v(zc, acc[c], X[c]);
        }
      }
      if (on) {
        st4(out.I + (size_t)b * F + h, X[0]);
#pragma unroll
        for (int c = 0; c < 3; ++c) st4(out.A + ((size_t)b * 3 + c) * F + h, X[1 + c]);
#pragma unroll
        for (int c = 0; c < 5; ++c) st4(out.S + ((size_t)b * 5 + c) * F + h, X[4 + c]);
      }
    }
    __syncwarp();   // the tile is re-zeroed lazily by the next atom
  }
}

// ------------------------------------------------------------------------------------------------------------ backward
// Wh [3][F/4][R][4]: Wh[g][h/4][k][h%4] = W_g[h][k] (one float4 per (4 channels, k)); bias rows are read from Wk[g][R][:].  dynamic smem: 2 R floats (means, betas) then per
// warp Z x (R+1) x 12 floats (G) + 9 F floats (p = dX * (U + V)).
__global__ void __launch_bounds__(256) embed_moments_bwd_kernel(const float* __restrict__ vec, const float* __restrict__ dist,
                                                                const int* __restrict__ nbr_j, const int* __restrict__ indptr,
                                                                const int* __restrict__ zslot, const float* __restrict__ Uz,
                                                                const float* __restrict__ Vz, const float* __restrict__ Wk,
                                                                const float* __restrict__ Wh, const float* __restrict__ means,
                                                                const float* __restrict__ betas, float alpha, float lo, float hi, int R,
                                                                int n, int F, int Z, In3 dout, float* __restrict__ g_dist,
                                                                float* __restrict__ g_vec) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int RP = R + 1;
  float* s_mu = smem;
  float* s_be = smem + R;
  const size_t per_warp = (size_t)Z * RP * CP + (size_t)9 * F;
  float* Gw = smem + ((2 * R + 3) / 4) * 4 + (size_t)wid * per_warp;
  float* pw = Gw + (size_t)Z * RP * CP;
  for (int t = threadIdx.x; t < R; t += blockDim.x) { s_mu[t] = means[t]; s_be[t] = betas[t]; }
  __syncthreads();
  for (int b = blockIdx.x * wpb + wid; b < n; b += gridDim.x * wpb) {
    const int k0 = indptr[b], k1 = indptr[b + 1];
    // species present in this row
    unsigned present = 0u;
    for (int e = k0 + lane; e < k1; e += 32) present |= 1u << zslot[nbr_j[e]];
    present = __reduce_or_sync(0xffffffffu, present);
    const int ub = zslot[b];
    // ---- step A: G[v][k][c] = sum_h dX[c][h] (U + V_v)[h] W~_g(c)[h][k] ----
    for (unsigned pm = present; pm; pm &= pm - 1) {
      const int v = __ffs(pm) - 1;
      float bias_part[9];
#pragma unroll
      for (int c = 0; c < 9; ++c) bias_part[c] = 0.f;
      for (int hh = 0; hh < F; hh += 128) {
        const int h = hh + lane * 4;
        if (h < F) {
          const float4 zc = add4v(ldg4(Uz + (size_t)ub * F + h), ldg4(Vz + (size_t)v * F + h));
          float4 p[9];
          p[0] = mul4v(ldg4(dout.I + (size_t)b * F + h), zc);
#pragma unroll
          for (int c = 0; c < 3; ++c) p[1 + c] = mul4v(ldg4(dout.A + ((size_t)b * 3 + c) * F + h), zc);
#pragma unroll
          for (int c = 0; c < 5; ++c) p[4 + c] = mul4v(ldg4(dout.S + ((size_t)b * 5 + c) * F + h), zc);
#pragma unroll
          for (int c = 0; c < 9; ++c) *reinterpret_cast<float4*>(pw + (size_t)c * F + h) = p[c];
          const float4 b1 = ldg4(Wk + ((size_t)0 * RP + R) * F + h), b2 = ldg4(Wk + ((size_t)1 * RP + R) * F + h),
                       b3 = ldg4(Wk + ((size_t)2 * RP + R) * F + h);
          bias_part[0] += p[0].x * b1.x + p[0].y * b1.y + p[0].z * b1.z + p[0].w * b1.w;
#pragma unroll
          for (int c = 1; c < 4; ++c) bias_part[c] += p[c].x * b2.x + p[c].y * b2.y + p[c].z * b2.z + p[c].w * b2.w;
#pragma unroll
          for (int c = 4; c < 9; ++c) bias_part[c] += p[c].x * b3.x + p[c].y * b3.y + p[c].z * b3.z + p[c].w * b3.w;
        }
      }
#pragma unroll
      for (int c = 0; c < 9; ++c) bias_part[c] = warp_sum(bias_part[c]);
      float* Gv = Gw + (size_t)v * RP * CP;
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 9; ++c) Gv[(size_t)R * CP + c] = bias_part[c];
      }
      __syncwarp();
      for (int k = lane; k < R; k += 32) {     // lane <-> radial index
        float g[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) g[c] = 0.f;
        for (int h = 0; h < F; h += 4) {
          float4 p[9];
#pragma unroll
          for (int c = 0; c < 9; ++c) p[c] = *reinterpret_cast<const float4*>(pw + (size_t)c * F + h);   // broadcast
          const float pv[9][4] = {{p[0].x, p[0].y, p[0].z, p[0].w}, {p[1].x, p[1].y, p[1].z, p[1].w}, {p[2].x, p[2].y, p[2].z, p[2].w},
                                  {p[3].x, p[3].y, p[3].z, p[3].w}, {p[4].x, p[4].y, p[4].z, p[4].w}, {p[5].x, p[5].y, p[5].z, p[5].w},
                                  {p[6].x, p[6].y, p[6].z, p[6].w}, {p[7].x, p[7].y, p[7].z, p[7].w}, {p[8].x, p[8].y, p[8].z, p[8].w}};
          const float4 w1 = ldg4(Wh + (((size_t)0 * F + h) / 4 * R + k) * 4), w2 = ldg4(Wh + (((size_t)1 * F + h) / 4 * R + k) * 4),
                       w3 = ldg4(Wh + (((size_t)2 * F + h) / 4 * R + k) * 4);
          const float wv[3][4] = {{w1.x, w1.y, w1.z, w1.w}, {w2.x, w2.y, w2.z, w2.w}, {w3.x, w3.y, w3.z, w3.w}};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            g[0] = fmaf(pv[0][q], wv[0][q], g[0]);
#pragma unroll
            for (int c = 1; c < 4; ++c) g[c] = fmaf(pv[c][q], wv[1][q], g[c]);
#pragma unroll
            for (int c = 4; c < 9; ++c) g[c] = fmaf(pv[c][q], wv[2][q], g[c]);
          }
        }
        float4* gp = reinterpret_cast<float4*>(Gv + (size_t)k * CP);
        gp[0] = make_float4(g[0], g[1], g[2], g[3]); gp[1] = make_float4(g[4], g[5], g[6], g[7]); gp[2] = make_float4(g[8], 0.f, 0.f, 0.f);
      }
      __syncwarp();
    }
    // ---- step B: per edge (lane <-> edge) ----
    for (int e0 = k0; e0 < k1; e0 += 32) {
      const int e = e0 + lane;
      const bool valid = e < k1;
      const Geo ge = load_geo(vec, dist, nbr_j, zslot, e, valid, lo, hi, alpha);
      if (valid) {
        const float* Gv = Gw + (size_t)ge.v * RP * CP;
        float P[9], Pd[9];
        {   // bias row: rho = 1, rho' = 0
          const float4 x = *reinterpret_cast<const float4*>(Gv + (size_t)R * CP), y = *reinterpret_cast<const float4*>(Gv + (size_t)R * CP + 4);
          P[0] = x.x; P[1] = x.y; P[2] = x.z; P[3] = x.w; P[4] = y.x; P[5] = y.y; P[6] = y.z; P[7] = y.w; P[8] = Gv[(size_t)R * CP + 8];
#pragma unroll
          for (int c = 0; c < 9; ++c) Pd[c] = 0.f;
        }
        const float two_alpha_u_c0 = 2.f * alpha * ge.u * ge.c0;
        for (int k = 0; k < R; ++k) {
          // rbf_k = c0 e, d rbf_k/dr = dc0 e + 2 alpha beta u t c0 e  with e = exp(-beta t^2), t = u - mu_k  (tn::expnorm_rbf)
          const float t = ge.u - s_mu[k], bk = s_be[k];
          const float ee = ex2_approx(-1.4426950408889634f * bk * t * t);
          const float rho = ge.c0 * ee, drho = ee * fmaf(two_alpha_u_c0 * bk, t, ge.dc0);
          const float4 x = *reinterpret_cast<const float4*>(Gv + (size_t)k * CP), y = *reinterpret_cast<const float4*>(Gv + (size_t)k * CP + 4);
          const float z = Gv[(size_t)k * CP + 8];
          const float gk[9] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w, z};
#pragma unroll
          for (int c = 0; c < 9; ++c) { P[c] = fmaf(gk[c], rho, P[c]); Pd[c] = fmaf(gk[c], drho, Pd[c]); }
        }
        const float r[3] = {ge.r0, ge.r1, ge.r2};
        float a[3], sb[5];
        tn::skew_compact(r, a); tn::sym_basis(r, sb);
        float s0 = P[0], s1 = Pd[0];
#pragma unroll
        for (int c = 0; c < 3; ++c) { s0 = fmaf(a[c], P[1 + c], s0); s1 = fmaf(a[c], Pd[1 + c], s1); }
#pragma unroll
        for (int c = 0; c < 5; ++c) { s0 = fmaf(sb[c], P[4 + c], s0); s1 = fmaf(sb[c], Pd[4 + c], s1); }
        const float gd = ge.dcut * s0 + ge.cut * s1;
        // d/d rhat: skew (A01 = -r2, A02 = r1, A12 = -r0) and the symmetric-traceless basis
        float gr[3] = {-ge.cut * P[3], ge.cut * P[2], -ge.cut * P[1]};
        tn::sym_basis_bwd_acc(r, &P[4], ge.cut, gr);
        // rhat = vec * inv (inv = 1/dist; 1 for the self loop, whose vec is 0): d/dvec = g_rhat * inv, d/ddist -= (g_rhat . rhat) * inv
        const bool self = (r[0] == 0.f && r[1] == 0.f && r[2] == 0.f) && ge.inv == 1.f;
        g_dist[e] = gd - (self ? 0.f : (gr[0] * r[0] + gr[1] * r[1] + gr[2] * r[2]) * ge.inv);
        g_vec[3 * (size_t)e] = gr[0] * ge.inv; g_vec[3 * (size_t)e + 1] = gr[1] * ge.inv; g_vec[3 * (size_t)e + 2] = gr[2] * ge.inv;
      }
    }
    __syncwarp();
  }
}

int pick_warps(size_t per_warp_bytes, size_t fixed_bytes) {
  int w = 8;
  while (w > 1 && fixed_bytes + (size_t)w * per_warp_bytes > 100 * 1024) w >>= 1;
  return w;
}

}  // namespace

extern "C" {

int tmd_embed_moments_fwd(const float* vec, const float* dist, const int* nbr_j, const int* indptr, const int* zslot, const float* Uz,
                          const float* Vz, const float* Wk, const float* means, const float* betas, double alpha, double lo, double hi,
                          int R, int n, int F, int Z, float* oI, float* oA, float* oS, void* stream) {
  TMD_REQUIRE(F > 0 && F % 4 == 0, "embed_moments_fwd: F must be a positive multiple of 4");
  TMD_REQUIRE(R > 0 && R <= 64, "embed_moments_fwd: 1 <= num_rbf <= 64");
  TMD_REQUIRE(Z > 0 && Z <= 32, "embed_moments_fwd: 1 <= number of species <= 32");
  TMD_REQUIRE(vec && dist && nbr_j && indptr && zslot && Uz && Vz && Wk && means && betas && oI && oA && oS, "embed_moments_fwd: null pointer");
  if (n == 0) return TMD_OK;
  const size_t per_warp = ((size_t)Z * (R + 1) * CP + 32 * CP) * sizeof(float);
  TMD_REQUIRE(per_warp <= 200 * 1024, "embed_moments_fwd: species x num_rbf tile exceeds shared memory");
  const int wpb = pick_warps(per_warp, 0);
  const size_t smem = (size_t)wpb * per_warp;
  const int blocks = tmd_blocks_for((long long)n * 32, wpb * 32);
  if (R > 32) {
    cudaFuncSetAttribute(embed_moments_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    embed_moments_fwd_kernel<true><<<blocks, wpb * 32, smem, (cudaStream_t)stream>>>(
        vec, dist, nbr_j, indptr, zslot, Uz, Vz, Wk, means, betas, (float)alpha, (float)lo, (float)hi, R, n, F, Z, Out3{oI, oA, oS});
  } else {
    cudaFuncSetAttribute(embed_moments_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    embed_moments_fwd_kernel<false><<<blocks, wpb * 32, smem, (cudaStream_t)stream>>>(
        vec, dist, nbr_j, indptr, zslot, Uz, Vz, Wk, means, betas, (float)alpha, (float)lo, (float)hi, R, n, F, Z, Out3{oI, oA, oS});
  }
  return tmd_check_launch("embed_moments_fwd");
}

int tmd_embed_moments_bwd(const float* vec, const float* dist, const int* nbr_j, const int* indptr, const int* zslot, const float* Uz,
                          const float* Vz, const float* Wk, const float* Wh, const float* means, const float* betas, double alpha,
                          double lo, double hi, int R, int n, int F, int Z, const float* dI, const float* dA, const float* dS,
                          float* g_dist, float* g_vec, void* stream) {
  TMD_REQUIRE(F > 0 && F % 4 == 0, "embed_moments_bwd: F must be a positive multiple of 4");
  TMD_REQUIRE(R > 0 && R <= 64, "embed_moments_bwd: 1 <= num_rbf <= 64");
  TMD_REQUIRE(Z > 0 && Z <= 32, "embed_moments_bwd: 1 <= number of species <= 32");
  TMD_REQUIRE(vec && dist && nbr_j && indptr && zslot && Uz && Vz && Wk && Wh && means && betas && dI && dA && dS && g_dist && g_vec,
              "embed_moments_bwd: null pointer");
  if (n == 0) return TMD_OK;
  const size_t fixed = (size_t)((2 * R + 3) / 4) * 4 * sizeof(float);
  const size_t per_warp = ((size_t)Z * (R + 1) * CP + (size_t)9 * F) * sizeof(float);
  TMD_REQUIRE(fixed + per_warp <= 200 * 1024, "embed_moments_bwd: species x num_rbf tile exceeds shared memory");
  const int wpb = pick_warps(per_warp, fixed);
  const size_t smem = fixed + (size_t)wpb * per_warp;
  cudaFuncSetAttribute(embed_moments_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  embed_moments_bwd_kernel<<<tmd_blocks_for((long long)n * 32, wpb * 32), wpb * 32, smem, (cudaStream_t)stream>>>(
      vec, dist, nbr_j, indptr, zslot, Uz, Vz, Wk, Wh, means, betas, (float)alpha, (float)lo, (float)hi, R, n, F, Z, In3{dI, dA, dS},
      g_dist, g_vec);
  return tmd_check_launch("embed_moments_bwd");
}

}  // extern "C"
