// a14: Equivariant-Transformer attention message (gather -> per-edge message -> segmented sum), forward and backward.
//
// Reference: EquivariantMultiHeadAttention.propagate/message/aggregate, torchmdnet/models/torchmd_et.py:355-426 --
// four index_select gathers, ~10 elementwise launches over [E,H,3hd] / [E,3,H,hd] tensors and two scatter-adds per
// layer (and the same again under autograd).  Here one warp owns one target atom t of the sorted symmetric CSR and
// streams its neighbours s; lanes stride the F channels with 128-bit loads (4 channels = a slice of one head), the
// per-head dot product q.k.dk is reduced with hd/4 shuffles, nothing of size [E, .] is ever materialised:
//   A_h      = sum_d q_t[h,d] k_s[h,d] dk_p[h,d]            attn_h = SiLU(A_h) * Ca
//   w_x[h,d] = v_s[h,x,d] dv_p[h,x,d] * Cv,  x in {a,b,c}   (Ca, Cv) = (C(r),1) or (1,C(r)) with vector_cutoff
//   x_t     += w_a * attn_h ;   vec_t[c] += vec_s[c] * w_b + w_c * d_c ,   d = (pos_s - pos_t)/r = -rhat(t,s)
// dk, dv are SiLU(Linear(rbf(r))): functions of r only, evaluated once per unordered pair (with d/dr) by the tcgen05
// radial chain (pair_chain_tc.cu) as one [R -> F+3F] layer; columns [0,F) = dk, [F,4F) = dv (per head: a|b|c).
// Backward: a target-centric kernel (dq, per-edge dE/dr and dE/d rhat) and a source-centric kernel (dk, dv, dvec) that
// walks the same rows using the symmetry of the list -- no atomics, deterministic.
#include "common.cuh"

namespace {

__device__ __forceinline__ float4 f4z() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4fma(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 f4axpy(float s, float4 a, float4 c) {
  return make_float4(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z), fmaf(s, a.w, c.w));
}
__device__ __forceinline__ float f4sum(float4 a) { return (a.x + a.y) + (a.z + a.w); }
__device__ __forceinline__ float f4dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// sum over the lanes that share one head (lanes_per_head = hd/4, a power of two <= 32)
__device__ __forceinline__ float head_sum(float v, int lanes_per_head) {
  for (int o = 1; o < lanes_per_head; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct EtGeom { float cut, dcut, ca, cv; };
__device__ __forceinline__ EtGeom et_geom(float r, float lo, float hi, int vcut) {
  EtGeom g;
  tn::cosine_cutoff(r, lo, hi, &g.cut, &g.dcut);
  g.ca = vcut ? 1.f : g.cut;
  g.cv = vcut ? g.cut : 1.f;
  return g;
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(256) et_msg_fwd_kernel(
    const float* __restrict__ q, const float* __restrict__ kk, const float* __restrict__ v, const float* __restrict__ vec,
    const float* __restrict__ dkv, const int* __restrict__ pid, const int* __restrict__ nbr_j,
    const int* __restrict__ indptr, const float* __restrict__ dist, const float* __restrict__ rhat, float lo, float hi,
    int vcut, int n, int F, int hd, float* __restrict__ xagg, float* __restrict__ vagg) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int f = lane * 4;
  const bool on = f < F;
  const int h = on ? f / hd : 0, d = on ? f - h * hd : 0;
  const int lph = hd >> 2;
  const long long vo = (long long)h * 3 * hd + d;  // offset of part a inside a [3F] row; parts b, c follow at +hd, +2hd
  for (int t = blockIdx.x * wpb + (threadIdx.x >> 5); t < n; t += gridDim.x * wpb) {
    const float4 qt = on ? ldg4(q + (long long)t * F + f) : f4z();
    float4 ax = f4z(), av0 = f4z(), av1 = f4z(), av2 = f4z();
    for (int k = indptr[t]; k < indptr[t + 1]; ++k) {
      const long long s = nbr_j[k], p = pid[k];
      const EtGeom g = et_geom(__ldg(dist + k), lo, hi, vcut);
      const float u0 = __ldg(rhat + 3 * k), u1 = __ldg(rhat + 3 * k + 1), u2 = __ldg(rhat + 3 * k + 2);
      float part = 0.f;
      float4 wa = f4z(), wb = f4z(), wc = f4z(), s0 = f4z(), s1 = f4z(), s2 = f4z();
      if (on) {
        const float* pr = dkv + p * 4 * (long long)F;
        part = f4dot(f4mul(qt, ldg4(kk + s * F + f)), ldg4(pr + f));
        const float* vs = v + s * 3 * (long long)F + vo;
        const float* dv = pr + F + vo;
        wa = f4scale(f4mul(ldg4(vs), ldg4(dv)), g.cv);
        wb = f4scale(f4mul(ldg4(vs + hd), ldg4(dv + hd)), g.cv);
        wc = f4scale(f4mul(ldg4(vs + 2 * hd), ldg4(dv + 2 * hd)), g.cv);
        const float* vv = vec + s * 3 * (long long)F + f;
        s0 = ldg4(vv); s1 = ldg4(vv + F); s2 = ldg4(vv + 2 * F);
      }
      const float A = head_sum(part, lph);
      const float attn = tn::silu(A) * g.ca;
      ax = f4axpy(attn, wa, ax);
      av0 = f4axpy(-u0, wc, f4fma(s0, wb, av0));
      av1 = f4axpy(-u1, wc, f4fma(s1, wb, av1));
      av2 = f4axpy(-u2, wc, f4fma(s2, wb, av2));
    }
    if (on) {
      st4(xagg + (long long)t * F + f, ax);
      float* o = vagg + (long long)t * 3 * F + f;
      st4(o, av0); st4(o + F, av1); st4(o + 2 * F, av2);
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward, target-centric
// dq_t, and for every slot k = (t, s): g_dist[k], g_rhat[k,3]
__global__ void __launch_bounds__(256) et_msg_bwd_target_kernel(
    const float* __restrict__ q, const float* __restrict__ kk, const float* __restrict__ v, const float* __restrict__ vec,
    const float* __restrict__ dkv, const float* __restrict__ dkvp, const float* __restrict__ gx,
    const float* __restrict__ gv, const int* __restrict__ pid, const int* __restrict__ nbr_j,
    const int* __restrict__ indptr, const float* __restrict__ dist, const float* __restrict__ rhat, float lo, float hi,
    int vcut, int n, int F, int hd, float* __restrict__ gq, float* __restrict__ g_dist, float* __restrict__ g_rhat) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int f = lane * 4;
  const bool on = f < F;
  const int h = on ? f / hd : 0, d = on ? f - h * hd : 0;
  const int lph = hd >> 2;
  const long long vo = (long long)h * 3 * hd + d;
  for (int t = blockIdx.x * wpb + (threadIdx.x >> 5); t < n; t += gridDim.x * wpb) {
    float4 qt = f4z(), gxt = f4z(), g0 = f4z(), g1 = f4z(), g2 = f4z(), aq = f4z();
    if (on) {
      qt = ldg4(q + (long long)t * F + f); gxt = ldg4(gx + (long long)t * F + f);
      const float* gp = gv + (long long)t * 3 * F + f;
      g0 = ldg4(gp); g1 = ldg4(gp + F); g2 = ldg4(gp + 2 * F);
    }
    for (int k = indptr[t]; k < indptr[t + 1]; ++k) {
      const long long s = nbr_j[k], p = pid[k];
      const EtGeom g = et_geom(__ldg(dist + k), lo, hi, vcut);
      const float u0 = __ldg(rhat + 3 * k), u1 = __ldg(rhat + 3 * k + 1), u2 = __ldg(rhat + 3 * k + 2);
      float part = 0.f, gattn_p = 0.f;
      float4 ks = f4z(), dk = f4z(), va = f4z(), vb = f4z(), vc = f4z(), da = f4z(), db = f4z(), dc = f4z();
      float4 gwb = f4z(), gwc = f4z();
      const float* pr = dkv + p * 4 * (long long)F;
      if (on) {
        ks = ldg4(kk + s * F + f); dk = ldg4(pr + f);
        part = f4dot(f4mul(qt, ks), dk);
        const float* vs = v + s * 3 * (long long)F + vo;
        const float* dv = pr + F + vo;
        va = ldg4(vs); vb = ldg4(vs + hd); vc = ldg4(vs + 2 * hd);
        da = ldg4(dv); db = ldg4(dv + hd); dc = ldg4(dv + 2 * hd);
        gattn_p = g.cv * f4dot(gxt, f4mul(va, da));                       // sum_d gx_t w_a
        const float* vv = vec + s * 3 * (long long)F + f;
        gwb = f4fma(g0, ldg4(vv), f4fma(g1, ldg4(vv + F), f4mul(g2, ldg4(vv + 2 * F))));  // sum_c gv_t[c] vec_s[c]
        gwc = f4axpy(-u0, g0, f4axpy(-u1, g1, f4scale(g2, -u2)));                        // sum_c gv_t[c] d_c
      }
      const float A = head_sum(part, lph);
      const float g_attn = head_sum(gattn_p, lph);
      const float sA = tn::silu(A), attn = sA * g.ca;
      const float gA = g_attn * tn::silu_grad(A) * g.ca;
      float gr = 0.f, gc = 0.f, gu0 = 0.f, gu1 = 0.f, gu2 = 0.f;
      if (on) {
        aq = f4axpy(gA, f4mul(ks, dk), aq);                               // dq_t += gA k_s dk
        const float* prp = dkvp + p * 4 * (long long)F;
        // d/dr through dk:  g_dk = gA q_t k_s
        gr = gA * f4dot(f4mul(qt, ks), ldg4(prp + f));
        // d/dr through dv:  g_dv_x = g_wx v_x Cv
        const float4 gwa = f4scale(gxt, attn);
        const float* dvp = prp + F + vo;
        gr += g.cv * (f4dot(f4mul(gwa, va), ldg4(dvp)) + f4dot(f4mul(gwb, vb), ldg4(dvp + hd)) + f4dot(f4mul(gwc, vc), ldg4(dvp + 2 * hd)));
        // through the cutoff
        if (vcut) gc = f4dot(gwa, f4mul(va, da)) + f4dot(gwb, f4mul(vb, db)) + f4dot(gwc, f4mul(vc, dc));
        // d/d rhat: vec message uses d = -rhat:  g_u_c = -sum gv_t[c] w_c
        const float4 wc = f4scale(f4mul(vc, dc), g.cv);
        gu0 = -f4dot(g0, wc); gu1 = -f4dot(g1, wc); gu2 = -f4dot(g2, wc);
      }
      if (!vcut && (lane % lph) == 0 && on) gc = g_attn * sA;             // one lane per head carries d attn / d Ca
      gr = warp_sum(gr + gc * g.dcut);
      gu0 = warp_sum(gu0); gu1 = warp_sum(gu1); gu2 = warp_sum(gu2);
      if (lane == 0) { g_dist[k] = gr; g_rhat[3 * k] = gu0; g_rhat[3 * k + 1] = gu1; g_rhat[3 * k + 2] = gu2; }
    }
    if (on) st4(gq + (long long)t * F + f, aq);
  }
}

// ------------------------------------------------------------------------------------------------ backward, source-centric
// dk_a, dv_a, dvec_a: atom a is the SOURCE of the edges whose targets are its neighbours b (list symmetric)
__global__ void __launch_bounds__(256) et_msg_bwd_source_kernel(
    const float* __restrict__ q, const float* __restrict__ kk, const float* __restrict__ v, const float* __restrict__ vec,
    const float* __restrict__ dkv, const float* __restrict__ gx, const float* __restrict__ gv,
    const int* __restrict__ pid, const int* __restrict__ nbr_j, const int* __restrict__ indptr,
    const float* __restrict__ dist, const float* __restrict__ rhat, float lo, float hi, int vcut, int n, int F, int hd,
    float* __restrict__ gk, float* __restrict__ gvv, float* __restrict__ gvec) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int f = lane * 4;
  const bool on = f < F;
  const int h = on ? f / hd : 0, d = on ? f - h * hd : 0;
  const int lph = hd >> 2;
  const long long vo = (long long)h * 3 * hd + d;
  for (int a = blockIdx.x * wpb + (threadIdx.x >> 5); a < n; a += gridDim.x * wpb) {
    float4 ka = f4z(), va = f4z(), vb = f4z(), vc = f4z(), s0 = f4z(), s1 = f4z(), s2 = f4z();
    float4 ak = f4z(), aa = f4z(), ab = f4z(), ac = f4z(), a0 = f4z(), a1 = f4z(), a2 = f4z();
    if (on) {
      ka = ldg4(kk + (long long)a * F + f);
      const float* vs = v + (long long)a * 3 * F + vo;
      va = ldg4(vs); vb = ldg4(vs + hd); vc = ldg4(vs + 2 * hd);
      const float* vv = vec + (long long)a * 3 * F + f;
      s0 = ldg4(vv); s1 = ldg4(vv + F); s2 = ldg4(vv + 2 * F);
    }
    for (int k = indptr[a]; k < indptr[a + 1]; ++k) {
      const long long b = nbr_j[k], p = pid[k];              // edge (target b, source a); same pair, same r
      const EtGeom g = et_geom(__ldg(dist + k), lo, hi, vcut);
      // rhat of slot (b, a) is -rhat of slot (a, b); the message uses d = -rhat(b, a) = +rhat(a, b)
      const float e0 = __ldg(rhat + 3 * k), e1 = __ldg(rhat + 3 * k + 1), e2 = __ldg(rhat + 3 * k + 2);
      float part = 0.f, gattn_p = 0.f;
      float4 qb = f4z(), dk = f4z(), da = f4z(), db = f4z(), dc = f4z(), gxb = f4z(), g0 = f4z(), g1 = f4z(), g2 = f4z();
      if (on) {
        const float* pr = dkv + p * 4 * (long long)F;
        qb = ldg4(q + b * F + f); dk = ldg4(pr + f);
        part = f4dot(f4mul(qb, ka), dk);
        const float* dv = pr + F + vo;
        da = ldg4(dv); db = ldg4(dv + hd); dc = ldg4(dv + 2 * hd);
        gxb = ldg4(gx + b * F + f);
        gattn_p = g.cv * f4dot(gxb, f4mul(va, da));
        const float* gp = gv + b * 3 * (long long)F + f;
        g0 = ldg4(gp); g1 = ldg4(gp + F); g2 = ldg4(gp + 2 * F);
      }
      const float A = head_sum(part, lph);
      const float g_attn = head_sum(gattn_p, lph);
      const float attn = tn::silu(A) * g.ca;
      const float gA = g_attn * tn::silu_grad(A) * g.ca;
      if (on) {
        ak = f4axpy(gA, f4mul(qb, dk), ak);                                   // dk_a += gA q_b dk
        aa = f4axpy(attn * g.cv, f4mul(gxb, da), aa);                         // dv_a (part a) += gx_b attn dv_a Cv
        const float4 gwb = f4fma(g0, s0, f4fma(g1, s1, f4mul(g2, s2)));
        const float4 gwc = f4axpy(e0, g0, f4axpy(e1, g1, f4scale(g2, e2)));
        ab = f4axpy(g.cv, f4mul(gwb, db), ab);
        ac = f4axpy(g.cv, f4mul(gwc, dc), ac);
        const float4 wb = f4scale(f4mul(vb, db), g.cv);                       // dvec_a[c] += gv_b[c] w_b
        a0 = f4fma(g0, wb, a0); a1 = f4fma(g1, wb, a1); a2 = f4fma(g2, wb, a2);
      }
    }
    if (on) {
      st4(gk + (long long)a * F + f, ak);
      float* o = gvv + (long long)a * 3 * F + vo;
      st4(o, aa); st4(o + hd, ab); st4(o + 2 * hd, ac);
      float* ov = gvec + (long long)a * 3 * F + f;
      st4(ov, a0); st4(ov + F, a1); st4(ov + 2 * F, a2);
    }
  }
}

inline bool et_dims_ok(int F, int hd) {
  if (F <= 0 || F > 128 || F % 4 || hd <= 0 || hd % 4 || F % hd) return false;
  const int lph = hd / 4;
  return (lph & (lph - 1)) == 0 && lph <= 32;
}

// ------------------------------------------------------------------------------------------------ a13: NeighborEmbedding
// x_nb[i] = sum_{e = (i, j), j != i} W_p(e) * T[z_j]        (reference models/utils.py:83-117: boolean-mask removal of the self
// loops, Linear(rbf) * C(r) per edge, embedding gather, index_add).  W_p = C(r) (W_d rbf(r) + b) is a function of r only: it
// comes per unordered pair from the radial chain (value + d/dr, tables for frozen weights); T is the neighbour embedding table
// [max_z, F] (64 KB, L1-resident).  One warp per atom, lanes over channels; backward w.r.t. geometry: g_dist[k] = <g_i, W'_p T[z_j]>.
__global__ void __launch_bounds__(256) nbr_embed_fwd_kernel(const float* __restrict__ wp, const float* __restrict__ table,
                                                            const int64_t* __restrict__ z, const int* __restrict__ pid,
                                                            const int* __restrict__ nbr_j, const int* __restrict__ indptr, int n, int F,
                                                            float* __restrict__ out) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    const int k0 = indptr[i], k1 = indptr[i + 1];
    for (int ff = 0; ff < F; ff += 128) {      // all lanes iterate (the shuffles need the whole warp); lanes past F idle
      const int f = ff + lane * 4;
      const bool on = f < F;
      float4 acc = f4z();
      for (int e0 = k0; e0 < k1; e0 += 32) {
        const int k = e0 + lane;
        int j = -1, p = 0, zj = 0;
        if (k < k1) { j = nbr_j[k]; p = pid[k]; zj = (int)z[j]; }
        const int m = min(32, k1 - e0);
        for (int t = 0; t < m; ++t) {
          const int jj = __shfl_sync(0xffffffffu, j, t), pp = __shfl_sync(0xffffffffu, p, t), zz = __shfl_sync(0xffffffffu, zj, t);
          if (jj == i || !on) continue;
          acc = f4fma(ldg4(wp + (long long)pp * F + f), ldg4(table + (long long)zz * F + f), acc);
        }
      }
      if (on) st4(out + (long long)i * F + f, acc);
    }
  }
}
__global__ void __launch_bounds__(256) nbr_embed_bwd_kernel(const float* __restrict__ wpp, const float* __restrict__ table,
                                                            const int64_t* __restrict__ z, const int* __restrict__ pid,
                                                            const int* __restrict__ nbr_j, const int* __restrict__ indptr, int n, int F,
                                                            const float* __restrict__ gout, float* __restrict__ g_dist) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    const int k0 = indptr[i], k1 = indptr[i + 1];
    for (int e0 = k0; e0 < k1; e0 += 32) {
      const int k = e0 + lane;
      int j = -1, p = 0, zj = 0;
      if (k < k1) { j = nbr_j[k]; p = pid[k]; zj = (int)z[j]; }
      const int m = min(32, k1 - e0);
      float keep = 0.f;
      for (int t = 0; t < m; ++t) {
        const int jj = __shfl_sync(0xffffffffu, j, t), pp = __shfl_sync(0xffffffffu, p, t), zz = __shfl_sync(0xffffffffu, zj, t);
        float s = 0.f;
        if (jj != i)
          for (int f = lane * 4; f < F; f += 128)
            s += f4dot(ldg4(gout + (long long)i * F + f), f4mul(ldg4(wpp + (long long)pp * F + f), ldg4(table + (long long)zz * F + f)));
        s = warp_sum(s);
        if (lane == t) keep = s;
      }
      if (lane < m) g_dist[e0 + lane] = keep;
    }
  }
}

}  // namespace

extern "C" {

// q, kk [N,F]; v [N,3F] (per head: a|b|c of hd each); vec [N,3,F]; dkv [P,4F] = (dk [F] | dv [3F]) per pair;
// CSR (indptr, nbr_j) with pid (slot -> pair), dist[E], rhat[E,3] (unit vector of the slot's own edge, 0 for self loops).
// F <= 128, hd % 4 == 0, hd/4 a power of two.
int tmd_et_msg_fwd(const float* q, const float* kk, const float* v, const float* vec, const float* dkv, const int* pid,
                   const int* nbr_j, const int* indptr, const float* dist, const float* rhat, double lo, double hi,
                   int vector_cutoff, int n, int F, int head_dim, float* xagg, float* vagg, void* stream) {
  TMD_REQUIRE(et_dims_ok(F, head_dim), "et_msg_fwd: need F <= 128, F % 4 == 0, head_dim % 4 == 0, head_dim/4 a power of two");
  TMD_REQUIRE(q && kk && v && vec && dkv && pid && nbr_j && indptr && dist && rhat && xagg && vagg, "et_msg_fwd: null pointer");
  if (n == 0) return TMD_OK;
  et_msg_fwd_kernel<<<tmd_blocks_for((long long)n * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      q, kk, v, vec, dkv, pid, nbr_j, indptr, dist, rhat, (float)lo, (float)hi, vector_cutoff, n, F, head_dim, xagg, vagg);
  return tmd_check_launch("et_msg_fwd");
}

// backward: gq, gk [N,F], gv_out [N,3F], gvec [N,3,F], g_dist[E], g_rhat[E,3] from gx [N,F], gv [N,3,F];
// dkvp = d(dkv)/dr per pair.
int tmd_et_msg_bwd(const float* q, const float* kk, const float* v, const float* vec, const float* dkv,
                   const float* dkvp, const float* gx, const float* gv, const int* pid, const int* nbr_j,
                   const int* indptr, const float* dist, const float* rhat, double lo, double hi, int vector_cutoff, int n,
                   int F, int head_dim, float* gq, float* gk, float* gv_out, float* gvec, float* g_dist, float* g_rhat,
                   void* stream) {
  TMD_REQUIRE(et_dims_ok(F, head_dim), "et_msg_bwd: need F <= 128, F % 4 == 0, head_dim % 4 == 0, head_dim/4 a power of two");
  TMD_REQUIRE(q && kk && v && vec && dkv && dkvp && gx && gv && pid && nbr_j && indptr && dist && rhat && gq && gk && gv_out && gvec && g_dist && g_rhat,
              "et_msg_bwd: null pointer");
  if (n == 0) return TMD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = tmd_blocks_for((long long)n * 32, 256);
  et_msg_bwd_target_kernel<<<grid, 256, 0, st>>>(q, kk, v, vec, dkv, dkvp, gx, gv, pid, nbr_j, indptr, dist, rhat, (float)lo,
                                                 (float)hi, vector_cutoff, n, F, head_dim, gq, g_dist, g_rhat);
  et_msg_bwd_source_kernel<<<grid, 256, 0, st>>>(q, kk, v, vec, dkv, gx, gv, pid, nbr_j, indptr, dist, rhat, (float)lo,
                                                 (float)hi, vector_cutoff, n, F, head_dim, gk, gv_out, gvec);
  return tmd_check_launch("et_msg_bwd");
}

// a13 (see the kernels): out [n,F] = neighbour embedding sum; wp / wpp [P_u,F] = C(r)(W_d rbf + b) per unordered pair and its d/dr;
// table [max_z,F] = NeighborEmbedding.embedding.weight; z int64 atomic numbers.  F % 4 == 0.
int tmd_nbr_embed_fwd(const float* wp, const float* table, const int64_t* z, const int* pid, const int* nbr_j, const int* indptr,
                      int n, int F, float* out, void* stream) {
  TMD_REQUIRE(F > 0 && F % 4 == 0, "nbr_embed_fwd: F must be a positive multiple of 4");
  TMD_REQUIRE(wp && table && z && pid && nbr_j && indptr && out, "nbr_embed_fwd: null pointer");
  if (n == 0) return TMD_OK;
  nbr_embed_fwd_kernel<<<tmd_blocks_for((long long)n * 32, 256), 256, 0, (cudaStream_t)stream>>>(wp, table, z, pid, nbr_j, indptr, n, F, out);
  return tmd_check_launch("nbr_embed_fwd");
}
int tmd_nbr_embed_bwd(const float* wpp, const float* table, const int64_t* z, const int* pid, const int* nbr_j, const int* indptr,
                      int n, int F, const float* gout, float* g_dist, void* stream) {
  TMD_REQUIRE(F > 0 && F % 4 == 0, "nbr_embed_bwd: F must be a positive multiple of 4");
  TMD_REQUIRE(wpp && table && z && pid && nbr_j && indptr && gout && g_dist, "nbr_embed_bwd: null pointer");
  if (n == 0) return TMD_OK;
  nbr_embed_bwd_kernel<<<tmd_blocks_for((long long)n * 32, 256), 256, 0, (cudaStream_t)stream>>>(wpp, table, z, pid, nbr_j, indptr, n, F, gout, g_dist);
  return tmd_check_launch("nbr_embed_bwd");
}

}  // extern "C"
