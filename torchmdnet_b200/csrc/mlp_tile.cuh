// Tile GEMM of the per-atom MLP kernels (head.cu: TensorNet tail + Scalar head; atom_mlp.cu: the embedding's scalar MLP).
//
// out[32 atoms, NC] += in_s[32, K] . W[K, NC] with fp32 operands and fp32-parity results: every operand is split on the fly into
// two TF32 terms (hi = cvt.rna.tf32(x), lo = cvt.rna.tf32(x - hi)) and each product is three warp-level tensor-core MMAs
// (lo.hi + hi.lo + hi.hi, small terms first) accumulated in fp32 -- the same 3xTF32 scheme as the tcgen05 kernels of the radial
// chain and the node linears.  These layers are ~1-3 GFLOP per 10k-atom frame (< 1 % of the frame's flops) on
// 32-atom tiles whose activations never leave shared memory between the layers; what limits them is instruction count and
// latency per tile, not tensor throughput, so the operands stay where the surrounding element-wise stages (norms, LayerNorm,
// SiLU) put them and the MMA is the register-operand `mma.sync.m16n8k8` form rather than a TMEM-accumulator pipeline: the
// FFMA version of this loop issued 21 instructions per 16 FMAs, this one ~60 per 12 x 1024.
// W is streamed through shared memory in chunks of KC rows with a three-stage cp.async pipeline (every CTA reads each weight row
// once, coalesced, two chunks ahead of its use); the staged rows are padded by 8 floats so that the B-fragment reads (4 rows x 8
// columns per instruction) hit 32 different banks; the activation tiles use a row stride = 4 (mod 32) floats for the same reason.
// Two details that the numbers depend on:
//  * every CTA walks the K chunks in a ROTATED order (start chunk = blockIdx mod #chunks).  All CTAs run the same layer at the
//    same time; in natural order the few hundred resident CTAs request the same 8 KB of W from the same L2 lines at the same
//    moment and the kernel runs at the pace of those lines (measured: 483 us for the embedding MLP, 10x its arithmetic time).
//  * the three MMAs of a k-step accumulate into a zeroed temporary that is then added to the running sum with an fp32 FADD.
//    The tensor core adds with truncation, so chaining all K/8 x 3 MMAs through one accumulator biases the result by
//    ~(number of MMAs) x 2^-24 (1.3e-5 relative at K = 576); with the temporary only the 8-term partial sums are truncated.
#pragma once
#include "common.cuh"

namespace mlp {

constexpr int TA = 32;       // atoms per CTA tile = two m16 MMA tiles
constexpr int HT = 256;      // threads per CTA (8 warps); warp w owns the n8 column tiles w, w + 8, w + 16, ...
constexpr int NW = HT / 32;
constexpr int KC = 16;       // weight rows per staged chunk = two MMA k-steps
constexpr int STAGES = 3;
constexpr int MAXNT = 6;     // column tiles per warp: NC <= 8 * NW * MAXNT = 384
constexpr int WPAD = 8;

__host__ __device__ inline size_t wbuf_floats(int max_nc) { return (size_t)STAGES * KC * (max_nc + WPAD); }

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
// KC rows x NC columns of W (row stride ldw) starting at row k0 -> wb [KC][NC + WPAD]
__device__ __forceinline__ void stage_w(float* wb, const float* __restrict__ W, int ldw, int k0, int NC, int tid) {
  const int nc4 = NC >> 2;
  for (int i = tid; i < KC * nc4; i += HT) {
    const int r = i / nc4, c = i - r * nc4;
    cp_async16(wb + (size_t)r * (NC + WPAD) + 4 * c, W + (size_t)(k0 + r) * ldw + 4 * c);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// acc[mt][nt][0..3]: rows 16 mt + g and 16 mt + g + 8 (g = lane / 4), columns 8 (warp + NW nt) + 2 (lane % 4) and + 1:
//   [0] = (g, 2t), [1] = (g, 2t + 1), [2] = (g + 8, 2t), [3] = (g + 8, 2t + 1).
// K % KC == 0 (16), NC % 8 == 0, NC <= 8 NW MAXNT, ld % 32 == 4.  All threads of the CTA must call (barriers inside).
__device__ __forceinline__ void tile_gemm_tc(const float* __restrict__ in_s, int ld, int K, const float* __restrict__ W, int ldw, int NC,
                                             float (&acc)[2][MAXNT][4], float* wbuf, int warp, int lane, int tid) {
  const int g = lane >> 2, t = lane & 3;
  const int ntiles = NC >> 3, wst = NC + WPAD;
  const int nchunks = K / KC, rot = blockIdx.x % nchunks;
  auto chunk_k0 = [&](int c) { int x = c + rot; if (x >= nchunks) x -= nchunks; return x * KC; };
  stage_w(wbuf, W, ldw, chunk_k0(0), NC, tid);
  if (nchunks > 1) stage_w(wbuf + (size_t)KC * wst, W, ldw, chunk_k0(1), NC, tid);
  else asm volatile("cp.async.commit_group;" ::: "memory");
  for (int c = 0; c < nchunks; ++c) {
    asm volatile("cp.async.wait_group 1;" ::: "memory");   // groups 0 .. c have landed (c + 1 may be in flight)
    __syncthreads();                                         // ... for every thread, and chunk c - 1 has been consumed by all
    if (c + 2 < nchunks) stage_w(wbuf + (size_t)((c + 2) % STAGES) * KC * wst, W, ldw, chunk_k0(c + 2), NC, tid);
    else asm volatile("cp.async.commit_group;" ::: "memory");   // keep one group per iteration
    const float* cur = wbuf + (size_t)(c % STAGES) * KC * wst;
    const int k0 = chunk_k0(c);
#pragma unroll
    for (int ks = 0; ks < KC; ks += 8) {
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const float* a = in_s + (size_t)(16 * mt + g) * ld + k0 + ks + t;
        split_tf32(a[0], ah[mt][0], al[mt][0]);
        split_tf32(a[(size_t)8 * ld], ah[mt][1], al[mt][1]);
        split_tf32(a[4], ah[mt][2], al[mt][2]);
        split_tf32(a[(size_t)8 * ld + 4], ah[mt][3], al[mt][3]);
      }
#pragma unroll
      for (int nt = 0; nt < MAXNT; ++nt) {
        const int tile = warp + NW * nt;
        if (tile < ntiles) {
          const float* b = cur + (size_t)(ks + t) * wst + 8 * tile + g;
          uint32_t bh0, bl0, bh1, bl1;
          split_tf32(b[0], bh0, bl0);
          split_tf32(b[(size_t)4 * wst], bh1, bl1);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            float tmp[4] = {0.f, 0.f, 0.f, 0.f};
            mma_tf32(tmp, al[mt], bh0, bh1);
            mma_tf32(tmp, ah[mt], bl0, bl1);
            mma_tf32(tmp, ah[mt], bh0, bh1);
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[mt][nt][r] += tmp[r];
          }
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();   // the staging buffers and in_s may be rewritten by the caller
}

// visit every accumulator element of this thread: fn(row in tile, column, value&)
template <typename Fn>
__device__ __forceinline__ void for_each_acc(float (&acc)[2][MAXNT][4], int NC, int warp, int lane, Fn fn) {
  const int g = lane >> 2, t = lane & 3, ntiles = NC >> 3;
#pragma unroll
  for (int nt = 0; nt < MAXNT; ++nt) {
    const int tile = warp + NW * nt;
    if (tile < ntiles) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int r = 0; r < 4; ++r) fn(16 * mt + g + 8 * (r >> 1), 8 * tile + 2 * t + (r & 1), acc[mt][nt][r]);
    }
  }
}
__device__ __forceinline__ void zero_acc(float (&acc)[2][MAXNT][4]) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < MAXNT; ++nt)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[mt][nt][r] = 0.f;
}

}  // namespace mlp
