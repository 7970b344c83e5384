// a6 (+ the radial projections of a5) on the 5th-generation tensor cores: tcgen05 / TMEM / TMA.
//
// The radial chain  rbf(r) -> [Linear -> SiLU] x L -> (x cutoff)  with its d/dr tangent is evaluated for every
// unordered atom pair as a sequence of GEMMs  D[2P, N] = A[2P, K] . W[N, K]^T  (row 2p = value of pair p, row 2p+1 =
// its tangent, so the activation derivative needs only a lane^1 shuffle in the epilogue).
//
// fp32 parity (energies/forces to 1e-4 relative, reference noise ~3e-6) rules out single-pass TF32 (1e-3).  Each
// operand is split into two TF32 terms, x = hi + lo (cvt.rna.tf32), and D accumulates lo.hi + hi.lo + hi.hi in the
// fp32 TMEM accumulator ("3xTF32", ~2^-21 relative): three tcgen05.mma.kind::tf32 per K-step at ~1/3 of the TF32
// peak, still ~5x the fp32 FFMA roofline.  The previous layer's epilogue writes its output already split (hi, lo),
// so operands go HBM -> TMA (SWIZZLE_128B) -> SMEM -> tensor core with no conversion pass.
//
// Kernel anatomy (one CTA per SM, persistent over (m_tile, n_block) work items, 192 threads):
//   warp 0      : TMA producer   - A_hi/A_lo [128 x 32] and W_hi/W_lo [BN x 32] fp32 boxes per K-block, mbarrier ring
//   warp 1      : MMA issuer     - allocates 512 TMEM columns, one elected lane issues 12 tcgen05.mma per K-block
//                                  (4 K-steps of 8 x 3 split terms), tcgen05.commit frees the SMEM stage / publishes D
//   warps 2..17 : epilogue       - tcgen05.ld 32x32b.x32 (thread = TMEM lane = output row), bias + SiLU + tangent
//                                  (+ cutoff on the last layer), re-split to (hi, lo) or final fp32 f / f' to HBM;
//                                  double-buffered accumulator (2 x 256 columns) overlaps it with the next item's MMAs
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 32;                 // 32 fp32 = 128 B = one SWIZZLE_128B row
constexpr int EPI_WARPS = 16;            // four warps per TMEM lane quarter, round-robin over 32-column chunks
constexpr int TC_THREADS = 64 + 32 * EPI_WARPS;
constexpr int MAX_N = 1024;              // bias staged in shared memory
constexpr int A_TILE_BYTES = BM * BK * 4;  // 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, kind::tf32, both operands K-major
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
// start address >> 4 | LBO (=1, ignored for swizzled K-major) << 16 | SBO (8 rows x 128 B = 1024 B) >> 4 << 32 |
// version 1 << 46 | layout SWIZZLE_128B (2) << 61
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 (1<<4), A=TF32 (2<<7), B=TF32 (2<<10), K-major both,
// N>>3 at bit 17, M>>4 at bit 24
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// same with A = B = F16 (format 0): kind::f16, K = 16 per instruction
__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 256-bit global store (STG.E.256 on sm_100a): one full 32-byte sector per lane and instruction
__device__ __forceinline__ void st8(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

struct GemmParams {
  int K;            // padded to a multiple of 32
  int N;            // output width, multiple of BN
  int BN;           // 64..256, multiple of 16
  int rows_per_pair;  // 2: (value, tangent) interleaved; 1: value only
  int act;          // SiLU
  int final_layer;  // write fp32 f / f' (and apply the cutoff if mul_cutoff) instead of the split activations
  int mul_cutoff;
  float lo, hi;
  const float* bias;     // [N]
  const float* pair_r;   // [P]
  const int* num_unique; // device counter
  int max_pairs;
  float* out_hi;   // hidden: [M_pad, N]
  float* out_lo;
  float* out_f;    // final: [P, N]
  float* out_fp;   // final, rows_per_pair == 2
  int f16;              // operands are fp16 (hi, lo) pairs (K-block = 64 elements) instead of tf32 (32 elements)
  int seg_t1, seg_t2;   // grouped GEMM: m-tiles [0,t1) use weight block 0, [t1,t2) block 1, [t2,..) block 2
  int b_seg_rows;       // rows of one weight block in the stacked B tensor map
};

template <int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm3x_tf32_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                   GemmParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-B aligned operand ring
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_tile_bytes = P.BN * BK * 4;
  const int stage_bytes = 2 * A_TILE_BYTES + 2 * b_tile_bytes;
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float bias_s[MAX_N];
  for (int i = threadIdx.x; i < P.N; i += TC_THREADS) bias_s[i] = P.bias[i];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_pairs = P.num_unique ? min(*P.num_unique, P.max_pairs) : P.max_pairs;
  const int rows_valid = n_pairs * P.rows_per_pair;
  const int m_tiles = (rows_valid + BM - 1) / BM;
  const int n_blks = P.N / P.BN;
  const int total_items = m_tiles * n_blks;
  const int bk_el = P.f16 ? 2 * BK : BK;   // elements per 128-byte K-block
  const int k_blocks = P.K / bk_el;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: all 512 columns (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int m_tile = item / n_blks, n_blk = item - m_tile * n_blks;
        const int b_row = n_blk * P.BN + ((m_tile >= P.seg_t1) + (m_tile >= P.seg_t2)) * P.b_seg_rows;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + (size_t)stage * stage_bytes;
          mbar_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
          tma_load_2d(&map_a_hi, &full_bar[stage], st, kb * bk_el, m_tile * BM);
          tma_load_2d(&map_a_lo, &full_bar[stage], st + A_TILE_BYTES, kb * bk_el, m_tile * BM);
          tma_load_2d(&map_b_hi, &full_bar[stage], st + 2 * A_TILE_BYTES, kb * bk_el, b_row);
          tma_load_2d(&map_b_lo, &full_bar[stage], st + 2 * A_TILE_BYTES + b_tile_bytes, kb * bk_el, b_row);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = P.f16 ? make_idesc_f16(BM, P.BN) : make_idesc_tf32(BM, P.BN);
    const bool f16 = P.f16 != 0;
    int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t st = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t a_hi = make_desc_sw128(st), a_lo = make_desc_sw128(st + A_TILE_BYTES);
          const uint64_t b_hi = make_desc_sw128(st + 2 * A_TILE_BYTES), b_lo = make_desc_sw128(st + 2 * A_TILE_BYTES + b_tile_bytes);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);  // 8 tf32 / 16 f16 = 32 B per K-step inside the 128 B swizzle row
            if (f16) {
              umma_f16(d_tmem, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
              umma_f16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
              umma_f16(d_tmem, a_hi + adv, b_hi + adv, idesc, 1);
            } else {
              umma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
              umma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
              umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, 1);
            }
          }
        }
        __syncwarp();
        if (elect_one()) umma_commit(&empty_bar[stage]);       // SMEM stage reusable when these MMAs retire
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tmem_full_bar[acc]);         // accumulator complete -> epilogue
      __syncwarp();
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..17; TMEM lane quarter = warp & 3)
    // jvp: lanes (2m, 2m+1) hold the value / tangent rows of pair m.  They swap halves of the 32-column chunk so that
    // each lane evaluates the activation ONCE for 16 columns and produces both h and dh for them.
    const int q = warp & 3;
    const int sub = (warp - 2) >> 2;  // which of the round-robin 32-column chunks this warp drains
    const int row_in_tile = q * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    const bool jvp = P.rows_per_pair == 2;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int m_tile = item / n_blks, n_blk = item - m_tile * n_blks;
      const int row = m_tile * BM + row_in_tile;
      const bool tangent = jvp && (row & 1);
      const int row_v = jvp ? (row & ~1) : row;          // the value row of this lane's pair
      const bool valid = row_v < rows_valid;
      const int pair = jvp ? (row >> 1) : row;
      float cut = 1.f, dcut = 0.f;
      if (P.final_layer && P.mul_cutoff) {
        const float r = valid ? __ldg(P.pair_r + pair) : 0.f;
        tn::cosine_cutoff(r, P.lo, P.hi, &cut, &dcut);
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
      for (int c0 = sub * 32; c0 < P.BN; c0 += 32 * (EPI_WARPS / 4)) {
        uint32_t v[32];
        tmem_ld32(t_row + (uint32_t)c0, v);
        const int col0 = n_blk * P.BN + c0;
        if (jvp) {
          const int cb = col0 + (tangent ? 16 : 0);        // this lane's 16 columns
          float h[16], dh[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float send = __uint_as_float(tangent ? v[i] : v[16 + i]);
            const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
            const float z = (tangent ? recv : __uint_as_float(v[i])) + bias_s[cb + i];
            const float t = tangent ? __uint_as_float(v[16 + i]) : recv;
            float hh = z, dd = t;
            if (P.act) {
              // sigmoid on the SFU (ex2 + rcp): ~1e-7 relative, far inside the 3xTF32 error budget
              const float sg = __fdividef(1.f, 1.f + __expf(-z));
              hh = z * sg;
              dd = sg * fmaf(z, 1.f - sg, 1.f) * t;
            }
            if (P.final_layer && P.mul_cutoff) { dd = fmaf(dd, cut, hh * dcut); hh *= cut; }
            h[i] = hh; dh[i] = dd;
          }
          if (valid) {
            if (P.final_layer) {
              float* d0 = P.out_f + (size_t)pair * P.N + cb;
              float* d1 = P.out_fp + (size_t)pair * P.N + cb;
              st8(d0, h); st8(d0 + 8, h + 8);
              st8(d1, dh); st8(d1 + 8, dh + 8);
            } else {
              const size_t o0 = (size_t)row_v * P.N + cb, o1 = o0 + P.N;
              if (P.f16) {
                // fp16 (hi, lo): 16 columns = 32 bytes = one 256-bit store per (row, array)
                __half* oh = reinterpret_cast<__half*>(P.out_hi);
                __half* ol = reinterpret_cast<__half*>(P.out_lo);
                float a[8], b[8], c[8], d[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  const __half2 hh = __floats2half2_rn(h[2 * u], h[2 * u + 1]);
                  const float2 hf = __half22float2(hh);
                  const __half2 hl = __floats2half2_rn(h[2 * u] - hf.x, h[2 * u + 1] - hf.y);
                  const __half2 dh2 = __floats2half2_rn(dh[2 * u], dh[2 * u + 1]);
                  const float2 df = __half22float2(dh2);
                  const __half2 dl = __floats2half2_rn(dh[2 * u] - df.x, dh[2 * u + 1] - df.y);
                  a[u] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&hh));
                  b[u] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&hl));
                  c[u] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&dh2));
                  d[u] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&dl));
                }
                st8(reinterpret_cast<float*>(oh + o0), a); st8(reinterpret_cast<float*>(ol + o0), b);
                st8(reinterpret_cast<float*>(oh + o1), c); st8(reinterpret_cast<float*>(ol + o1), d);
              } else
#pragma unroll
              for (int i = 0; i < 16; i += 8) {
                float a[8], b[8], c[8], d[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  a[u] = tf32_rna(h[i + u]);  b[u] = tf32_rna(h[i + u] - a[u]);
                  c[u] = tf32_rna(dh[i + u]); d[u] = tf32_rna(dh[i + u] - c[u]);
                }
                st8(P.out_hi + o0 + i, a); st8(P.out_lo + o0 + i, b);
                st8(P.out_hi + o1 + i, c); st8(P.out_lo + o1 + i, d);
              }
            }
          }
        } else {
          float o[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float z = __uint_as_float(v[i]) + bias_s[col0 + i];
            float res = P.act ? z * __fdividef(1.f, 1.f + __expf(-z)) : z;
            if (P.final_layer && P.mul_cutoff) res *= cut;
            o[i] = res;
          }
          if (valid) {
            if (P.final_layer) {
              float* dst = P.out_f + (size_t)pair * P.N + col0;
#pragma unroll
              for (int i = 0; i < 32; i += 4) st4(dst + i, make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]));
            } else {
              if (P.f16) {
                __half* oh = reinterpret_cast<__half*>(P.out_hi) + (size_t)row * P.N + col0;
                __half* ol = reinterpret_cast<__half*>(P.out_lo) + (size_t)row * P.N + col0;
#pragma unroll
                for (int i = 0; i < 32; i += 16) {
                  float a[8], b[8];
#pragma unroll
                  for (int u = 0; u < 8; ++u) {
                    const __half2 hh = __floats2half2_rn(o[i + 2 * u], o[i + 2 * u + 1]);
                    const float2 hf = __half22float2(hh);
                    const __half2 hl = __floats2half2_rn(o[i + 2 * u] - hf.x, o[i + 2 * u + 1] - hf.y);
                    a[u] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&hh));
                    b[u] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&hl));
                  }
                  st8(reinterpret_cast<float*>(oh + i), a); st8(reinterpret_cast<float*>(ol + i), b);
                }
              } else {
              float* dhp = P.out_hi + (size_t)row * P.N + col0;
              float* dlp = P.out_lo + (size_t)row * P.N + col0;
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                float hi4[4], lo4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { hi4[u] = tf32_rna(o[i + u]); lo4[u] = tf32_rna(o[i + u] - hi4[u]); }
                st4(dhp + i, make_float4(hi4[0], hi4[1], hi4[2], hi4[3]));
                st4(dlp + i, make_float4(lo4[0], lo4[1], lo4[2], lo4[3]));
              }
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// layer-0 operand: rows (2p, 2p+1) = (rbf(r_p), d rbf/dr) split into (hi, lo), K padded to K_pad with zeros
__global__ void pair_rbf_split_kernel(const float* __restrict__ pair_r, const int* __restrict__ num_unique, int max_pairs,
                                      int R, int K_pad, int rows_per_pair, const float* __restrict__ means,
                                      const float* __restrict__ betas, float alpha, float lo, float hi,
                                      float* __restrict__ a_hi, float* __restrict__ a_lo, int f16) {
  const int n_pairs = min(*num_unique, max_pairs);
  const long long total = (long long)n_pairs * K_pad;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(idx / K_pad), k = (int)(idx - (long long)p * K_pad);
    float v = 0.f, dv = 0.f;
    if (k < R) {
      const float r = pair_r[p];
      float c0, dc0;
      tn::cosine_cutoff(r, 0.f, hi, &c0, &dc0);
      const float u = expf(alpha * (lo - r));
      tn::expnorm_rbf(u, alpha, c0, dc0, means[k], betas[k], &v, &dv);
    }
    const size_t r0 = (size_t)p * rows_per_pair * K_pad + k;
    if (f16) {
      __half* hh = reinterpret_cast<__half*>(a_hi); __half* hl = reinterpret_cast<__half*>(a_lo);
      __half x = __float2half_rn(v);
      hh[r0] = x; hl[r0] = __float2half_rn(v - __half2float(x));
      if (rows_per_pair == 2) { x = __float2half_rn(dv); hh[r0 + K_pad] = x; hl[r0 + K_pad] = __float2half_rn(dv - __half2float(x)); }
      continue;
    }
    float h = tf32_rna(v);
    a_hi[r0] = h; a_lo[r0] = tf32_rna(v - h);
    if (rows_per_pair == 2) { h = tf32_rna(dv); a_hi[r0 + K_pad] = h; a_lo[r0 + K_pad] = tf32_rna(dv - h); }
  }
}

// (I [N,1,F], A [N,3,F], S [N,5,F]) -> TF32 (hi, lo) in one row-segmented buffer [R_tot, F]; segment starts 0, r1, r2
__global__ void irreps_split_kernel(const float* __restrict__ I, const float* __restrict__ A, const float* __restrict__ S,
                                    int N, int F, int r1, int r2, float* __restrict__ hi, float* __restrict__ lo) {
  const long long total = 9LL * N * F;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / F;
    const int c = (int)(idx - row * F);
    float x; long long orow;
    if (row < N) { x = I[row * F + c]; orow = row; }
    else if (row < 4LL * N) { x = A[(row - N) * F + c]; orow = r1 + (row - N); }
    else { x = S[(row - 4LL * N) * F + c]; orow = r2 + (row - 4LL * N); }
    const float h = tf32_rna(x);
    hi[orow * F + c] = h; lo[orow * F + c] = tf32_rna(x - h);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows, K] fp32 row-major, box = [box_rows, 32 cols], 128 B swizzle
static int make_map(CUtensorMap* m, const float* base, int rows, int K, int box_rows, int f16 = 0) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { tmd_set_error("cuTensorMapEncodeTiled unavailable"); return TMD_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)K * (f16 ? 2 : 4)};
  cuuint32_t box[2] = {(cuuint32_t)(f16 ? 2 * BK : BK), (cuuint32_t)box_rows};   // 128 bytes of K per row
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { tmd_set_error("cuTensorMapEncodeTiled failed"); return TMD_ERR_CUDA; }
  return TMD_OK;
}

static int pick_bn(int N) {
  // TMD_TC_BN (dev knob): preferred n-block; default = the widest of 256/192/128/64 that divides N
  static int pref = -1;
  if (pref < 0) { const char* e = getenv("TMD_TC_BN"); pref = e ? atoi(e) : 0; }
  if (pref > 0 && pref <= 256 && pref % 16 == 0 && N % pref == 0) return pref;
  const int cands[] = {256, 192, 128, 64};
  for (int c : cands) if (N % c == 0) return c;
  return 0;
}

}  // namespace

extern "C" {

// Tensor-core radial chain on pairs (same contract as tmd_pair_chain, 3xTF32 on tcgen05).
//   W_hi[l], W_lo[l] : TF32 split of the nn.Linear weight [N_l, K_l_pad] (row-major, K padded to a multiple of 32 with 0)
//   dims = {R, N_0, N_1, N_2}; hidden widths must be multiples of 32, every N_l a multiple of 64
//   rbf_hi/rbf_lo    : optional layer-0 operand from tmd_pair_rbf_split (shared by all chains of a frame), else NULL
//   act_a_hi/lo, act_b_hi/lo : ping-pong workspaces, each >= rows_pad * max(K_l_pad) floats where
//                              rows_pad = roundup(max_pairs * (out_fp ? 2 : 1), 128)
int tmd_pair_rbf_split(const float* pair_r, const int* num_unique, int max_pairs, int R, int rows_per_pair,
                       const float* means, const float* betas, double alpha, double lo, double hi, float* a_hi,
                       float* a_lo, int use_f16, void* stream) {
  TMD_REQUIRE(pair_r && num_unique && means && betas && a_hi && a_lo, "pair_rbf_split: null pointer");
  TMD_REQUIRE(rows_per_pair == 1 || rows_per_pair == 2, "pair_rbf_split: rows_per_pair must be 1 or 2");
  if (max_pairs == 0) return TMD_OK;
  const int bk = use_f16 ? 2 * BK : BK;
  const int K0 = ((R + bk - 1) / bk) * bk;
  const long long total = (long long)max_pairs * K0;
  pair_rbf_split_kernel<<<tmd_blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      pair_r, num_unique, max_pairs, R, K0, rows_per_pair, means, betas, (float)alpha, (float)lo, (float)hi, a_hi, a_lo, use_f16);
  return tmd_check_launch("pair_rbf_split");
}

int tmd_pair_chain_tc(const float* pair_r, const int* num_unique, int max_pairs, const float* const* W_hi,
                      const float* const* W_lo, const float* const* bias, const int* dims, int n_layers, int act,
                      int mul_cutoff, const float* means, const float* betas, double alpha, double lo, double hi,
                      const float* rbf_hi, const float* rbf_lo, float* act_a_hi, float* act_a_lo, float* act_b_hi,
                      float* act_b_lo, float* out_f, float* out_fp, int use_f16, void* stream) {
  TMD_REQUIRE(pair_r && num_unique && W_hi && W_lo && bias && dims && means && betas && out_f, "pair_chain_tc: null pointer");
  TMD_REQUIRE((rbf_hi == nullptr) == (rbf_lo == nullptr), "pair_chain_tc: rbf_hi / rbf_lo must be given together");
  TMD_REQUIRE(act_a_hi && act_a_lo && act_b_hi && act_b_lo, "pair_chain_tc: null workspace");
  TMD_REQUIRE(n_layers >= 1 && n_layers <= 3, "pair_chain_tc: n_layers must be 1..3");
  if (max_pairs == 0) return TMD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int rpp = out_fp ? 2 : 1;
  const int rows_pad = ((max_pairs * rpp + BM - 1) / BM) * BM;
  const int R = dims[0];
  const int bk = use_f16 ? 2 * BK : BK;
  const int K0 = ((R + bk - 1) / bk) * bk;
  if (!rbf_hi) {  // layer-0 operand not precomputed by tmd_pair_rbf_split: build it into workspace A
    const long long total = (long long)max_pairs * K0;
    pair_rbf_split_kernel<<<tmd_blocks_for(total, 256), 256, 0, st>>>(pair_r, num_unique, max_pairs, R, K0, rpp, means, betas,
                                                                      (float)alpha, (float)lo, (float)hi, act_a_hi, act_a_lo, use_f16);
  }
  const float* in_hi = rbf_hi ? rbf_hi : act_a_hi;
  const float* in_lo = rbf_hi ? rbf_lo : act_a_lo;
  float* o_hi = act_b_hi; float* o_lo = act_b_lo;
  float* n_hi = act_a_hi; float* n_lo = act_a_lo;   // where the layer after this one writes
  int K = K0;
  for (int l = 0; l < n_layers; ++l) {
    const int N = dims[l + 1];
    const int BN = pick_bn(N);
    TMD_REQUIRE(BN > 0, "pair_chain_tc: layer widths must be multiples of 64");
    TMD_REQUIRE(l == n_layers - 1 || N % bk == 0, "pair_chain_tc: hidden widths must be multiples of 32 (tf32) / 64 (f16)");
    TMD_REQUIRE(N <= MAX_N, "pair_chain_tc: layer wider than 1024");
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    int rc;
    if ((rc = make_map(&ma_hi, in_hi, rows_pad, K, BM, use_f16))) return rc;
    if ((rc = make_map(&ma_lo, in_lo, rows_pad, K, BM, use_f16))) return rc;
    if ((rc = make_map(&mb_hi, W_hi[l], N, K, BN, use_f16))) return rc;
    if ((rc = make_map(&mb_lo, W_lo[l], N, K, BN, use_f16))) return rc;
    GemmParams P;
    P.K = K; P.N = N; P.BN = BN; P.rows_per_pair = rpp; P.act = act; P.final_layer = (l == n_layers - 1);
    P.mul_cutoff = mul_cutoff; P.lo = (float)lo; P.hi = (float)hi; P.bias = bias[l]; P.pair_r = pair_r;
    P.num_unique = num_unique; P.max_pairs = max_pairs; P.out_hi = o_hi; P.out_lo = o_lo; P.out_f = out_f; P.out_fp = out_fp;
    P.seg_t1 = P.seg_t2 = 0x7fffffff; P.b_seg_rows = 0; P.f16 = use_f16;
    const int stage_bytes = 2 * A_TILE_BYTES + 2 * BN * BK * 4;
    const int max_items = (rows_pad / BM) * (N / BN);
    const int grid = max_items < 148 ? max_items : 148;
    if (stage_bytes * 3 + 1024 <= 200 * 1024) {
      const int smem = stage_bytes * 3 + 1024;
      cudaFuncSetAttribute(gemm3x_tf32_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      gemm3x_tf32_kernel<3><<<grid, TC_THREADS, smem, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, P);
    } else {
      const int smem = stage_bytes * 2 + 1024;
      cudaFuncSetAttribute(gemm3x_tf32_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      gemm3x_tf32_kernel<2><<<grid, TC_THREADS, smem, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, P);
    }
    // next layer reads what this one wrote
    in_hi = o_hi; in_lo = o_lo;
    float* t;
    t = o_hi; o_hi = n_hi; n_hi = t;
    t = o_lo; o_lo = n_lo; n_lo = t;
    K = N;
  }
  return tmd_check_launch("pair_chain_tc");
}

// a9: the three bias-free F x F node linears of TensorNet (I, A, S rows each with its own weight) as ONE grouped 3xTF32
// GEMM on tcgen05.  W_hi / W_lo: TF32 split of the stacked weights [3F, F] (rows 0..F-1 act on I, F..2F-1 on A, 2F..3F-1 on S).
// Row-segmented buffers of R_tot = r128(N) + r128(3N) + r128(5N) rows: ws_hi, ws_lo (operands) and out (fp32 result);
// the I / A / S blocks of `out` start at rows 0, r128(N), r128(N) + r128(3N).  zero_bias: F zeros.
int tmd_irreps_linear_tc(const float* I, const float* A, const float* S, int N, int F, const float* W_hi, const float* W_lo,
                         const float* zero_bias, float* ws_hi, float* ws_lo, float* out, void* stream) {
  TMD_REQUIRE(I && A && S && W_hi && W_lo && zero_bias && ws_hi && ws_lo && out, "irreps_linear_tc: null pointer");
  TMD_REQUIRE(F % 64 == 0 && F <= 256, "irreps_linear_tc: F must be a multiple of 64, <= 256");
  if (N == 0) return TMD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  auto r128 = [](long long x) { return (int)(((x + BM - 1) / BM) * BM); };
  const int r1 = r128(N), r2 = r1 + r128(3LL * N), rtot = r2 + r128(5LL * N);
  irreps_split_kernel<<<tmd_blocks_for(9LL * N * F, 256), 256, 0, st>>>(I, A, S, N, F, r1, r2, ws_hi, ws_lo);
  const int BN = pick_bn(F);
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  if ((rc = make_map(&ma_hi, ws_hi, rtot, F, BM))) return rc;
  if ((rc = make_map(&ma_lo, ws_lo, rtot, F, BM))) return rc;
  if ((rc = make_map(&mb_hi, W_hi, 3 * F, F, BN))) return rc;
  if ((rc = make_map(&mb_lo, W_lo, 3 * F, F, BN))) return rc;
  GemmParams P;
  P.K = F; P.N = F; P.BN = BN; P.rows_per_pair = 1; P.act = 0; P.final_layer = 1; P.mul_cutoff = 0; P.lo = 0.f; P.hi = 1.f;
  P.bias = zero_bias; P.pair_r = nullptr; P.num_unique = nullptr; P.max_pairs = rtot; P.out_hi = nullptr; P.out_lo = nullptr;
  P.out_f = out; P.out_fp = nullptr; P.seg_t1 = r1 / BM; P.seg_t2 = r2 / BM; P.b_seg_rows = F; P.f16 = 0;
  const int stage_bytes = 2 * A_TILE_BYTES + 2 * BN * BK * 4;
  const int items = (rtot / BM) * (F / BN);
  const int grid = items < 148 ? items : 148;
  if (stage_bytes * 3 + 1024 <= 200 * 1024) {
    const int smem = stage_bytes * 3 + 1024;
    cudaFuncSetAttribute(gemm3x_tf32_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    gemm3x_tf32_kernel<3><<<grid, TC_THREADS, smem, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, P);
  } else {
    const int smem = stage_bytes * 2 + 1024;
    cudaFuncSetAttribute(gemm3x_tf32_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    gemm3x_tf32_kernel<2><<<grid, TC_THREADS, smem, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, P);
  }
  return tmd_check_launch("irreps_linear_tc");
}

}  // extern "C"
