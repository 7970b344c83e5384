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

// raw SFU ops (flush-to-zero, no range fix-ups): sigmoid(z) = rcp(1 + ex2(-z log2 e)); +-inf behave (sg -> 0 / 1)
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

struct JvpOut {
  float* out_f; float* out_fp; float* out_hi; float* out_lo;
  size_t pair, row_v; int N, cb; bool valid;
};

// (value, tangent) lane-pair exchange + bias + SiLU (+ cutoff): v = this lane's 32 accumulator columns; on return h / dh hold the
// activation and its tangent for this lane's 16 columns (value lane: columns 0..15 of the chunk, tangent lane: 16..31).
template <bool ACT, bool CUT>
__device__ __forceinline__ void jvp_math(const uint32_t* v, bool tangent, const float* bias16, float cut, float dcut, float* h, float* dh) {
  float bb[16];
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    const float4 b4 = *reinterpret_cast<const float4*>(bias16 + i);
    bb[i] = b4.x; bb[i + 1] = b4.y; bb[i + 2] = b4.z; bb[i + 3] = b4.w;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float send = __uint_as_float(tangent ? v[i] : v[16 + i]);
    const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
    const float z = (tangent ? recv : __uint_as_float(v[i])) + bb[i];
    const float t = tangent ? __uint_as_float(v[16 + i]) : recv;
    float hh = z, dd = t;
    if (ACT) {
      const float sg = rcp_ftz(1.f + ex2_ftz(-1.4426950408889634f * z));   // ~1e-7 relative, far inside the split-term budget
      hh = z * sg;
      dd = sg * fmaf(z, 1.f - sg, 1.f) * t;
    }
    if (CUT) { dd = fmaf(dd, cut, hh * dcut); hh *= cut; }
    h[i] = hh; dh[i] = dd;
  }
}

// One 32-column accumulator chunk of the (value, tangent) row pair held by lanes (2m, 2m+1): the lanes swap halves so that each
// evaluates the activation once for 16 columns and produces both h and dh for them.
template <bool ACT, bool FINAL, bool CUT, bool F16>
__device__ __forceinline__ void jvp_chunk(const uint32_t* v, bool tangent, const float* bias16, float cut, float dcut, const JvpOut& o) {
  float h[16], dh[16];
  jvp_math<ACT, FINAL && CUT>(v, tangent, bias16, cut, dcut, h, dh);
  if (!o.valid) return;
  if (FINAL) {
    float* d0 = o.out_f + o.pair * o.N + o.cb;
    float* d1 = o.out_fp + o.pair * o.N + o.cb;
    st8(d0, h); st8(d0 + 8, h + 8);
    st8(d1, dh); st8(d1 + 8, dh + 8);
    return;
  }
  const size_t o0 = o.row_v * o.N + o.cb, o1 = o0 + o.N;
  if (F16) {
    // fp16 (hi, lo): 16 columns = 32 bytes = one 256-bit store per (row, array)
    __half* oh = reinterpret_cast<__half*>(o.out_hi);
    __half* ol = reinterpret_cast<__half*>(o.out_lo);
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
  } else {
#pragma unroll
    for (int i = 0; i < 16; i += 8) {
      float a[8], b[8], c[8], d[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        a[u] = tf32_rna(h[i + u]);  b[u] = tf32_rna(h[i + u] - a[u]);
        c[u] = tf32_rna(dh[i + u]); d[u] = tf32_rna(dh[i + u] - c[u]);
      }
      st8(o.out_hi + o0 + i, a); st8(o.out_lo + o0 + i, b);
      st8(o.out_hi + o1 + i, c); st8(o.out_lo + o1 + i, d);
    }
  }
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
  int dbg;              // dev probe (TMD_TC_DBG): 1 epilogue off, 2 MMA off, 3 no global stores, 4 no A loads
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
          tma_load_2d(&map_a_hi, &full_bar[stage], st, kb * bk_el, P.dbg == 4 ? 0 : m_tile * BM);
          tma_load_2d(&map_a_lo, &full_bar[stage], st + A_TILE_BYTES, kb * bk_el, P.dbg == 4 ? 0 : m_tile * BM);
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
        if (P.dbg != 2 && elect_one()) {
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
      for (int c0 = sub * 32; c0 < P.BN && P.dbg != 1; c0 += 32 * (EPI_WARPS / 4)) {
        uint32_t v[32];
        tmem_ld32(t_row + (uint32_t)c0, v);
        const int col0 = n_blk * P.BN + c0;
        if (jvp) {
          const int cb = col0 + (tangent ? 16 : 0);        // this lane's 16 columns
          const JvpOut o{P.out_f, P.out_fp, P.out_hi, P.out_lo, (size_t)pair, (size_t)row_v, P.N, cb, valid && P.dbg != 3};
          const float* bs = bias_s + cb;
          if (P.final_layer) {
            if (P.mul_cutoff) { if (P.act) jvp_chunk<true, true, true, false>(v, tangent, bs, cut, dcut, o); else jvp_chunk<false, true, true, false>(v, tangent, bs, cut, dcut, o); }
            else              { if (P.act) jvp_chunk<true, true, false, false>(v, tangent, bs, cut, dcut, o); else jvp_chunk<false, true, false, false>(v, tangent, bs, cut, dcut, o); }
          } else if (P.f16)   { if (P.act) jvp_chunk<true, false, false, true>(v, tangent, bs, cut, dcut, o); else jvp_chunk<false, false, false, true>(v, tangent, bs, cut, dcut, o); }
          else                { if (P.act) jvp_chunk<true, false, false, false>(v, tangent, bs, cut, dcut, o); else jvp_chunk<false, false, false, false>(v, tangent, bs, cut, dcut, o); }
        } else {
          float o[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float z = __uint_as_float(v[i]) + bias_s[col0 + i];
            float res = P.act ? z * rcp_ftz(1.f + ex2_ftz(-1.4426950408889634f * z)) : z;
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

// ---------------------------------------------------------------------------------------------------------------------
// Fused tail of the radial chain: layers 2 and 3 in ONE kernel, the 2F-wide hidden activation h2 never leaves the SM.
//   h1 [rows, F] (fp16 hi/lo, written by the layer-1 launch of gemm3x)  --W2-->  h2 = SiLU(.)  --W3-->  f, f' (fp32, x cutoff)
// Per 128-row tile (64 pairs x (value, tangent)):
//   SMEM  A3 : the layer-3 A operand, 2F/64 K-blocks x (hi 16 KB + lo 16 KB) in the TMA SWIZZLE_128B K-major layout, WRITTEN BY
//              THE EPILOGUE WARPS (st.shared + fence.proxy.async).  The h1 tile (layer-2 A operand, F/64 K-blocks) is TMA-loaded
//              into the LAST F/64 K-blocks of the same region: it is dead once layer 2's MMAs retire, which is exactly when the
//              second half of h2 is written there.
//         ring: 3 slots of one weight K-block (NC rows x 128 B x (hi, lo)), streamed from L2 by TMA
//   TMEM  [0, 2NC) layer-2 accumulator as two NC-column chunks, [2NC, 4NC) layer-3 accumulator chunks, double-buffered (NC = F)
//   warps 0: weight TMA   1: MMA issuer   2: h1-tile TMA   3..18: epilogue (TMEM lane quarter = warp & 3, 32-column sub-chunk)
// HBM traffic per chain call drops from 2.53 GB to ~1.4 GB (h2 is 1.13 GB of write + read).
struct TailParams {
  int F;                 // 64 or 128: K2 = F, N2 = 2F = K3, N3 = 3F
  int mul_cutoff;
  float lo, hi;
  const float* bias2;    // [2F]
  const float* bias3;    // [3F]
  const float* pair_r;
  const int* num_unique;
  int max_pairs;
  float* out_f;          // [P, 3F]
  float* out_fp;
  int dbg;               // dev probe (TMD_TC_DBG): 1 no global stores, 2 no shared-memory stores, 3 no epilogue math, 4 no MMA
};

constexpr int TAIL_EPI_WARPS = 16;
constexpr int TAIL_THREADS = 96 + 32 * TAIL_EPI_WARPS;
constexpr int TAIL_SLOTS = 3;
constexpr int KBLK_BYTES = 2 * A_TILE_BYTES;   // one A K-block: hi + lo

__global__ void __launch_bounds__(TAIL_THREADS, 1)
chain_tail_f16_kernel(const __grid_constant__ CUtensorMap map_h1_hi, const __grid_constant__ CUtensorMap map_h1_lo,
                      const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
                      const __grid_constant__ CUtensorMap map_w3_hi, const __grid_constant__ CUtensorMap map_w3_lo,
                      TailParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int NC = P.F;
  const int KB2 = P.F / 64, KB3 = 2 * P.F / 64;
  uint8_t* a3 = smem;                                   // KB3 K-blocks
  uint8_t* a2 = a3 + (size_t)(KB3 - KB2) * KBLK_BYTES;  // aliases the last KB2 K-blocks of a3
  const int slot_bytes = NC * 128 * 2;
  uint8_t* ring = a3 + (size_t)KB3 * KBLK_BYTES;
  __shared__ __align__(8) uint64_t full_bar[TAIL_SLOTS], empty_bar[TAIL_SLOTS];
  __shared__ __align__(8) uint64_t a2_full, a3_free, acc2_full[2], acc2_empty[2], a3_ready[2], acc3_full[2], acc3_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_pairs = min(*P.num_unique, P.max_pairs);
  const int rows_valid = n_pairs * 2;
  const int m_tiles = (rows_valid + BM - 1) / BM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TAIL_SLOTS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&a2_full, 1); mbar_init(&a3_free, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc2_full[a], 1); mbar_init(&acc2_empty[a], TAIL_EPI_WARPS); mbar_init(&a3_ready[a], TAIL_EPI_WARPS);
      mbar_init(&acc3_full[a], 1); mbar_init(&acc3_empty[a], TAIL_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ weight stream: 2*KB2 + 3*KB3 slots per tile
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
        for (int g = 0; g < 5; ++g) {                    // g 0,1: layer-2 chunks; 2..4: layer-3 chunks
          const bool l2 = g < 2;
          const int kbs = l2 ? KB2 : KB3;
          const int row0 = (l2 ? g : g - 2) * NC;
          for (int kb = 0; kb < kbs; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* st = ring + (size_t)stage * slot_bytes;
            mbar_expect_tx(&full_bar[stage], (uint32_t)slot_bytes);
            tma_load_2d(l2 ? &map_w2_hi : &map_w3_hi, &full_bar[stage], st, kb * 64, row0);
            tma_load_2d(l2 ? &map_w2_lo : &map_w3_lo, &full_bar[stage], st + slot_bytes / 2, kb * 64, row0);
            if (++stage == TAIL_SLOTS) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ h1 tile -> the tail of the A3 region
    if (elect_one()) {
      int it = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
        if (it > 0) mbar_wait(&a3_free, (uint32_t)((it - 1) & 1));   // layer-3 MMAs of the previous tile have retired
        mbar_expect_tx(&a2_full, (uint32_t)(KB2 * KBLK_BYTES));
        for (int kb = 0; kb < KB2; ++kb) {
          tma_load_2d(&map_h1_hi, &a2_full, a2 + (size_t)kb * KBLK_BYTES, kb * 64, tile * BM);
          tma_load_2d(&map_h1_lo, &a2_full, a2 + (size_t)kb * KBLK_BYTES + A_TILE_BYTES, kb * 64, tile * BM);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_f16(BM, NC);
    int stage = 0; uint32_t phase = 0;
    int it = 0, g3 = 0;
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
      const uint32_t tp = (uint32_t)(it & 1);
      mbar_wait(&a2_full, tp);
      for (int c = 0; c < 2; ++c) {                       // layer 2: D2[:, c*NC : (c+1)*NC]
        mbar_wait(&acc2_empty[c], tp ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(c * NC);
        for (int kb = 0; kb < KB2; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if (P.dbg != 4 && elect_one()) {
            const uint32_t sa = smem_u32(a2 + (size_t)kb * KBLK_BYTES);
            const uint32_t sb = smem_u32(ring + (size_t)stage * slot_bytes);
            const uint64_t a_hi = make_desc_sw128(sa), a_lo = make_desc_sw128(sa + A_TILE_BYTES);
            const uint64_t b_hi = make_desc_sw128(sb), b_lo = make_desc_sw128(sb + slot_bytes / 2);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adv = (uint64_t)((k * 32) >> 4);
              umma_f16(d_tmem, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
              umma_f16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
              umma_f16(d_tmem, a_hi + adv, b_hi + adv, idesc, 1);
            }
          }
          __syncwarp();
          if (elect_one()) umma_commit(&empty_bar[stage]);
          __syncwarp();
          if (++stage == TAIL_SLOTS) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(&acc2_full[c]);
        __syncwarp();
      }
      for (int c = 0; c < 3; ++c, ++g3) {                 // layer 3: D3[:, c*NC : (c+1)*NC], A = h2 in shared memory
        const int b = g3 & 1;
        mbar_wait(&acc3_empty[b], (uint32_t)(((g3 >> 1) & 1) ^ 1));
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(2 * NC + b * NC);
        for (int kb = 0; kb < KB3; ++kb) {
          if (c == 0 && kb == 0) mbar_wait(&a3_ready[0], tp);          // first half of h2 written (and proxy-fenced)
          if (c == 0 && kb == KB3 / 2) mbar_wait(&a3_ready[1], tp);    // second half
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if (P.dbg != 4 && elect_one()) {
            const uint32_t sa = smem_u32(a3 + (size_t)kb * KBLK_BYTES);
            const uint32_t sb = smem_u32(ring + (size_t)stage * slot_bytes);
            const uint64_t a_hi = make_desc_sw128(sa), a_lo = make_desc_sw128(sa + A_TILE_BYTES);
            const uint64_t b_hi = make_desc_sw128(sb), b_lo = make_desc_sw128(sb + slot_bytes / 2);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adv = (uint64_t)((k * 32) >> 4);
              umma_f16(d_tmem, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
              umma_f16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
              umma_f16(d_tmem, a_hi + adv, b_hi + adv, idesc, 1);
            }
          }
          __syncwarp();
          if (elect_one()) umma_commit(&empty_bar[stage]);
          __syncwarp();
          if (++stage == TAIL_SLOTS) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(&acc3_full[b]);
        __syncwarp();
      }
      if (elect_one()) umma_commit(&a3_free);              // every MMA reading this tile's A3 has retired
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 3..18
    const int q = warp & 3;
    const int sub = (warp - 3) >> 2;                      // 32-column sub-chunk of an NC-column chunk (NC = 64: sub 2,3 idle)
    const bool has_cols = sub * 32 < NC;
    const int row_in_tile = q * 32 + lane;
    const bool tangent = lane & 1;
    const int r_v = row_in_tile & ~1;                     // value row of this lane's pair inside the tile
    int it = 0, g3 = 0;
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
      const uint32_t tp = (uint32_t)(it & 1);
      const int row_v = tile * BM + r_v;
      const bool valid = row_v < rows_valid;
      const int pair = row_v >> 1;
      float cut = 1.f, dcut = 0.f;
      if (P.mul_cutoff) {
        const float r = valid ? __ldg(P.pair_r + pair) : 0.f;
        tn::cosine_cutoff(r, P.lo, P.hi, &cut, &dcut);
      }
      for (int c = 0; c < 2; ++c) {                       // h2 chunk c -> A3 K-blocks
        mbar_wait(&acc2_full[c], tp);
        tcgen05_fence_after();
        if (has_cols) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * NC + sub * 32), v);
          const int kcol = c * NC + sub * 32 + (tangent ? 16 : 0);     // this lane's 16 columns of h2 = K index of layer 3
          float h[16], dh[16];
          if (P.dbg != 3) jvp_math<true, false>(v, tangent, P.bias2 + kcol, 1.f, 0.f, h, dh);
          else { for (int i = 0; i < 16; ++i) { h[i] = __uint_as_float(v[i]); dh[i] = __uint_as_float(v[16 + i]); } }
          uint8_t* kb_base = a3 + (size_t)(kcol >> 6) * KBLK_BYTES;
          const int j0 = (kcol & 63) >> 3;                               // 16-byte chunk index inside the 128-byte row
#pragma unroll
          for (int rr = 0; rr < 2 && P.dbg != 2; ++rr) {                // rr 0: value row (h), 1: tangent row (dh)
            const int r = r_v + rr;
            const float* src = rr ? dh : h;
            uint8_t* rowp = kb_base + (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              uint32_t hi4[4], lo4[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float x0 = src[jj * 8 + 2 * u], x1 = src[jj * 8 + 2 * u + 1];
                const __half2 hh = __floats2half2_rn(x0, x1);
                const float2 hf = __half22float2(hh);
                const __half2 hl = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
                hi4[u] = *reinterpret_cast<const uint32_t*>(&hh);
                lo4[u] = *reinterpret_cast<const uint32_t*>(&hl);
              }
              const int off = ((j0 + jj) ^ (r & 7)) * 16;
              *reinterpret_cast<uint4*>(rowp + off) = make_uint4(hi4[0], hi4[1], hi4[2], hi4[3]);
              *reinterpret_cast<uint4*>(rowp + A_TILE_BYTES + off) = make_uint4(lo4[0], lo4[1], lo4[2], lo4[3]);
            }
          }
        }
        tcgen05_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA's async proxy
        __syncwarp();
        if (lane == 0) { mbar_arrive(&acc2_empty[c]); mbar_arrive(&a3_ready[c]); }
      }
      for (int c = 0; c < 3; ++c, ++g3) {                 // f, f' chunk c -> HBM
        const int b = g3 & 1;
        mbar_wait(&acc3_full[b], (uint32_t)((g3 >> 1) & 1));
        tcgen05_fence_after();
        if (has_cols) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(2 * NC + b * NC + sub * 32), v);
          const int cb = c * NC + sub * 32 + (tangent ? 16 : 0);
          float h[16], dh[16];
          if (P.dbg == 3) { for (int i = 0; i < 16; ++i) { h[i] = __uint_as_float(v[i]); dh[i] = __uint_as_float(v[16 + i]); } }
          else if (P.mul_cutoff) jvp_math<true, true>(v, tangent, P.bias3 + cb, cut, dcut, h, dh);
          else jvp_math<true, false>(v, tangent, P.bias3 + cb, cut, dcut, h, dh);
          if (valid && P.dbg != 1) {
            const int N3 = 3 * NC;
            float* d0 = P.out_f + (size_t)pair * N3 + cb;
            float* d1 = P.out_fp + (size_t)pair * N3 + cb;
            st8(d0, h); st8(d0 + 8, h + 8);
            st8(d1, dh); st8(d1 + 8, dh + 8);
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc3_empty[b]);
      }
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

// ---------------------------------------------------------------------------------------------------------------------
// a9: the three bias-free F x F node linears of TensorNet as ONE grouped 3xTF32 GEMM that reads the fp32 irreps DIRECTLY.
// The A operand is TMA-loaded as raw fp32 K-blocks; converter warps split it in shared memory, in place (hi = rna_tf32(x)
// overwrites x, lo = rna_tf32(x - hi) goes to the twin tile at the same swizzled offsets -- the split is element-wise, so the
// SWIZZLE_128B layout is untouched), publish it to the async proxy (fence.proxy.async) and hand the stage to the MMA warp.
// No split pass over HBM, no padded operand / result buffers: the epilogue stores straight into the three output tensors.
//   warps 0: TMA   1: MMA issuer   2..9: converters   10..17: epilogue (TMEM lane quarter = warp & 3)
struct IrrParams {
  int F;            // K = N = BN = F (64, 128, 192 or 256)
  int n_atoms;
  int t1, t2, tiles;  // m-tiles [0,t1) are I rows, [t1,t2) A rows, [t2,tiles) S rows
  float* out[3];
};
constexpr int IRR_CONV_WARPS = 8;
constexpr int IRR_EPI_WARPS = 8;
constexpr int IRR_THREADS = 64 + 32 * (IRR_CONV_WARPS + IRR_EPI_WARPS);

template <int STAGES>
__global__ void __launch_bounds__(IRR_THREADS, 1)
irreps_linear_kernel(const __grid_constant__ CUtensorMap map_i, const __grid_constant__ CUtensorMap map_a,
                     const __grid_constant__ CUtensorMap map_s, const __grid_constant__ CUtensorMap map_b_hi,
                     const __grid_constant__ CUtensorMap map_b_lo, IrrParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int BN = P.F;
  const int b_tile_bytes = BN * BK * 4;
  const int stage_bytes = 2 * A_TILE_BYTES + 2 * b_tile_bytes;
  const int k_blocks = P.F / BK;
  __shared__ __align__(8) uint64_t full_bar[STAGES], conv_bar[STAGES], empty_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&conv_bar[s], IRR_CONV_WARPS); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], IRR_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
        const int seg = (tile >= P.t1) + (tile >= P.t2);
        const int ltile = tile - (seg == 0 ? 0 : (seg == 1 ? P.t1 : P.t2));
        const CUtensorMap* ma = seg == 0 ? &map_i : (seg == 1 ? &map_a : &map_s);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + (size_t)stage * stage_bytes;
          mbar_expect_tx(&full_bar[stage], (uint32_t)(A_TILE_BYTES + 2 * b_tile_bytes));
          tma_load_2d(ma, &full_bar[stage], st, kb * BK, ltile * BM);
          tma_load_2d(&map_b_hi, &full_bar[stage], st + 2 * A_TILE_BYTES, kb * BK, seg * P.F);
          tma_load_2d(&map_b_lo, &full_bar[stage], st + 2 * A_TILE_BYTES + b_tile_bytes, kb * BK, seg * P.F);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_tf32(BM, BN);
    int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&conv_bar[stage], phase);                 // operands split and proxy-fenced by the converters
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t st = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t a_hi = make_desc_sw128(st), a_lo = make_desc_sw128(st + A_TILE_BYTES);
          const uint64_t b_hi = make_desc_sw128(st + 2 * A_TILE_BYTES), b_lo = make_desc_sw128(st + 2 * A_TILE_BYTES + b_tile_bytes);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            umma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
            umma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
            umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, 1);
          }
        }
        __syncwarp();
        if (elect_one()) umma_commit(&empty_bar[stage]);
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tmem_full_bar[acc]);
      __syncwarp();
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp < 2 + IRR_CONV_WARPS) {
    // ------------------------------------------------------------------ converters: fp32 K-block -> (hi, lo) TF32, in place
    const int cidx = (warp - 2) * 32 + lane;
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        uint8_t* st = smem + (size_t)stage * stage_bytes;
#pragma unroll
        for (int u = 0; u < A_TILE_BYTES / 16 / (32 * IRR_CONV_WARPS); ++u) {
          const int off = (cidx + u * 32 * IRR_CONV_WARPS) * 16;
          const float4 x = *reinterpret_cast<const float4*>(st + off);
          float4 h, l;
          h.x = tf32_rna(x.x); h.y = tf32_rna(x.y); h.z = tf32_rna(x.z); h.w = tf32_rna(x.w);
          l.x = tf32_rna(x.x - h.x); l.y = tf32_rna(x.y - h.y); l.z = tf32_rna(x.z - h.z); l.w = tf32_rna(x.w - h.w);
          *reinterpret_cast<float4*>(st + off) = h;
          *reinterpret_cast<float4*>(st + A_TILE_BYTES + off) = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> fp32 rows of the output tensors
    const int q = warp & 3;
    const int sub = (warp - 2 - IRR_CONV_WARPS) >> 2;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
      const int seg = (tile >= P.t1) + (tile >= P.t2);
      const int ltile = tile - (seg == 0 ? 0 : (seg == 1 ? P.t1 : P.t2));
      const long long rows_seg = (long long)P.n_atoms * (seg == 0 ? 1 : (seg == 1 ? 3 : 5));
      const long long row = (long long)ltile * BM + q * 32 + lane;
      float* orow = P.out[seg] + row * P.F;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
      for (int c0 = sub * 32; c0 < BN; c0 += 32 * (IRR_EPI_WARPS / 4)) {
        uint32_t v[32];
        tmem_ld32(t_row + (uint32_t)c0, v);
        if (row < rows_seg) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) st8(orow + c0 + i, reinterpret_cast<const float*>(v) + i);
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
  static int fuse_tail = -1;   // TMD_TC_FUSE=0 (dev knob) keeps the three per-layer launches
  if (fuse_tail < 0) { const char* e = getenv("TMD_TC_FUSE"); fuse_tail = e ? atoi(e) : 1; }
  const bool tail_ok = fuse_tail && use_f16 && n_layers == 3 && out_fp && act && (dims[1] == 64 || dims[1] == 128) &&
                       dims[2] == 2 * dims[1] && dims[3] == 3 * dims[1];
  for (int l = 0; l < n_layers; ++l) {
    if (l == 1 && tail_ok) {
      // layers 2 + 3 in one kernel; in_hi / in_lo = h1 written by the layer-1 launch above
      const int F = dims[1];
      TMD_REQUIRE((reinterpret_cast<uintptr_t>(bias[1]) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias[2]) & 15) == 0,
                  "pair_chain_tc: biases must be 16-byte aligned");
      CUtensorMap mh_hi, mh_lo, m2_hi, m2_lo, m3_hi, m3_lo;
      int rc;
      if ((rc = make_map(&mh_hi, in_hi, rows_pad, F, BM, 1))) return rc;
      if ((rc = make_map(&mh_lo, in_lo, rows_pad, F, BM, 1))) return rc;
      if ((rc = make_map(&m2_hi, W_hi[1], 2 * F, F, F, 1))) return rc;
      if ((rc = make_map(&m2_lo, W_lo[1], 2 * F, F, F, 1))) return rc;
      if ((rc = make_map(&m3_hi, W_hi[2], 3 * F, 2 * F, F, 1))) return rc;
      if ((rc = make_map(&m3_lo, W_lo[2], 3 * F, 2 * F, F, 1))) return rc;
      TailParams T;
      T.F = F; T.mul_cutoff = mul_cutoff; T.lo = (float)lo; T.hi = (float)hi; T.bias2 = bias[1]; T.bias3 = bias[2];
      T.pair_r = pair_r; T.num_unique = num_unique; T.max_pairs = max_pairs; T.out_f = out_f; T.out_fp = out_fp;
      { const char* e = getenv("TMD_TC_DBG"); T.dbg = e ? atoi(e) : 0; }
      const int smem = (2 * F / 64) * KBLK_BYTES + TAIL_SLOTS * F * 128 * 2 + 1024;
      const int tiles = rows_pad / BM;
      cudaFuncSetAttribute(chain_tail_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      chain_tail_f16_kernel<<<tiles < 148 ? tiles : 148, TAIL_THREADS, smem, st>>>(mh_hi, mh_lo, m2_hi, m2_lo, m3_hi, m3_lo, T);
      break;
    }
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
    { const char* e = getenv("TMD_TC_DBG"); P.dbg = e ? atoi(e) : 0; }
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

// a9: the three bias-free F x F node linears of TensorNet (I [N,1,F], A [N,3,F], S [N,5,F], each with its own weight) as ONE
// grouped 3xTF32 GEMM on tcgen05, reading the fp32 irreps directly (split in shared memory) and writing oI / oA / oS.
// W_hi / W_lo: TF32 split of the stacked weights [3F, F] (rows 0..F-1 act on I, F..2F-1 on A, 2F..3F-1 on S).
int tmd_irreps_linear_tc(const float* I, const float* A, const float* S, int N, int F, const float* W_hi, const float* W_lo,
                         float* oI, float* oA, float* oS, void* stream) {
  TMD_REQUIRE(I && A && S && W_hi && W_lo && oI && oA && oS, "irreps_linear_tc: null pointer");
  TMD_REQUIRE(F % 64 == 0 && F <= 256, "irreps_linear_tc: F must be a multiple of 64, <= 256");
  if (N == 0) return TMD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  auto tiles_of = [](long long rows) { return (int)((rows + BM - 1) / BM); };
  IrrParams P;
  P.F = F; P.n_atoms = N; P.t1 = tiles_of(N); P.t2 = P.t1 + tiles_of(3LL * N); P.tiles = P.t2 + tiles_of(5LL * N);
  P.out[0] = oI; P.out[1] = oA; P.out[2] = oS;
  CUtensorMap mi, ma, ms, mb_hi, mb_lo;
  int rc;
  if ((rc = make_map(&mi, I, N, F, BM))) return rc;
  if ((rc = make_map(&ma, A, 3 * N, F, BM))) return rc;
  if ((rc = make_map(&ms, S, 5 * N, F, BM))) return rc;
  if ((rc = make_map(&mb_hi, W_hi, 3 * F, F, F))) return rc;
  if ((rc = make_map(&mb_lo, W_lo, 3 * F, F, F))) return rc;
  const int stage_bytes = 2 * A_TILE_BYTES + 2 * F * BK * 4;
  const int grid = P.tiles < 148 ? P.tiles : 148;
  if (stage_bytes * 3 + 1024 <= 200 * 1024) {
    const int smem = stage_bytes * 3 + 1024;
    cudaFuncSetAttribute(irreps_linear_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    irreps_linear_kernel<3><<<grid, IRR_THREADS, smem, st>>>(mi, ma, ms, mb_hi, mb_lo, P);
  } else {
    const int smem = stage_bytes * 2 + 1024;
    cudaFuncSetAttribute(irreps_linear_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    irreps_linear_kernel<2><<<grid, IRR_THREADS, smem, st>>>(mi, ma, ms, mb_hi, mb_lo, P);
  }
  return tmd_check_launch("irreps_linear_tc");
}

}  // extern "C"
