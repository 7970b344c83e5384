// a7 (interaction message passing, forward and backward) as CTA-cooperative gathers.
//
// The warp-per-atom kernels of message_passing.cu / edge_fused.cu gather the neighbour's node row (9 F-vectors = 4.6 KB at
// F = 128) once per EDGE; atoms that sit next to each other share most of their neighbours, and the kernels were limited by
// L2->SM bytes (8-9 TB/s of ~12), not by HBM (profiles/r01_mp_ncu.md).  Here a CTA owns a GROUP of G = 15 rows that are
// spatial neighbours (consecutive entries of the neighbour list's cell order), and
//   * every DISTINCT neighbour row of the group is staged ONCE in shared memory: a producer warp walks the group's sorted
//     neighbour union and issues three bulk async copies (cp.async.bulk, I / A / S segments) per row into an mbarrier ring;
//     the 15 consumer warps (one row each) read it with LDS.  On the C4 water box 15 rows have ~200 distinct neighbours for
//     ~800 edges, so gather traffic drops from 4.6 KB to ~1.1 KB per edge;
//   * the per-pair radial rows f_p (and f'_p in the backward) are streamed through a private per-warp ring of bulk async
//     copies issued D edges ahead, so bytes in flight are bounded by shared memory (170-220 KB per SM), not by registers;
//   * a consumer only synchronises on the union elements its own row contains: the plan stores, per edge, the position of
//     its neighbour in the group's union (`upos`), and per union element the number of rows that need it (`ucount`); the
//     producer pre-arrives on the slot's "empty" barrier for the rows that do not.
// Every row still accumulates its edges in ascending-j order with the same FMA sequence as the warp-per-atom kernels, so
// results are BIT-IDENTICAL to them (tests/test_gpu_fused.py::test_cooperative_message_passing_bit_exact).
//
// The per-frame plan (tmd_mp_plan) is one launch shared by the 3 forward + 3 backward launches of a 3-layer frame.
// Reference semantics: torchmdnet/extensions/warp_kernels/tensornet_mp.py:42-131 (message_passing_fwd / _bwd).
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int G = 15;            // rows per group = consumer warps per CTA (+ 1 producer warp = 16 warps: registers are
                                 // allocated per 4 warps, so 17 warps would be charged as 20 and cap the kernel at 96 registers)
constexpr int UMAX = 2048;       // distinct neighbours of a group the staged path handles (else: direct-gather fallback)
constexpr int GCAP = 4096;       // edges of a group the plan kernel sorts (else: fallback)
constexpr int PLAN_THREADS = 256;
constexpr int COOP_THREADS = 32 * (G + 1);
constexpr unsigned SELF_BIT = 0x8000u;

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
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
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
// non-blocking probe of a phase (test_wait never suspends the thread) and a pure polling wait built on it
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
template <bool POLL>
__device__ __forceinline__ void mbar_wait_c(uint64_t* bar, uint32_t parity) {
  if (POLL) { while (!mbar_test(bar, parity)) {} }
  else mbar_wait(bar, parity);
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float4 fma4(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

struct CIrr { const float* I; const float* A; const float* S; };
struct MIrr { float* I; float* A; float* S; };
#ifdef TMD_COOP_PROF
__device__ long long* g_prof = nullptr;   // [groups][16 warps][8]: total, issued spin, full wait, f wait, lane-0 tail, edges, start, end
#define PROF_T0() const long long _t0 = clock64()
#define PROF_ADD(var) (var) += clock64() - _t0
#else
#define PROF_T0()
#define PROF_ADD(var)
#endif
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

// ------------------------------------------------------------------------------------------------------------ plan
// One CTA per group g of G rows (rows = order[G g .. G g + G) when an order is given and *order_flag != 0, else consecutive
// atoms).  Output: union_j / ucount [goff[g] .. goff[g] + gcnt[g]) = sorted distinct neighbours of the group and how many of
// its rows contain each; upos[k] = position of nbr_j[k] in that union (| SELF_BIT for the self loop); gcnt[g] = -1 when the
// group exceeds the staged path's capacity (the MP kernels then gather directly for it).  Space is allocated from `cursor`
// (one int, zeroed by the caller's memset): sum of gcnt <= number of edges.
__global__ void __launch_bounds__(PLAN_THREADS) mp_plan_kernel(const int* __restrict__ indptr, const int* __restrict__ nbr_j,
                                                               const int* __restrict__ num_pairs, int max_pairs, int n, int gs,
                                                               const int* __restrict__ order, const int* __restrict__ order_flag,
                                                               int* __restrict__ rows_out, int* __restrict__ cursor,
                                                               int* __restrict__ union_j, unsigned char* __restrict__ ucount,
                                                               unsigned short* __restrict__ upos, int* __restrict__ goff,
                                                               int* __restrict__ gcnt) {
  __shared__ int keys[GCAP];
  __shared__ unsigned char cnt[GCAP];      // run lengths, indexed by unique position
  __shared__ int s_row[G], s_k0[G], s_len[G], s_base[G + 1];
  __shared__ int s_warp[PLAN_THREADS / 32];
  __shared__ int s_off;
  const int g = blockIdx.x, tid = threadIdx.x;
  const int E = min(*num_pairs, max_pairs);
  const bool use_order = order && (!order_flag || *order_flag != 0);
  if (tid < G) {
    const int r = g * gs + tid;
    int b = -1, k0 = 0, len = 0;
    if (tid < gs && r < n) {
      b = use_order ? order[r] : r;
      k0 = min(indptr[b], E);
      len = min(indptr[b + 1], E) - k0;
    }
    s_row[tid] = b; s_k0[tid] = k0; s_len[tid] = len;
    if (tid < gs && r < n) rows_out[r] = b;
  }
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int w = 0; w < G; ++w) { s_base[w] = run; run += s_len[w]; }
    s_base[G] = run;
  }
  __syncthreads();
  const int total = s_base[G];
  if (total > GCAP || total == 0) {
    if (tid == 0) { gcnt[g] = total == 0 ? 0 : -1; goff[g] = 0; }
    return;
  }
  int n2 = PLAN_THREADS;
  while (n2 < total) n2 <<= 1;
  for (int w = 0; w < G; ++w)
    for (int e = tid; e < s_len[w]; e += PLAN_THREADS) keys[s_base[w] + e] = nbr_j[s_k0[w] + e];
  for (int i = total + tid; i < n2; i += PLAN_THREADS) keys[i] = 0x7fffffff;
  __syncthreads();
  // bitonic sort of n2 keys
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n2; i += PLAN_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const int a = keys[i], b = keys[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  // unique: thread t owns the contiguous chunk [t C, (t+1) C)
  const int C = n2 / PLAN_THREADS;   // <= 16
  int v[GCAP / PLAN_THREADS];
  int heads = 0;
  const int prev = (tid * C > 0) ? keys[tid * C - 1] : -1;
#pragma unroll
  for (int q = 0; q < GCAP / PLAN_THREADS; ++q) {
    if (q < C) {
      const int i = tid * C + q;
      v[q] = keys[i];
      const int before = q == 0 ? prev : v[q - 1];
      if (i < total && v[q] != before) ++heads;
    }
  }
  // exclusive scan of `heads` over the block
  const int lane = tid & 31, wid = tid >> 5;
  int x = heads;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) s_warp[wid] = x;
  for (int i = tid; i < GCAP; i += PLAN_THREADS) cnt[i] = 0;
  __syncthreads();
  int wbase = 0, U = 0;
  for (int w = 0; w < PLAN_THREADS / 32; ++w) { if (w < wid) wbase += s_warp[w]; U += s_warp[w]; }
  int pos = wbase + x - heads;   // unique position of this thread's first head (exclusive prefix)
  __syncthreads();               // every thread holds its chunk in registers: keys[] may now be overwritten in place
  if (U > UMAX) {
    if (tid == 0) { gcnt[g] = -1; goff[g] = 0; }
    return;
  }
  {
    int p = pos - 1;             // unique index of the element before this chunk
#pragma unroll
    for (int q = 0; q < GCAP / PLAN_THREADS; ++q) {
      if (q < C) {
        const int i = tid * C + q;
        if (i < total) {
          const int before = q == 0 ? prev : v[q - 1];
          if (v[q] != before) { ++p; keys[p] = v[q]; }
        }
      }
    }
  }
  __syncthreads();
  if (tid == 0) s_off = atomicAdd(cursor, U);
  __syncthreads();
  const int off = s_off;
  // per edge: position in the union (binary search) + run-length histogram
  for (int w = 0; w < G; ++w) {
    const int b = s_row[w];
    for (int e = tid; e < s_len[w]; e += PLAN_THREADS) {
      const int k = s_k0[w] + e;
      const int j = nbr_j[k];
      int lo = 0, hi = U - 1;
      while (lo < hi) { const int m = (lo + hi) >> 1; if (keys[m] < j) lo = m + 1; else hi = m; }
      upos[k] = (unsigned short)(lo | (j == b ? SELF_BIT : 0u));
      atomicAdd(reinterpret_cast<unsigned int*>(cnt) + (lo >> 2), 1u << (8 * (lo & 3)));   // byte counter (<= G per byte)
    }
  }
  __syncthreads();
  for (int t = tid; t < U; t += PLAN_THREADS) { union_j[off + t] = keys[t]; ucount[off + t] = cnt[t]; }
  if (tid == 0) { gcnt[g] = U; goff[g] = off; }
}

// ------------------------------------------------------------------------------------------------------------ kernels
// shared-memory layout of the cooperative kernels (dynamic):
//   [0, 8 S)                    full[S]   mbarriers (tx-count: one staged node row)
//   [8 S, 16 S)                 empty[S]  mbarriers (count G: one arrival per row of the group)
//   [16 S, 16 S + 8 G D)        fbar[G][D] mbarriers (tx-count: the radial row(s) of one edge)
//   then issued[S] ints: number of uses of slot s the producer has issued so far.  A consumer skips the union elements its
//   row does not contain, so it can be several uses of a slot ahead of the slot's current one; mbarrier parity waits only
//   tell adjacent phases apart, so before waiting for use n of slot s it first spins until issued[s] > n (the producer issues
//   use n only after every consumer of use n-1 has released the slot, i.e. the barrier is then in phase n).
//   ring  S x 9F floats, then per-warp radial rings G x D x (3F or 6F) floats
template <int S, int D, bool BWD>
struct CoopSmem {
  static constexpr int kBars = 2 * S + G * D;
  static __host__ __device__ size_t bar_bytes() { return ((size_t)kBars * 8 + (size_t)S * 4 + 127) / 128 * 128; }   // + issued[S]
  static __host__ __device__ size_t ring_bytes(int F) { return (size_t)S * 9 * F * 4; }
  static __host__ __device__ size_t fring_bytes(int F) { return (size_t)G * D * (BWD ? 6 : 3) * F * 4; }
  static __host__ __device__ size_t total(int F) { return bar_bytes() + ring_bytes(F) + fring_bytes(F); }
};

// producer warp: lane s owns ring slot s and fills it with the union elements t = s, s + S, s + 2S, ... (the slots are independent,
// so 16-24 lanes issue their barrier operations and bulk copies side by side instead of one lane walking the union serially).
// The lanes poll their slot's "empty" barrier with the non-blocking test_wait in ONE converged loop: a blocking try_wait in
// divergent lanes would park the whole warp on one lane's barrier while other lanes' slots are already free.
template <int S>
__device__ __forceinline__ void producer_loop(uint64_t* full, uint64_t* empty, volatile int* issued, float* ring, const CIrr& T, int F,
                                              const int* __restrict__ union_j, const unsigned char* __restrict__ ucount,
                                              int off, int U, int lane) {
  static_assert(S <= 32, "one producer lane per ring slot");
  const uint32_t row_bytes = 9u * F * 4u;
  const int sl = lane < S ? lane : 0;
  float* dst = ring + (size_t)sl * 9 * F;
  int t = lane, n = 0;
  bool active = lane < S && t < U;
  long long j = active ? union_j[off + t] : 0;
  int c = active ? (int)ucount[off + t] : G;
  while (__any_sync(0xffffffffu, active)) {
    const bool ready = active && (n == 0 || mbar_test(empty + sl, (n - 1) & 1));
    if (ready) {
      mbar_expect_tx(full + sl, row_bytes);
      bulk_g2s(dst, T.I + j * F, (uint32_t)F * 4u, full + sl);
      bulk_g2s(dst + F, T.A + j * 3 * F, 3u * F * 4u, full + sl);
      bulk_g2s(dst + 4 * F, T.S + j * 5 * F, 5u * F * 4u, full + sl);
      if (c < G) mbar_arrive_n(empty + sl, (uint32_t)(G - c));   // rows of the group that do not contain j
      issued[sl] = n + 1;
      t += S; ++n;
      active = t < U;
      if (active) { j = union_j[off + t]; c = (int)ucount[off + t]; }
    }
  }
}

// forward: M_b = sum_{k in row b} f[pid[k]] (x) Y[nbr_j[k]]      (F <= 128, F % 4 == 0)
template <int S, int D, bool POLL>
__global__ void __launch_bounds__(COOP_THREADS, 1) mp_coop_fwd_kernel(CIrr Y, const float* __restrict__ f, const int* __restrict__ pid,
                                                                     const int* __restrict__ nbr_j, const int* __restrict__ indptr,
                                                                     const int* __restrict__ num_pairs, int max_pairs, int n, int F, int gs,
                                                                     const int* __restrict__ rows, const int* __restrict__ union_j,
                                                                     const unsigned char* __restrict__ ucount,
                                                                     const unsigned short* __restrict__ upos,
                                                                     const int* __restrict__ goff, const int* __restrict__ gcnt,
                                                                     MIrr M) {
  extern __shared__ __align__(128) unsigned char smem[];
  using L = CoopSmem<S, D, false>;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + S;
  uint64_t* fbar = empty + S;
  volatile int* issued = reinterpret_cast<volatile int*>(fbar + G * D);
  float* ring = reinterpret_cast<float*>(smem + L::bar_bytes());
  float* fring = ring + (size_t)S * 9 * F;
  const int g = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int U = gcnt[g], off = goff[g];
  const int E = min(*num_pairs, max_pairs);
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, G); issued[s] = 0; }
    for (int i = 0; i < G * D; ++i) mbar_init(fbar + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (wid == G) {   // ---------------------------------------------------------------- producer warp
    if (U > 0) producer_loop<S>(full, empty, issued, ring, Y, F, union_j, ucount, off, U, lane);
    return;
  }
  // ------------------------------------------------------------------------------------ consumer warp = one row
  const int r = g * gs + wid;
  if (wid >= gs || r >= n) return;
#ifdef TMD_COOP_PROF
  long long p_f = 0, p_iss = 0, p_full = 0, p_tail = 0;
  const long long p_start = clock64();
#endif
  const int b = rows[r];
  const int k0 = min(indptr[b], E), nE = min(indptr[b + 1], E) - k0;
  const int h = lane * 4;
  const bool on = h < F;
  float4 acc[9];
#pragma unroll
  for (int c = 0; c < 9; ++c) acc[c] = zero4();
  if (U < 0) {   // group over the staged path's capacity: direct gathers (same arithmetic)
    if (on)
      for (int k = k0; k < k0 + nE; ++k) {
        const long long j = nbr_j[k], p = pid[k];
        const float* fe = f + p * 3 * (long long)F + h;
        const float4 w1 = ldg4(fe), w2 = ldg4(fe + F), w3 = ldg4(fe + 2 * F);
        float4 t[9];
        ld9(Y, j, F, h, t);
        acc[0] = fma4(w1, t[0], acc[0]);
#pragma unroll
        for (int c = 1; c < 4; ++c) acc[c] = fma4(w2, t[c], acc[c]);
#pragma unroll
        for (int c = 4; c < 9; ++c) acc[c] = fma4(w3, t[c], acc[c]);
      }
    if (on) st9(M, b, F, h, acc);
    return;
  }
  float* myf = fring + (size_t)wid * D * 3 * F;
  uint64_t* mybar = fbar + wid * D;
  const uint32_t frow_bytes = 3u * F * 4u;
  // edge metadata (pair id, union position) lives in registers, 32 edges per lane-parallel load and one chunk ahead: no global
  // load sits on the per-edge critical path (the refill of the radial ring needs pid[e + D], possibly from the next chunk)
  int pid_c = lane < nE ? pid[k0 + lane] : 0, up_c = lane < nE ? (int)upos[k0 + lane] : 0;
  int pid_n = 32 + lane < nE ? pid[k0 + 32 + lane] : 0, up_n = 32 + lane < nE ? (int)upos[k0 + 32 + lane] : 0;
  for (int d = 0; d < D && d < nE; ++d) {
    const long long p0 = __shfl_sync(0xffffffffu, pid_c, d);
    if (lane == 0) {
      mbar_expect_tx(mybar + d, frow_bytes);
      bulk_g2s(myf + (size_t)d * 3 * F, f + p0 * 3 * F, frow_bytes, mybar + d);
    }
  }
  for (int e0 = 0; e0 < nE; e0 += 32) {
    const int m = min(32, nE - e0);
    for (int i = 0; i < m; ++i) {
      const int e = e0 + i, d = e % D;
      const int t = __shfl_sync(0xffffffffu, up_c, i) & (int)(SELF_BIT - 1);
      const int q = i + D;
      const long long pn = q < 32 ? __shfl_sync(0xffffffffu, pid_c, q & 31) : __shfl_sync(0xffffffffu, pid_n, q & 31);
      const int s = t % S;
      { PROF_T0(); mbar_wait_c<POLL>(mybar + d, (e / D) & 1); PROF_ADD(p_f); }
      { PROF_T0(); while (issued[s] <= t / S) __nanosleep(64); PROF_ADD(p_iss); }
      { PROF_T0(); mbar_wait_c<POLL>(full + s, (t / S) & 1); PROF_ADD(p_full); }
      if (on) {
        const float* fe = myf + (size_t)d * 3 * F + h;
        const float* row = ring + (size_t)s * 9 * F + h;
        const float4 w1 = *reinterpret_cast<const float4*>(fe), w2 = *reinterpret_cast<const float4*>(fe + F),
                     w3 = *reinterpret_cast<const float4*>(fe + 2 * F);
        acc[0] = fma4(w1, *reinterpret_cast<const float4*>(row), acc[0]);
#pragma unroll
        for (int c = 1; c < 4; ++c) acc[c] = fma4(w2, *reinterpret_cast<const float4*>(row + c * F), acc[c]);
#pragma unroll
        for (int c = 4; c < 9; ++c) acc[c] = fma4(w3, *reinterpret_cast<const float4*>(row + c * F), acc[c]);
      }
      __syncwarp();
      { PROF_T0();
      if (lane == 0) {
        mbar_arrive(empty + s);
        if (e + D < nE) {
          mbar_expect_tx(mybar + d, frow_bytes);
          bulk_g2s(myf + (size_t)d * 3 * F, f + pn * 3 * F, frow_bytes, mybar + d);
        }
      }
      __syncwarp(); PROF_ADD(p_tail); }
    }
    pid_c = pid_n; up_c = up_n;
    pid_n = e0 + 64 + lane < nE ? pid[k0 + e0 + 64 + lane] : 0;
    up_n = e0 + 64 + lane < nE ? (int)upos[k0 + e0 + 64 + lane] : 0;
  }
  if (on) st9(M, b, F, h, acc);
#ifdef TMD_COOP_PROF
  if (g_prof && lane == 0) {
    long long* o = g_prof + ((long long)g * 16 + wid) * 8;
    const long long p_end = clock64();
    o[0] = p_end - p_start; o[1] = p_iss; o[2] = p_full; o[3] = p_f; o[4] = p_tail; o[5] = nE; o[6] = p_start; o[7] = p_end;
  }
#endif
}

// backward: dY_b = sum_k f_p dM_j,   g_dist[k] = < f'_p , dM_j * Y_b >  (0 for the self loop), slot k = (b, j)
template <int S, int D, bool POLL>
__global__ void __launch_bounds__(COOP_THREADS, 1) mp_coop_bwd_kernel(CIrr Y, CIrr dM, const float* __restrict__ f,
                                                                     const float* __restrict__ fp, const int* __restrict__ pid,
                                                                     const int* __restrict__ nbr_j, const int* __restrict__ indptr,
                                                                     const int* __restrict__ num_pairs, int max_pairs, int n, int F, int gs,
                                                                     const int* __restrict__ rows, const int* __restrict__ union_j,
                                                                     const unsigned char* __restrict__ ucount,
                                                                     const unsigned short* __restrict__ upos,
                                                                     const int* __restrict__ goff, const int* __restrict__ gcnt,
                                                                     CIrr add, MIrr dY, float* __restrict__ g_dist) {
  extern __shared__ __align__(128) unsigned char smem[];
  using L = CoopSmem<S, D, true>;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + S;
  uint64_t* fbar = empty + S;
  volatile int* issued = reinterpret_cast<volatile int*>(fbar + G * D);
  float* ring = reinterpret_cast<float*>(smem + L::bar_bytes());
  float* fring = ring + (size_t)S * 9 * F;
  const int g = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int U = gcnt[g], off = goff[g];
  const int E = min(*num_pairs, max_pairs);
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, G); issued[s] = 0; }
    for (int i = 0; i < G * D; ++i) mbar_init(fbar + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (wid == G) {
    if (U > 0) producer_loop<S>(full, empty, issued, ring, dM, F, union_j, ucount, off, U, lane);
    return;
  }
  const int r = g * gs + wid;
  if (wid >= gs || r >= n) return;
  const int b = rows[r];
  const int k0 = min(indptr[b], E), nE = min(indptr[b + 1], E) - k0;
  const int h = lane * 4;
  const bool on = h < F;
  float4 acc[9], yb[9];
#pragma unroll
  for (int c = 0; c < 9; ++c) { acc[c] = zero4(); yb[c] = zero4(); }
  if (on) ld9(Y, b, F, h, yb);
  if (U < 0) {   // direct-gather fallback
    for (int k = k0; k < k0 + nE; ++k) {
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
        for (int c = 1; c < 4; ++c) { acc[c] = fma4(w2, dj[c], acc[c]); g2 = fma4(dj[c], yb[c], g2); }
#pragma unroll
        for (int c = 4; c < 9; ++c) { acc[c] = fma4(w3, dj[c], acc[c]); g3 = fma4(dj[c], yb[c], g3); }
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
    if (on) { if (add.I) { float4 ad[9]; ld9(add, b, F, h, ad); for (int c = 0; c < 9; ++c) acc[c] = add4(acc[c], ad[c]); } st9(dY, b, F, h, acc); }
    return;
  }
  float* myf = fring + (size_t)wid * D * 6 * F;
  uint64_t* mybar = fbar + wid * D;
  const uint32_t frow_bytes = 3u * F * 4u;
  int pid_c = lane < nE ? pid[k0 + lane] : 0, up_c = lane < nE ? (int)upos[k0 + lane] : 0;
  int pid_n = 32 + lane < nE ? pid[k0 + 32 + lane] : 0, up_n = 32 + lane < nE ? (int)upos[k0 + 32 + lane] : 0;
  for (int d = 0; d < D && d < nE; ++d) {
    const long long p0 = __shfl_sync(0xffffffffu, pid_c, d);
    if (lane == 0) {
      mbar_expect_tx(mybar + d, frow_bytes * (fp ? 2u : 1u));
      bulk_g2s(myf + (size_t)d * 6 * F, f + p0 * 3 * F, frow_bytes, mybar + d);
      if (fp) bulk_g2s(myf + (size_t)d * 6 * F + 3 * F, fp + p0 * 3 * F, frow_bytes, mybar + d);
    }
  }
  for (int e0 = 0; e0 < nE; e0 += 32) {
    const int m = min(32, nE - e0);
    float keep = 0.f;                              // lane i keeps g_dist of edge e0 + i: one coalesced store per 32 edges
    for (int i = 0; i < m; ++i) {
      const int e = e0 + i, d = e % D;
      const int up = __shfl_sync(0xffffffffu, up_c, i);
      const int q = i + D;
      const long long pn = q < 32 ? __shfl_sync(0xffffffffu, pid_c, q & 31) : __shfl_sync(0xffffffffu, pid_n, q & 31);
      const int t_ = up & (int)(SELF_BIT - 1);
      const int s = t_ % S;
      mbar_wait_c<POLL>(mybar + d, (e / D) & 1);
      while (issued[s] <= t_ / S) __nanosleep(64);
      mbar_wait_c<POLL>(full + s, (t_ / S) & 1);
      float t = 0.f;
      if (on) {
        const float* fe = myf + (size_t)d * 6 * F + h;
        const float* row = ring + (size_t)s * 9 * F + h;
        const float4 w1 = *reinterpret_cast<const float4*>(fe), w2 = *reinterpret_cast<const float4*>(fe + F),
                     w3 = *reinterpret_cast<const float4*>(fe + 2 * F);
        const float4 d0 = *reinterpret_cast<const float4*>(row);
        acc[0] = fma4(w1, d0, acc[0]);
        float4 g2 = zero4(), g3 = zero4();
#pragma unroll
        for (int c = 1; c < 4; ++c) {
          const float4 dc = *reinterpret_cast<const float4*>(row + c * F);
          acc[c] = fma4(w2, dc, acc[c]); g2 = fma4(dc, yb[c], g2);
        }
#pragma unroll
        for (int c = 4; c < 9; ++c) {
          const float4 dc = *reinterpret_cast<const float4*>(row + c * F);
          acc[c] = fma4(w3, dc, acc[c]); g3 = fma4(dc, yb[c], g3);
        }
        if (fp) {
          const float* fpe = fe + 3 * F;
          t = dot4(*reinterpret_cast<const float4*>(fpe), mul4(d0, yb[0])) + dot4(*reinterpret_cast<const float4*>(fpe + F), g2) +
              dot4(*reinterpret_cast<const float4*>(fpe + 2 * F), g3);
        }
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(empty + s);
        if (e + D < nE) {
          mbar_expect_tx(mybar + d, frow_bytes * (fp ? 2u : 1u));
          bulk_g2s(myf + (size_t)d * 6 * F, f + pn * 3 * F, frow_bytes, mybar + d);
          if (fp) bulk_g2s(myf + (size_t)d * 6 * F + 3 * F, fp + pn * 3 * F, frow_bytes, mybar + d);
        }
      }
      if (fp) {
        t = warp_sum(t);
        if (lane == i) keep = (up & (int)SELF_BIT) ? 0.f : t;
      }
    }
    if (fp && lane < m) g_dist[k0 + e0 + lane] = keep;
    pid_c = pid_n; up_c = up_n;
    pid_n = e0 + 64 + lane < nE ? pid[k0 + e0 + 64 + lane] : 0;
    up_n = e0 + 64 + lane < nE ? (int)upos[k0 + e0 + 64 + lane] : 0;
  }
  if (on) {
    if (add.I) {   // gradient that reached Y_b along another path (P2): summed here instead of by a separate launch
      float4 ad[9];
      ld9(add, b, F, h, ad);
#pragma unroll
      for (int c = 0; c < 9; ++c) acc[c] = add4(acc[c], ad[c]);
    }
    st9(dY, b, F, h, acc);
  }
}

}  // namespace

extern "C" {

#ifdef TMD_COOP_PROF
int tmd_mp_coop_set_profile(long long* buf) { cudaMemcpyToSymbol(g_prof, &buf, sizeof(buf)); return 0; }
#endif

// largest number of rows per cooperative group (= consumer warps per CTA)
int tmd_mp_group_size(void) { return G; }

// rows per group for n rows: every launch is one CTA per group on 148 SMs with one resident CTA each, and a group's time is
// proportional to its rows, so the frame cost is rounds x rows; among G-3 .. G pick the size with the smallest product (a
// 10 125-atom frame: 15 rows -> 675 groups = 5 rounds (4.56 filled), 14 rows -> 724 groups = 5 rounds (4.89 filled): 7 % less).
// The plan arrays goff / gcnt have ceil(n / rows) entries.
int tmd_mp_pick_group(int n) {
  int best = G; long long best_cost = -1;
  for (int gs = G; gs >= G - 3 && gs >= 1; --gs) {
    const long long groups = (n + gs - 1) / gs, rounds = (groups + 147) / 148;
    const long long cost = rounds * gs;
    if (best_cost < 0 || cost < best_cost) { best = gs; best_cost = cost; }
  }
  return best;
}

// workspace sizes (elements): rows[n] int, union_j[max_pairs] int, ucount[max_pairs] u8, upos[max_pairs] u16,
// goff / gcnt [ceil(n / tmd_mp_group_size())] int, cursor[1] int (zeroed by this call)
int tmd_mp_plan(const int* indptr, const int* nbr_j, const int* num_pairs, int max_pairs, int n, int gs, const int* order,
                const int* order_flag, int* rows, int* cursor, int* union_j, unsigned char* ucount, unsigned short* upos,
                int* goff, int* gcnt, void* stream) {
  TMD_REQUIRE(indptr && nbr_j && num_pairs && rows && cursor && union_j && ucount && upos && goff && gcnt, "mp_plan: null pointer");
  TMD_REQUIRE(gs >= 1 && gs <= G, "mp_plan: rows per group must be 1 .. tmd_mp_group_size()");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(cursor, 0, sizeof(int), st);
  if (n == 0) return tmd_check_launch("mp_plan(empty)");
  mp_plan_kernel<<<(n + gs - 1) / gs, PLAN_THREADS, 0, st>>>(indptr, nbr_j, num_pairs, max_pairs, n, gs, order, order_flag, rows, cursor,
                                                          union_j, ucount, upos, goff, gcnt);
  return tmd_check_launch("mp_plan");
}

int tmd_mp_coop_fwd(const float* YI, const float* YA, const float* YS, const float* f, const int* pid, const int* nbr_j,
                    const int* indptr, const int* num_pairs, int max_pairs, int n, int F, int gs, const int* rows, const int* union_j,
                    const unsigned char* ucount, const unsigned short* upos, const int* goff, const int* gcnt, float* MI,
                    float* MA, float* MS, void* stream) {
  TMD_REQUIRE(F > 0 && F % 4 == 0 && F <= 128, "mp_coop_fwd: F must be a multiple of 4, <= 128");
  TMD_REQUIRE(gs >= 1 && gs <= G, "mp_coop_fwd: rows per group must be 1 .. tmd_mp_group_size()");
  TMD_REQUIRE(YI && YA && YS && f && pid && nbr_j && indptr && num_pairs && rows && union_j && ucount && upos && goff && gcnt && MI && MA && MS,
              "mp_coop_fwd: null pointer");
  if (n == 0) return TMD_OK;
  static const int cfg = getenv("TMD_COOP_FWD") ? atoi(getenv("TMD_COOP_FWD")) : 0;
#define TMD_COOP_FWD_LAUNCH(S_, D_, P_)                                                                                                  \
  do {                                                                                                                              \
    const size_t smem = CoopSmem<S_, D_, false>::total(F);                                                                          \
    cudaFuncSetAttribute(mp_coop_fwd_kernel<S_, D_, P_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                    \
    mp_coop_fwd_kernel<S_, D_, P_><<<(n + gs - 1) / gs, COOP_THREADS, smem, (cudaStream_t)stream>>>(                                     \
        CIrr{YI, YA, YS}, f, pid, nbr_j, indptr, num_pairs, max_pairs, n, F, gs, rows, union_j, ucount, upos, goff, gcnt, MIrr{MI, MA, MS}); \
  } while (0)
  if (cfg == 1) TMD_COOP_FWD_LAUNCH(32, 2, false);
  else if (cfg == 2) TMD_COOP_FWD_LAUNCH(32, 3, false);
  else if (cfg == 3) TMD_COOP_FWD_LAUNCH(28, 3, false);
  else TMD_COOP_FWD_LAUNCH(24, 4, false);
#undef TMD_COOP_FWD_LAUNCH
  return tmd_check_launch("mp_coop_fwd");
}

int tmd_mp_coop_bwd(const float* YI, const float* YA, const float* YS, const float* dMI, const float* dMA, const float* dMS,
                    const float* f, const float* fp, const int* pid, const int* nbr_j, const int* indptr, const int* num_pairs,
                    int max_pairs, int n, int F, int gs, const int* rows, const int* union_j, const unsigned char* ucount,
                    const unsigned short* upos, const int* goff, const int* gcnt, const float* addI, const float* addA,
                    const float* addS, float* dYI, float* dYA, float* dYS, float* g_dist, void* stream) {
  TMD_REQUIRE(F > 0 && F % 4 == 0 && F <= 128, "mp_coop_bwd: F must be a multiple of 4, <= 128");
  TMD_REQUIRE(gs >= 1 && gs <= G, "mp_coop_bwd: rows per group must be 1 .. tmd_mp_group_size()");
  TMD_REQUIRE(YI && YA && YS && dMI && dMA && dMS && f && pid && nbr_j && indptr && num_pairs && rows && union_j && ucount && upos && goff && gcnt && dYI && dYA && dYS,
              "mp_coop_bwd: null pointer");
  TMD_REQUIRE(!fp || g_dist, "mp_coop_bwd: fp given without g_dist");
  TMD_REQUIRE((addI && addA && addS) || (!addI && !addA && !addS), "mp_coop_bwd: the add triple is all-or-nothing");
  if (n == 0) return TMD_OK;
  static const int cfg = getenv("TMD_COOP_BWD") ? atoi(getenv("TMD_COOP_BWD")) : 0;
#define TMD_COOP_BWD_LAUNCH(S_, D_, P_)                                                                                                  \
  do {                                                                                                                              \
    const size_t smem = CoopSmem<S_, D_, true>::total(F);                                                                           \
    cudaFuncSetAttribute(mp_coop_bwd_kernel<S_, D_, P_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                    \
    mp_coop_bwd_kernel<S_, D_, P_><<<(n + gs - 1) / gs, COOP_THREADS, smem, (cudaStream_t)stream>>>(                                     \
        CIrr{YI, YA, YS}, CIrr{dMI, dMA, dMS}, f, fp, pid, nbr_j, indptr, num_pairs, max_pairs, n, F, gs, rows, union_j, ucount, upos,      \
        goff, gcnt, CIrr{addI, addA, addS}, MIrr{dYI, dYA, dYS}, g_dist);                                                                                 \
  } while (0)
  if (cfg == 1) TMD_COOP_BWD_LAUNCH(28, 2, false);
  else if (cfg == 2) TMD_COOP_BWD_LAUNCH(20, 2, false);
  else if (cfg == 3) TMD_COOP_BWD_LAUNCH(16, 3, false);
  else TMD_COOP_BWD_LAUNCH(24, 2, false);
#undef TMD_COOP_BWD_LAUNCH
  return tmd_check_launch("mp_coop_bwd");
}

}  // extern "C"
