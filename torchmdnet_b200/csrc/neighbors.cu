// a1/a2/a3: neighbour list, its position gradient and the COO->CSR/CSC transform.
//
// Replaces torchmdnet/extensions/warp_kernels/neighbors_brute.py:32-205 (+ neighbors_cell.py)
// and graph_transform.py:33-79.  B200-first differences:
//   * deterministic: one warp owns one row i, scans candidate j in ascending order and
//     compacts hits with ballot/popc, so the list comes out sorted by (i, j) and *is* the CSR
//     (row pointer = exclusive scan of the per-row counts).  No global atomic slot counter,
//     no host-side boolean-mask compaction, no separate COO->CSR pass.
//   * candidates are restricted to the row's own `batch` segment (binary search when `batch`
//     is sorted, which is how every caller builds it), so 8 x 5000-atom boxes cost 8 x 5000^2,
//     not 40000^2; for one rectangular periodic box they come from a cell grid (27 cells, O(N)),
//     staged in shared memory and rank-sorted so the emitted rows are identical to the scan's.
//   * arithmetic identical to the reference: canonical orientation (larger index first),
//     delta = pos_a - pos_b, sequential z,y,x minimum-image rounding with roundf (ties away),
//     separate roundings for every mul/sub (no FMA contraction), cut-offs compared squared.
#include "common.cuh"

namespace {

template <typename T> __device__ __forceinline__ T mul_rn(T a, T b);
template <> __device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T sub_rn(T a, T b);
template <> __device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
template <> __device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
template <typename T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
template <typename T> __device__ __forceinline__ T div_rn(T a, T b);
template <> __device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
template <> __device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float round_away(float x) { return roundf(x); }
__device__ __forceinline__ double round_away(double x) { return round(x); }
__device__ __forceinline__ float sqrt_rn(float x) { return __fsqrt_rn(x); }
__device__ __forceinline__ double sqrt_rn(double x) { return __dsqrt_rn(x); }

// seg[i] = [first, last+1) index range of atoms sharing batch[i], assuming `batch` is sorted;
// flag[0] is raised if it is not (rows then scan the whole array).
__global__ void batch_segments_kernel(const int64_t* __restrict__ batch, int n, int* __restrict__ seg_lo,
                                      int* __restrict__ seg_hi, int* __restrict__ unsorted_flag) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int64_t b = batch[i];
    if (i > 0 && batch[i - 1] > b) atomicOr(unsorted_flag, 1);
    int lo = 0, hi = i;  // first index with batch >= b
    while (lo < hi) { int m = (lo + hi) >> 1; if (batch[m] < b) lo = m + 1; else hi = m; }
    seg_lo[i] = lo;
    lo = i; hi = n;      // first index with batch > b
    while (lo < hi) { int m = (lo + hi) >> 1; if (batch[m] <= b) lo = m + 1; else hi = m; }
    seg_hi[i] = lo;
  }
}

template <typename T> struct BoxRow { T b00, b10, b11, b20, b21, b22; };

// minimum-image delta of the directed pair (i, j) computed in the canonical orientation (larger index first) with the
// reference's operation order; returns whether lo2 <= d2 < hi2
template <typename T>
__device__ __forceinline__ bool pair_delta(const T* __restrict__ pos, int i, int j, T xi, T yi, T zi, int periodic,
                                           const BoxRow<T>& B, T lo2, T hi2, T& dx, T& dy, T& dz, T& d2) {
  const bool fwd = i > j;
  const T xj = pos[3 * j], yj = pos[3 * j + 1], zj = pos[3 * j + 2];
  dx = fwd ? sub_rn(xi, xj) : sub_rn(xj, xi);
  dy = fwd ? sub_rn(yi, yj) : sub_rn(yj, yi);
  dz = fwd ? sub_rn(zi, zj) : sub_rn(zj, zi);
  if (periodic) {
    const T s3 = round_away(div_rn(dz, B.b22));
    dx = sub_rn(dx, mul_rn(s3, B.b20)); dy = sub_rn(dy, mul_rn(s3, B.b21)); dz = sub_rn(dz, mul_rn(s3, B.b22));
    const T s2 = round_away(div_rn(dy, B.b11));
    dx = sub_rn(dx, mul_rn(s2, B.b10)); dy = sub_rn(dy, mul_rn(s2, B.b11));
    const T s1 = round_away(div_rn(dx, B.b00));
    dx = sub_rn(dx, mul_rn(s1, B.b00));
  }
  d2 = add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(dz, dz));
  const bool ok = (d2 < hi2) && (d2 >= lo2);
  if (!fwd) { dx = -dx; dy = -dy; dz = -dz; }
  return ok;
}

// ---- cell grids: one per periodic rectangular box.  hdr = {total cells, 0, 0, ok}; boxhdr[4 b ..] = {nx, ny, nz, first cell of box b}.
// A single box (box_stride == 0) is box 0 for every atom.  Any non-rectangular box, a box with fewer than 3 cells along an axis,
// or more cells than the workspace holds switches the whole call to the ordered segment scan (ok = 0).
__global__ void cell_setup_kernel(const float* boxf, const double* boxd, int box_stride, int n_boxes, int periodic, double cutoff,
                                  int max_cells, int* __restrict__ hdr, int* __restrict__ boxhdr) {
  int ok = (periodic && max_cells >= 27 && n_boxes >= 1) ? 1 : 0;
  long long total = 0;
  for (int bx = 0; ok && bx < n_boxes; ++bx) {
    double b[9];
    for (int k = 0; k < 9; ++k) b[k] = boxf ? (double)boxf[(size_t)bx * box_stride + k] : boxd[(size_t)bx * box_stride + k];
    if (b[3] != 0.0 || b[6] != 0.0 || b[7] != 0.0 || b[1] != 0.0 || b[2] != 0.0 || b[5] != 0.0) { ok = 0; break; }  // rectangular only
    const double L[3] = {b[0], b[4], b[8]};
    int nc[3];
    for (int d = 0; d < 3; ++d) {
      nc[d] = (int)floor(L[d] / (cutoff * 1.0001));   // cell edge strictly larger than the cutoff
      if (nc[d] < 3) ok = 0;                           // 27 distinct neighbour cells needed
    }
    if (n_boxes == 1)
      while (ok && (long long)nc[0] * nc[1] * nc[2] > max_cells) {  // coarsen the longest axis until it fits
        int d = (nc[0] >= nc[1] && nc[0] >= nc[2]) ? 0 : (nc[1] >= nc[2] ? 1 : 2);
        if (--nc[d] < 3) ok = 0;
      }
    boxhdr[4 * bx] = nc[0]; boxhdr[4 * bx + 1] = nc[1]; boxhdr[4 * bx + 2] = nc[2]; boxhdr[4 * bx + 3] = (int)total;
    total += (long long)nc[0] * nc[1] * nc[2];
    if (total > max_cells) ok = 0;
  }
  hdr[0] = ok ? (int)total : 0; hdr[1] = 0; hdr[2] = 0; hdr[3] = ok;
}
template <typename T>
__global__ void cell_assign_kernel(const T* __restrict__ pos, const int64_t* __restrict__ batch, const T* __restrict__ box, int box_stride,
                                   int n, const int* __restrict__ hdr, const int* __restrict__ boxhdr, int* __restrict__ cell_of,
                                   int* __restrict__ cell_count) {
  if (!hdr[3]) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int64_t bx = box_stride ? batch[i] : 0;
    const T* bb = box + (size_t)bx * box_stride;
    const T L[3] = {bb[0], bb[4], bb[8]};
    const int nc[3] = {boxhdr[4 * bx], boxhdr[4 * bx + 1], boxhdr[4 * bx + 2]};
    const int ny = nc[1], nz = nc[2];
    int c[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      T f = pos[3 * i + d] / L[d];
      f -= floor(f);
      int k = (int)(f * (T)nc[d]);
      c[d] = k < 0 ? 0 : (k >= nc[d] ? nc[d] - 1 : k);
    }
    const int id = boxhdr[4 * bx + 3] + (c[0] * ny + c[1]) * nz + c[2];
    cell_of[i] = id;
    atomicAdd(cell_count + id, 1);
  }
}
__global__ void cell_fill_kernel(int n, const int* __restrict__ hdr, const int* __restrict__ cell_of,
                                 const int* __restrict__ cell_start, int* __restrict__ cursor, int* __restrict__ cell_atoms) {
  if (!hdr[3]) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = cell_of[i];
    cell_atoms[cell_start[c] + atomicAdd(cursor + c, 1)] = i;
  }
}

constexpr int ROW_CAP = 192;  // neighbours of one atom staged in shared memory for the in-row sort (cell path)

// One warp owns one row i.  Candidates come either from the row's batch segment in ascending order (brute, already sorted)
// or from the 27 neighbouring cells (unordered -> staged in shared memory and rank-sorted by j, so the emitted row is
// identical to the brute one).  FILL=false only counts.
template <typename T, bool FILL>
__global__ void __launch_bounds__(256) neighbor_rows_kernel(
    const T* __restrict__ pos, const int64_t* __restrict__ batch, const T* __restrict__ box, int box_stride,
    int periodic, int n, T lo2, T hi2, int loop, int transpose, const int* __restrict__ seg_lo,
    const int* __restrict__ seg_hi, const int* __restrict__ unsorted_flag, const int* __restrict__ indptr,
    int max_pairs, int* __restrict__ deg, int* __restrict__ out_i, int* __restrict__ out_j, T* __restrict__ out_vec,
    T* __restrict__ out_dist, const int* __restrict__ cell_hdr, const int* __restrict__ boxhdr, const int* __restrict__ cell_of,
    const int* __restrict__ cell_start, const int* __restrict__ cell_atoms) {
  extern __shared__ unsigned char nb_smem[];
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int wid = threadIdx.x >> 5;
  const int nwarps = gridDim.x * warps_per_block;
  const bool unsorted = (*unsorted_flag) != 0;
  const bool use_cells = cell_hdr && cell_hdr[3] != 0;
  // per-warp staging (cell path, FILL only): j[ROW_CAP], then dx, dy, dz, d2 as T
  int* sj = reinterpret_cast<int*>(nb_smem) + wid * ROW_CAP;
  T* sv = reinterpret_cast<T*>(nb_smem + (size_t)warps_per_block * ROW_CAP * sizeof(int)) + (size_t)wid * ROW_CAP * 4;
  for (int i = blockIdx.x * warps_per_block + wid; i < n; i += nwarps) {
    const T xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
    const int64_t bi = batch[i];
    BoxRow<T> B = {0, 0, 0, 0, 0, 0};
    if (periodic) {
      const T* bx = box + (size_t)bi * box_stride;
      B.b00 = bx[0]; B.b10 = bx[3]; B.b11 = bx[4]; B.b20 = bx[6]; B.b21 = bx[7]; B.b22 = bx[8];
    }
    const int base = FILL ? indptr[i] : 0;
    int count = 0;
    bool row_done = false;
    if (use_cells) {
      // ------------------------------------------------------------------ cell path
      const int64_t bxi = box_stride ? bi : 0;          // the row's own box grid
      const int nx = boxhdr[4 * bxi], ny = boxhdr[4 * bxi + 1], nz = boxhdr[4 * bxi + 2], coff = boxhdr[4 * bxi + 3];
      const int ci = cell_of[i] - coff;
      const int cz = ci % nz, cy = (ci / nz) % ny, cx = ci / (nz * ny);
      for (int nb = 0; nb < 27; ++nb) {
        const int ax = (cx + nb / 9 - 1 + nx) % nx, ay = (cy + (nb / 3) % 3 - 1 + ny) % ny, az = (cz + nb % 3 - 1 + nz) % nz;
        const int c2 = coff + (ax * ny + ay) * nz + az;
        const int s0 = cell_start[c2], s1 = cell_start[c2 + 1];
        for (int t0 = s0; t0 < s1; t0 += 32) {
          const int t = t0 + lane;
          bool ok = false;
          int j = -1;
          T dx = 0, dy = 0, dz = 0, d2 = 0;
          if (t < s1) {
            j = cell_atoms[t];
            if (batch[j] == bi && (transpose || j <= i)) {
              if (j == i) ok = loop != 0;
              else ok = pair_delta(pos, i, j, xi, yi, zi, periodic, B, lo2, hi2, dx, dy, dz, d2);
            }
          }
          const unsigned m = __ballot_sync(0xffffffffu, ok);
          if (FILL && ok) {
            const int k = count + __popc(m & ((1u << lane) - 1u));
            if (k < ROW_CAP) { sj[k] = j; sv[4 * k] = dx; sv[4 * k + 1] = dy; sv[4 * k + 2] = dz; sv[4 * k + 3] = d2; }
          }
          count += __popc(m);
        }
      }
      if (!FILL) {
        row_done = true;
      } else if (count <= ROW_CAP) {
        __syncwarp();
        for (int e = lane; e < count; e += 32) {   // rank sort by j (unique within a row)
          const int j = sj[e];
          int rank = 0;
          for (int u = 0; u < count; ++u) rank += (sj[u] < j);
          const int slot = base + rank;
          if (slot < max_pairs) {
            out_i[slot] = i; out_j[slot] = j;
            out_vec[3 * (size_t)slot] = sv[4 * e]; out_vec[3 * (size_t)slot + 1] = sv[4 * e + 1]; out_vec[3 * (size_t)slot + 2] = sv[4 * e + 2];
            out_dist[slot] = sqrt_rn(sv[4 * e + 3]);
          }
        }
        __syncwarp();
        row_done = true;
      } else {
        count = 0;  // very dense row: emit it with the ordered segment scan below
      }
    }
    if (!row_done) {
      // ------------------------------------------------------------------ ordered scan of the batch segment
      int jlo = unsorted ? 0 : seg_lo[i];
      int jhi = unsorted ? n : seg_hi[i];
      if (!transpose) jhi = min(jhi, i + 1);  // half list keeps only j <= i (the canonical orientation)
      for (int j0 = jlo; j0 < jhi; j0 += 32) {
        const int j = j0 + lane;
        bool ok = false;
        T dx = 0, dy = 0, dz = 0, d2 = 0;
        if (j < jhi && batch[j] == bi) {
          if (j == i) ok = loop != 0;
          else ok = pair_delta(pos, i, j, xi, yi, zi, periodic, B, lo2, hi2, dx, dy, dz, d2);
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (FILL && ok) {
          const int slot = base + count + __popc(m & ((1u << lane) - 1u));
          if (slot < max_pairs) {
            out_i[slot] = i; out_j[slot] = j;
            out_vec[3 * (size_t)slot] = dx; out_vec[3 * (size_t)slot + 1] = dy; out_vec[3 * (size_t)slot + 2] = dz;
            out_dist[slot] = sqrt_rn(d2);
          }
        }
        count += __popc(m);
      }
    }
    if (!FILL && lane == 0) deg[i] = count;
  }
}

// single-CTA exclusive scan: out[0]=0, out[k+1]=sum_{t<=k} in[t]; total -> *total_out (optional)
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const int* __restrict__ in, int n, int* __restrict__ out,
                                                               int* __restrict__ total_out) {
  constexpr int ITEMS = 8;   // consecutive elements per thread and pass: 8192 per pass instead of 1024
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) { carry_s = 0; out[0] = 0; }
  __syncthreads();
  for (int base = 0; base < n; base += 1024 * ITEMS) {
    const int i0 = base + threadIdx.x * ITEMS;
    int v[ITEMS];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) { v[k] = (i0 + k < n) ? in[i0 + k] : 0; sum += v[k]; }
    int x = sum;
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
    const int incl = x + (w > 0 ? warp_tot[w - 1] : 0) + carry_s;   // inclusive prefix up to this thread's last element
    int run = incl - sum;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) { run += v[k]; if (i0 + k < n) out[i0 + k + 1] = run; }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

// rev[e] = slot of the transposed pair (j,i) (binary search in the sorted row j); -1 if absent
__global__ void reverse_edge_kernel(const int* __restrict__ indptr, const int* __restrict__ nbr_i,
                                    const int* __restrict__ nbr_j, const int* __restrict__ num_pairs, int max_pairs,
                                    int* __restrict__ rev) {
  const int E = min(*num_pairs, max_pairs);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    const int i = nbr_i[e], j = nbr_j[e];
    int lo = indptr[j], hi = min(indptr[j + 1], E);
    int r = -1;
    while (lo < hi) { int m = (lo + hi) >> 1; int v = nbr_j[m]; if (v < i) lo = m + 1; else if (v > i) hi = m; else { r = m; break; } }
    rev[e] = r;
  }
}

// a2: dpos_i = sum_{e in row i} (g_e - g_rev(e)),  g_e = dvec_e + ddist_e * vec_e / dist_e (0 where dist == 0)
// (reference: extensions/neighbor_utils.py:11-46, there as two index_add_ with atomics)
template <typename T>
__device__ __forceinline__ T warp_sum_t(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// one warp per atom: lanes stride over the row's slots (fixed summation tree -> deterministic), no atomics
template <typename T>
__global__ void __launch_bounds__(256) neighbor_grad_csr_kernel(const int* __restrict__ indptr, const int* __restrict__ rev,
                                                                const int* __restrict__ num_pairs, int max_pairs, int n,
                                                                const T* __restrict__ vec, const T* __restrict__ dist,
                                                                const T* __restrict__ gvec, const T* __restrict__ gdist,
                                                                T* __restrict__ gpos) {
  const int E = min(*num_pairs, max_pairs);
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    T ax = 0, ay = 0, az = 0;
    const int e1 = min(indptr[i + 1], E);
    for (int e = indptr[i] + lane; e < e1; e += 32) {
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        const int k = side == 0 ? e : rev[e];
        if (k < 0) continue;
        const T d = dist[k];
        if (d == T(0)) continue;
        const T s = gdist ? gdist[k] / d : T(0);
        T gx = s * vec[3 * (size_t)k], gy = s * vec[3 * (size_t)k + 1], gz = s * vec[3 * (size_t)k + 2];
        if (gvec) { gx += gvec[3 * (size_t)k]; gy += gvec[3 * (size_t)k + 1]; gz += gvec[3 * (size_t)k + 2]; }
        if (side == 0) { ax += gx; ay += gy; az += gz; } else { ax -= gx; ay -= gy; az -= gz; }
      }
    }
    ax = warp_sum_t(ax); ay = warp_sum_t(ay); az = warp_sum_t(az);
    if (lane == 0) { gpos[3 * (size_t)i] = ax; gpos[3 * (size_t)i + 1] = ay; gpos[3 * (size_t)i + 2] = az; }
  }
}

// generic (any edge order, -1 padded) position gradient, atomics; used by the half-list / unsorted API path
template <typename T>
__global__ void neighbor_grad_coo_kernel(const int64_t* __restrict__ nbr, long long P, const T* __restrict__ vec,
                                         const T* __restrict__ dist, const T* __restrict__ gvec,
                                         const T* __restrict__ gdist, T* __restrict__ gpos) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < P; e += (long long)gridDim.x * blockDim.x) {
    const int64_t i = nbr[e], j = nbr[P + e];
    if (i < 0) continue;
    const T d = dist[e];
    if (d == T(0)) continue;
    const T s = gdist ? gdist[e] / d : T(0);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      T g = s * vec[3 * e + c] + (gvec ? gvec[3 * e + c] : T(0));
      atomicAdd(gpos + 3 * i + c, g);
      atomicAdd(gpos + 3 * j + c, -g);
    }
  }
}

// ---- a3: COO -> CSR / CSC with per-row slots sorted by edge id (deterministic) -----------------
__global__ void degree_kernel(const int* __restrict__ ei, int E, int n, int* __restrict__ row_deg, int* __restrict__ col_deg) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    const int r = ei[e], c = ei[E + e];
    if (r < 0 || c < 0 || r >= n || c >= n) continue;
    atomicAdd(row_deg + r, 1); atomicAdd(col_deg + c, 1);
  }
}
__global__ void csr_fill_kernel(const int* __restrict__ ei, int E, int n, const int* __restrict__ row_ptr,
                                const int* __restrict__ col_ptr, int* __restrict__ row_cur, int* __restrict__ col_cur,
                                int* __restrict__ row_data, int* __restrict__ col_data) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    const int r = ei[e], c = ei[E + e];
    if (r < 0 || c < 0 || r >= n || c >= n) continue;
    row_data[row_ptr[r] + atomicAdd(row_cur + r, 1)] = e;
    col_data[col_ptr[c] + atomicAdd(col_cur + c, 1)] = e;
  }
}
// sort each row's edge ids ascending (rows are short), then fill the neighbour index
__global__ void csr_sort_rows_kernel(const int* __restrict__ ei, int E, int n, const int* __restrict__ ptr,
                                     int* __restrict__ data, int* __restrict__ indices, int other_row) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const int s = ptr[r], t = ptr[r + 1];
    for (int a = s + 1; a < t; ++a) {
      const int v = data[a];
      int b = a - 1;
      while (b >= s && data[b] > v) { data[b + 1] = data[b]; --b; }
      data[b + 1] = v;
    }
    for (int a = s; a < t; ++a) indices[a] = ei[other_row * E + data[a]];
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------ C ABI
template <typename T>
static int neighbor_csr_impl(const T* pos, const int64_t* batch, const T* box, int box_stride, int periodic, int n,
                             double cutoff_lower, double cutoff_upper, int loop, int transpose, int max_pairs,
                             int max_cells, int n_boxes, int* indptr, int* nbr_i, int* nbr_j, T* vec, T* dist, int* num_pairs,
                             int* rev, int* workspace, cudaStream_t st) {
  TMD_REQUIRE(n >= 0 && max_pairs >= 0 && max_cells >= 0, "neighbor_csr: negative size");
  TMD_REQUIRE(pos && batch && indptr && nbr_i && nbr_j && vec && dist && num_pairs && workspace,
              "neighbor_csr: null pointer");
  TMD_REQUIRE(!periodic || box, "neighbor_csr: periodic without box");
  if (n == 0) { cudaMemsetAsync(num_pairs, 0, sizeof(int), st); cudaMemsetAsync(indptr, 0, sizeof(int), st); return tmd_check_launch("neighbor_csr(empty)"); }
  // workspace (ints): seg_lo[n] seg_hi[n] deg[n] flag[1] | hdr[4] cell_of[n] cell_atoms[n] cell_count[C+1] cell_start[C+1] cursor[C+1]
  //                   boxhdr[4 n_boxes]
  int* seg_lo = workspace; int* seg_hi = workspace + n; int* deg = workspace + 2 * n; int* flag = workspace + 3 * n;
  int* hdr = workspace + 3 * n + 1; int* cell_of = hdr + 4; int* cell_atoms = cell_of + n;
  int* cell_count = cell_atoms + n; int* cell_start = cell_count + (max_cells + 1); int* cursor = cell_start + (max_cells + 1);
  int* boxhdr = cursor + (max_cells + 1);
  TMD_REQUIRE(n_boxes >= 0 && (box_stride == 0 ? n_boxes <= 1 : true), "neighbor_csr: n_boxes must be the number of box matrices");
  const bool cells = max_cells >= 27 && periodic && n_boxes >= 1 && n_boxes <= 4096;
  cudaMemsetAsync(flag, 0, sizeof(int) * 5, st);  // flag + hdr (hdr.ok = 0 unless the setup kernel enables the grid)
  batch_segments_kernel<<<tmd_blocks_for(n, 256), 256, 0, st>>>(batch, n, seg_lo, seg_hi, flag);
  if (cells) {
    cudaMemsetAsync(cell_count, 0, sizeof(int) * 3 * (max_cells + 1), st);
    const bool f32 = sizeof(T) == 4;
    cell_setup_kernel<<<1, 1, 0, st>>>(f32 ? reinterpret_cast<const float*>(box) : nullptr,
                                       f32 ? nullptr : reinterpret_cast<const double*>(box), box_stride, n_boxes, periodic,
                                       cutoff_upper, max_cells, hdr, boxhdr);
    cell_assign_kernel<T><<<tmd_blocks_for(n, 256), 256, 0, st>>>(pos, batch, box, box_stride, n, hdr, boxhdr, cell_of, cell_count);
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(cell_count, max_cells, cell_start, nullptr);
    cell_fill_kernel<<<tmd_blocks_for(n, 256), 256, 0, st>>>(n, hdr, cell_of, cell_start, cursor, cell_atoms);
  }
  // cut-offs are squared in double and rounded to T, like the reference's python-float kernel args
  const T lo2 = (T)(cutoff_lower * cutoff_lower), hi2 = (T)(cutoff_upper * cutoff_upper);
  const int blocks = tmd_blocks_for((long long)n * 32, 256, 148 * 8);
  const size_t smem = cells ? (size_t)8 * ROW_CAP * (sizeof(int) + 4 * sizeof(T)) : 0;
  if (smem > 48 * 1024) {
    cudaFuncSetAttribute(neighbor_rows_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(neighbor_rows_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  const int* chdr = cells ? hdr : nullptr;
  neighbor_rows_kernel<T, false><<<blocks, 256, smem, st>>>(pos, batch, box, box_stride, periodic, n, lo2, hi2, loop,
                                                             transpose, seg_lo, seg_hi, flag, nullptr, max_pairs, deg,
                                                             nullptr, nullptr, nullptr, nullptr, chdr, boxhdr, cell_of, cell_start, cell_atoms);
  exclusive_scan_kernel<<<1, 1024, 0, st>>>(deg, n, indptr, num_pairs);
  neighbor_rows_kernel<T, true><<<blocks, 256, smem, st>>>(pos, batch, box, box_stride, periodic, n, lo2, hi2, loop,
                                                            transpose, seg_lo, seg_hi, flag, indptr, max_pairs, deg,
                                                            nbr_i, nbr_j, vec, dist, chdr, boxhdr, cell_of, cell_start, cell_atoms);
  if (rev) reverse_edge_kernel<<<tmd_blocks_for(max_pairs, 256), 256, 0, st>>>(indptr, nbr_i, nbr_j, num_pairs, max_pairs, rev);
  return tmd_check_launch("neighbor_csr");
}

extern "C" {

int tmd_neighbor_csr_f32(const float* pos, const int64_t* batch, const float* box, int box_stride, int periodic,
                         int n, double cutoff_lower, double cutoff_upper, int loop, int transpose, int max_pairs,
                         int max_cells, int n_boxes, int* indptr, int* nbr_i, int* nbr_j, float* vec, float* dist, int* num_pairs,
                         int* rev, int* workspace, void* stream) {
  return neighbor_csr_impl<float>(pos, batch, box, box_stride, periodic, n, cutoff_lower, cutoff_upper, loop,
                                  transpose, max_pairs, max_cells, n_boxes, indptr, nbr_i, nbr_j, vec, dist, num_pairs, rev,
                                  workspace, (cudaStream_t)stream);
}
int tmd_neighbor_csr_f64(const double* pos, const int64_t* batch, const double* box, int box_stride, int periodic,
                         int n, double cutoff_lower, double cutoff_upper, int loop, int transpose, int max_pairs,
                         int max_cells, int n_boxes, int* indptr, int* nbr_i, int* nbr_j, double* vec, double* dist, int* num_pairs,
                         int* rev, int* workspace, void* stream) {
  return neighbor_csr_impl<double>(pos, batch, box, box_stride, periodic, n, cutoff_lower, cutoff_upper, loop,
                                   transpose, max_pairs, max_cells, n_boxes, indptr, nbr_i, nbr_j, vec, dist, num_pairs, rev,
                                   workspace, (cudaStream_t)stream);
}

int tmd_neighbor_grad_csr_f32(const int* indptr, const int* rev, const int* num_pairs, int max_pairs, int n,
                              const float* vec, const float* dist, const float* gvec, const float* gdist,
                              float* gpos, void* stream) {
  TMD_REQUIRE(indptr && rev && num_pairs && vec && dist && gpos, "neighbor_grad_csr: null pointer");
  if (n == 0) return TMD_OK;
  neighbor_grad_csr_kernel<float><<<tmd_blocks_for((long long)n * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      indptr, rev, num_pairs, max_pairs, n, vec, dist, gvec, gdist, gpos);
  return tmd_check_launch("neighbor_grad_csr");
}
int tmd_neighbor_grad_csr_f64(const int* indptr, const int* rev, const int* num_pairs, int max_pairs, int n,
                              const double* vec, const double* dist, const double* gvec, const double* gdist,
                              double* gpos, void* stream) {
  TMD_REQUIRE(indptr && rev && num_pairs && vec && dist && gpos, "neighbor_grad_csr: null pointer");
  if (n == 0) return TMD_OK;
  neighbor_grad_csr_kernel<double><<<tmd_blocks_for((long long)n * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      indptr, rev, num_pairs, max_pairs, n, vec, dist, gvec, gdist, gpos);
  return tmd_check_launch("neighbor_grad_csr");
}
int tmd_neighbor_grad_coo_f32(const int64_t* nbr, long long P, const float* vec, const float* dist, const float* gvec,
                              const float* gdist, float* gpos_zeroed, void* stream) {
  TMD_REQUIRE(nbr && vec && dist && gpos_zeroed, "neighbor_grad_coo: null pointer");
  if (P == 0) return TMD_OK;
  neighbor_grad_coo_kernel<float><<<tmd_blocks_for(P, 256), 256, 0, (cudaStream_t)stream>>>(nbr, P, vec, dist, gvec, gdist, gpos_zeroed);
  return tmd_check_launch("neighbor_grad_coo");
}
int tmd_neighbor_grad_coo_f64(const int64_t* nbr, long long P, const double* vec, const double* dist,
                              const double* gvec, const double* gdist, double* gpos_zeroed, void* stream) {
  TMD_REQUIRE(nbr && vec && dist && gpos_zeroed, "neighbor_grad_coo: null pointer");
  if (P == 0) return TMD_OK;
  neighbor_grad_coo_kernel<double><<<tmd_blocks_for(P, 256), 256, 0, (cudaStream_t)stream>>>(nbr, P, vec, dist, gvec, gdist, gpos_zeroed);
  return tmd_check_launch("neighbor_grad_coo");
}

// a3: edge_index int32 [2,E] (entries < 0 are padding) -> CSR(row = edge_index[0]) and CSC(col = edge_index[1]).
// workspace: 2*(n+1) ints, zeroed by this call.
int tmd_graph_transform(const int* edge_index, int E, int n, int* row_data, int* row_indices, int* row_indptr,
                        int* col_data, int* col_indices, int* col_indptr, int* workspace, void* stream) {
  TMD_REQUIRE(E >= 0 && n >= 0, "graph_transform: negative size");
  TMD_REQUIRE(row_indptr && col_indptr && workspace, "graph_transform: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int* rdeg = workspace; int* cdeg = workspace + (n + 1);
  cudaMemsetAsync(workspace, 0, sizeof(int) * 2 * (n + 1), st);
  if (E > 0) degree_kernel<<<tmd_blocks_for(E, 256), 256, 0, st>>>(edge_index, E, n, rdeg, cdeg);
  exclusive_scan_kernel<<<1, 1024, 0, st>>>(rdeg, n, row_indptr, nullptr);
  exclusive_scan_kernel<<<1, 1024, 0, st>>>(cdeg, n, col_indptr, nullptr);
  cudaMemsetAsync(workspace, 0, sizeof(int) * 2 * (n + 1), st);
  if (E > 0) {
    csr_fill_kernel<<<tmd_blocks_for(E, 256), 256, 0, st>>>(edge_index, E, n, row_indptr, col_indptr, rdeg, cdeg, row_data, col_data);
    csr_sort_rows_kernel<<<tmd_blocks_for(n, 128), 128, 0, st>>>(edge_index, E, n, row_indptr, row_data, row_indices, 1);
    csr_sort_rows_kernel<<<tmd_blocks_for(n, 128), 128, 0, st>>>(edge_index, E, n, col_indptr, col_data, col_indices, 0);
  }
  return tmd_check_launch("graph_transform");
}

}  // extern "C"
