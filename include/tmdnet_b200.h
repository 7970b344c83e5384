/*
 * libtmdnet_b200.so -- C ABI of the B200-native (sm_100a) TensorNet / ET energy+force hot path.
 *
 * This is the drop-in boundary that replaces torchmd-net's `torchmdnet/extensions` (NVIDIA-Warp kernels
 * wrapped as torch custom ops).  Conventions, shared by every entry point:
 *   - plain pointers + sizes, no torch / Warp types; all pointers are DEVICE pointers;
 *   - the CALLER allocates every output and workspace (the reference's op wrappers do the same with
 *     torch.empty/zeros, e.g. warp_ops/neighbors.py:69-72, warp_ops/tensornet_mp.py:57-59); kernels never
 *     allocate or free; inputs are read-only;
 *   - `stream` is a cudaStream_t; every call is an asynchronous launch with no host synchronisation, hence
 *     CUDA-graph capturable (reference: wp.launch(..., stream=torch stream), warp_kernels/utils.py:84-88);
 *   - return value: 0 on success, <0 for invalid arguments / launch errors (text via tmd_last_error());
 *     device-side conditions (neighbour-list overflow) are reported through an output counter, exactly like
 *     the reference's `num_pairs` (models/utils.py:297-300);
 *   - fp32 tensors are F-last and contiguous: X [N,3,3,F]; I [N,1,F], A [N,3,F], S [N,5,F] with component
 *     order A = (01,02,12), S = (00,01,02,11,12), S22 = -(S00+S11) (warp_kernels/compose_tensor.py:54-98).
 *     F must be a multiple of 4 for the message-passing entry points (128-bit loads).
 *
 * Reference paths below are relative to torchmd/torchmd-net @ 833c0770.
 */
#ifndef TMDNET_B200_H
#define TMDNET_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int tmd_abi_version(void);
const char* tmd_last_error(void);

/* ---- a1: neighbour list ---------------------------------------------------------------------------------
 * Replaces torch.ops.torchmdnet.warp_neighbor_brute_fwd / warp_neighbor_cell_fwd
 * (extensions/warp_ops/neighbors.py:34-148, :196-289; kernels warp_kernels/neighbors_brute.py:32-205,
 * neighbors_cell.py:17-147).  Same pair semantics (same `batch`, cutoff_lower <= r < cutoff_upper compared
 * squared, self pairs bypass the lower cutoff, triclinic minimum image by sequential z,y,x rounding, canonical
 * orientation = larger index first, transposed pair negated).  Output is sorted by (i, j) and therefore IS a
 * CSR: indptr[n+1], nbr_i/nbr_j/vec/dist[max_pairs].  num_pairs[0] = pairs found (may exceed max_pairs; slots
 * past max_pairs are dropped).  Slots >= num_pairs are left untouched (caller pre-fills -1 / 0).
 * box: [1 or n_batch,3,3]; box_stride = 0 (shared) or 9 (per sample).  rev (optional, may be NULL): slot of the
 * transposed pair.  transpose=0 gives the half list (j <= i).  max_cells > 0 enables the O(N) cell-grid candidate
 * search: one grid per periodic rectangular box with >= 3 cells per axis (n_boxes = number of box matrices, 1 for a shared
 * box; the reference's cell kernel handles a single box only, warp_ops/neighbors.py:222-229); decided on the device, identical
 * output; any other input (triclinic, small or non-periodic boxes, more cells than max_cells) uses the ordered batch-segment
 * scan.  workspace: 5*n + 8 + 3*(max_cells+1) + 4*n_boxes int32. */
int tmd_neighbor_csr_f32(const float* pos, const int64_t* batch, const float* box, int box_stride, int periodic,
                         int n, double cutoff_lower, double cutoff_upper, int loop, int transpose, int max_pairs,
                         int max_cells, int n_boxes, int* indptr, int* nbr_i, int* nbr_j, float* vec, float* dist, int* num_pairs,
                         int* rev, int* workspace, void* stream);
int tmd_neighbor_csr_f64(const double* pos, const int64_t* batch, const double* box, int box_stride, int periodic,
                         int n, double cutoff_lower, double cutoff_upper, int loop, int transpose, int max_pairs,
                         int max_cells, int n_boxes, int* indptr, int* nbr_i, int* nbr_j, double* vec, double* dist, int* num_pairs,
                         int* rev, int* workspace, void* stream);

/* ---- a2: position gradient of the neighbour outputs -------------------------------------------------------
 * Replaces neighbor_grad_positions (extensions/neighbor_utils.py:11-46; two index_add_ with atomics).
 * _csr: deterministic gather over the symmetric CSR (needs rev); _coo: any -1 padded int64 [2,P] list, atomics,
 * gpos must be zero-initialised.  gvec / gdist may be NULL. */
int tmd_neighbor_grad_csr_f32(const int* indptr, const int* rev, const int* num_pairs, int max_pairs, int n,
                              const float* vec, const float* dist, const float* gvec, const float* gdist,
                              float* gpos, void* stream);
int tmd_neighbor_grad_csr_f64(const int* indptr, const int* rev, const int* num_pairs, int max_pairs, int n,
                              const double* vec, const double* dist, const double* gvec, const double* gdist,
                              double* gpos, void* stream);
int tmd_neighbor_grad_coo_f32(const int64_t* neighbors, long long P, const float* vec, const float* dist,
                              const float* gvec, const float* gdist, float* gpos_zeroed, void* stream);
int tmd_neighbor_grad_coo_f64(const int64_t* neighbors, long long P, const double* vec, const double* dist,
                              const double* gvec, const double* gdist, double* gpos_zeroed, void* stream);

/* ---- a3: COO -> CSR + CSC ------------------------------------------------------------------------------------
 * Replaces nvtnet::count_row_col_primitive + nvtnet::convert_to_sparse_primitive + 2 torch.cumsum
 * (extensions/warp_ops/graph_transform.py:160-179; warp_kernels/graph_transform.py:33-79).
 * edge_index int32 [2,E], negative entries are padding.  Slots inside a row are ordered by edge id
 * (deterministic; the reference's order is atomic-dependent).  workspace: 2*(n+1) int32. */
int tmd_graph_transform(const int* edge_index, int E, int n, int* row_data, int* row_indices, int* row_indptr,
                        int* col_data, int* col_indices, int* col_indptr, int* workspace, void* stream);

/* ---- a5: embedding ("radial") message passing -----------------------------------------------------------------
 * Replaces tensornet::radial_message_passing_{fwd,bwd}_primitive (warp_ops/tensornet_radial_mp.py:37-191;
 * kernels warp_kernels/tensornet_radial_mp.py:43-189).  rhat [E,3], edge_attr [E,3,F]; CSR row_indptr[n+1] with
 * optional edge-id indirection row_data (NULL: slot k is edge k).  The sign of A follows the reference's PyTorch
 * path (models/tensornet.py:84-103).  Backward has one writer per edge (no fp atomics). */
int tmd_radial_mp_fwd(const float* rhat, const float* edge_attr, const int* row_data, const int* row_indptr, int n,
                      int F, float* out_I, float* out_A, float* out_S, void* stream);
int tmd_radial_mp_bwd(const float* rhat, const float* edge_attr, const int* row_data, const int* row_indptr, int n,
                      int F, const float* dI, const float* dA, const float* dS, float* d_edge_attr, float* d_rhat,
                      void* stream);

/* ---- a7: interaction message passing ----------------------------------------------------------------------------
 * Replaces tensornet::message_passing_{fwd,bwd}_primitive (warp_ops/tensornet_mp.py:37-210; kernels
 * warp_kernels/tensornet_mp.py:42-131).  fwd gathers over CSR rows, bwd over the CSC columns (d_edge_attr may be
 * NULL when the edge weights need no gradient). */
int tmd_mp_fwd(const float* I, const float* A, const float* S, const float* edge_attr, const int* row_data,
               const int* row_indices, const int* row_indptr, int n, int F, float* out_I, float* out_A, float* out_S,
               void* stream);
int tmd_mp_bwd(const float* I, const float* A, const float* S, const float* edge_attr, const float* dout_I,
               const float* dout_A, const float* dout_S, const int* col_data, const int* col_indices,
               const int* col_indptr, int n, int F, float* dI, float* dA, float* dS, float* d_edge_attr, void* stream);

/* ---- a8/a10: per-(atom,channel) tensor algebra, the reference's granular op set ----------------------------------
 * Replace tensornet::{compose_tensor,decompose_tensor,tensor_matmul_o3_3x3,tensor_matmul_so3_3x3,tensor_norm3}
 * _{fwd,bwd}_primitive (warp_ops/*.py; kernels warp_kernels/compose_tensor.py:54-148, decompose_tensor.py:56-140,
 * equivariant_o3_matmul.py:42-101, equivariant_so3_matmul.py:42-99, tensor_norm3.py:48-138). */
int tmd_compose_fwd(const float* I, const float* A, const float* S, float* X, int N, int F, void* stream);
int tmd_compose_bwd(const float* dX, float* dI, float* dA, float* dS, int N, int F, void* stream);
int tmd_decompose_fwd(const float* X, float* I, float* A, float* S, int N, int F, void* stream);
int tmd_decompose_bwd(const float* dI, const float* dA, const float* dS, float* dX, int N, int F, void* stream);
int tmd_matmul_o3_fwd(const float* A, const float* B, float* C, int N, int F, void* stream);
int tmd_matmul_o3_bwd(const float* A, const float* B, const float* dC, float* dA, float* dB, int N, int F, void* stream);
int tmd_matmul_so3_fwd(const float* A, const float* B, float* C, int N, int F, void* stream);
int tmd_matmul_so3_bwd(const float* A, const float* B, const float* dC, float* dA, float* dB, int N, int F, void* stream);
int tmd_norm3_fwd(const float* X, float* out, int N, int F, void* stream);
int tmd_norm3_bwd(const float* dout, const float* X, float* dX, int N, int F, void* stream);

/* ---- a8/a10 fused: compact-irreps stages used by the B200 model path -----------------------------------------------
 * One kernel each for what models/tensornet.py expresses as chains of the granular ops:
 *   norm2c : |X|^2 [N,F]                          (tensornet.py:589, tensor_norm of the composed tensor)
 *   norm3c : (3I^2,|A|^2,|S|^2) [N,3F]            (tensornet.py:384-398)
 *   gate   : (I g0, A g1, S g2), g [N,3,F]        (tensornet.py:607-615)
 *   p1     : X/(|X|^2+1)                          (tensornet.py:745)
 *   p2     : decompose(C)/(|C|^2+1), C = qf (YM+MY) [o3=1] or 2 YM [o3=0]     (tensornet.py:786-806)
 *   p3     : X + dX + qf dX.dX                    (tensornet.py:812)
 * qf [N] = 1 + 0.1 q_atom. */
int tmd_norm2c_fwd(const float* I, const float* A, const float* S, float* out, int N, int F, void* stream);
int tmd_norm2c_bwd(const float* I, const float* A, const float* S, const float* dout, float* dI, float* dA, float* dS, int N, int F, void* stream);
int tmd_norm3c_fwd(const float* I, const float* A, const float* S, float* out, int N, int F, void* stream);
int tmd_norm3c_bwd(const float* I, const float* A, const float* S, const float* dout, float* dI, float* dA, float* dS, int N, int F, void* stream);
int tmd_gate_fwd(const float* I, const float* A, const float* S, const float* g, float* oI, float* oA, float* oS, int N, int F, void* stream);
int tmd_gate_bwd(const float* I, const float* A, const float* S, const float* g, const float* dI, const float* dA, const float* dS,
                 float* gI, float* gA, float* gS, float* dg, int N, int F, void* stream);
int tmd_p1_fwd(const float* I, const float* A, const float* S, float* oI, float* oA, float* oS, int N, int F, void* stream);
int tmd_p1_bwd(const float* I, const float* A, const float* S, const float* dI, const float* dA, const float* dS,
               float* gI, float* gA, float* gS, int N, int F, void* stream);
int tmd_p2_fwd(const float* YI, const float* YA, const float* YS, const float* MI, const float* MA, const float* MS,
               const float* qf, int o3, float* oI, float* oA, float* oS, int N, int F, void* stream);
int tmd_p2_bwd(const float* YI, const float* YA, const float* YS, const float* MI, const float* MA, const float* MS,
               const float* qf, int o3, const float* dI, const float* dA, const float* dS, float* gYI, float* gYA,
               float* gYS, float* gMI, float* gMA, float* gMS, int N, int F, void* stream);
int tmd_p3_fwd(const float* XI, const float* XA, const float* XS, const float* DI, const float* DA, const float* DS,
               const float* qf, float* oI, float* oA, float* oS, int N, int F, void* stream);
int tmd_p3_bwd(const float* DI, const float* DA, const float* DS, const float* qf, const float* dI, const float* dA,
               const float* dS, float* gXI, float* gXA, float* gXS, float* gDI, float* gDA, float* gDS, int N, int F, void* stream);

/* ---- a4+a5+a6+a7 fused: radial pipeline on unique pairs -----------------------------------------------------------
 * Replaces, for inference+forces, the reference's per-edge chain ExpNormalSmearing -> CosineCutoff -> 3x(Linear+SiLU)
 * (models/tensornet.py:738-743; models/utils.py:356-407,500-528) and the embedding's radial projections
 * (tensornet.py:558-567), which the reference runs as cuBLAS SGEMMs + elementwise launches on every directed edge.
 *   tmd_pair_index : unique pairs of the symmetric CSR (canonical edge j <= i): pair_ptr[n+1], pid[E] (edge -> pair),
 *                    pair_edge[P], pair_r[P], num_unique[1].  workspace: n int32.
 *   tmd_pair_chain : out_f[p,:] = chain(rbf(r_p)) (x cutoff), out_fp[p,:] = d/dr of it (NULL: value only).
 *                    WTl: transposed Linear weight [K_l][N_l]; dims = {R, N_0, N_1, N_2}; N_l % 32 == 0, <= 384.
 *   tmd_mp_pair_bwd: backward of tmd_mp_fwd with pair-indexed weights: dY and (if fp != NULL) dE/dr_p = <dE/df_p, f'_p>
 *                    fused into the transposed gather -- no [E,3F] gradient, no backward GEMM.
 *   tmd_embed_fwd/bwd: embedding message passing with w = C(r) (U_i + V_j) dp_p computed in-kernel (U, V per atom:
 *                    emb2([Z_i;Z_j]) = U_i + V_j); bwd returns dE/d dist[e], dE/d rhat[e,3]. */
int tmd_pair_index(const int* indptr, const int* nbr_i, const int* nbr_j, const int* rev, const int* num_edges,
                   int max_edges, int n, const float* dist, int* pair_ptr, int* pid, int* pair_edge, float* pair_r,
                   int* num_unique, int* workspace, void* stream);
int tmd_pair_chain(const float* pair_r, const int* num_unique, int max_pairs, const float* WT0, const float* b0,
                   const float* WT1, const float* b1, const float* WT2, const float* b2, const int* dims, int n_layers,
                   int act, int mul_cutoff, const float* means, const float* betas, double alpha, double lo, double hi,
                   float* out_f, float* out_fp, void* stream);
/* Tensor-core version of tmd_pair_chain: every layer is a 3xTF32 GEMM on tcgen05 (TMA-fed, TMEM accumulators).
 * W_hi/W_lo[l]: TF32 (cvt.rna) split of the Linear weight [N_l, K_l] with K_0 zero-padded to a multiple of 32;
 * hidden widths % 32 == 0, every N_l % 64 == 0.  Workspaces act_{a,b}_{hi,lo}: >= roundup(max_pairs*(out_fp?2:1),128)
 * x max(K_l) floats each.  Same outputs as tmd_pair_chain to ~1e-6 relative.
 * use_f16 = 1: the split terms are fp16 instead of tf32 (same 11-bit significands, half the bytes, twice the MMA rate;
 * W_hi/W_lo and the workspaces then hold __half, K padded to multiples of 64, hidden widths % 64 == 0).  fp16 has a
 * 5-bit exponent: the caller must bound |activations| and |d/dr| below 65504 (extensions/fused.py does, from the weight
 * norms, and falls back to tf32 otherwise). */
int tmd_pair_chain_tc(const float* pair_r, const int* num_unique, int max_pairs, const float* const* W_hi,
                      const float* const* W_lo, const float* const* bias, const int* dims, int n_layers, int act,
                      int mul_cutoff, const float* means, const float* betas, double alpha, double lo, double hi,
                      const float* rbf_hi, const float* rbf_lo, float* act_a_hi, float* act_a_lo, float* act_b_hi,
                      float* act_b_lo, float* out_f, float* out_fp, int use_f16, void* stream);
/* ---- 8(f).1: tabulated radial function ---------------------------------------------------------------------------------
 * SURVEY 8(f) item 1 ("tabulating each layer's radial function r -> MLP(RBF(r)) C(r) ... gated by a parity check"): for frozen
 * weights the chain of tmd_pair_chain* is a fixed R -> R^N function of one scalar (models/tensornet.py:738-743).  table is
 * [n_int][4][N] fp32 = (f_i, f_{i+1}-f_i, h f'_i, h f'_{i+1}) on the intervals [i h, (i+1) h), h = 2^-k (nodes exact in fp32),
 * built on the host in fp64 and accepted only if it reproduces the exact path (extensions/fused.py: RadialChain.table).
 * Writes out_f [P,N] and, if out_fp != NULL, d/dr -- same contract as tmd_pair_chain_tc. */
int tmd_radial_table(const float* pair_r, const int* num_unique, int max_pairs, const float* table, int n_int, double h, int N,
                     float* out_f, float* out_fp, void* stream);
/* ---- a9: node linears -----------------------------------------------------------------------------------------------
 * Replaces the three bias-free nn.Linear(F,F) that TensorNet applies to I / A / S (models/tensornet.py:752-754, :808-810,
 * :607-615; cuBLAS SGEMMs in the reference) by one grouped 3xTF32 tcgen05 GEMM that reads the fp32 irreps directly (the
 * (hi, lo) split of the A operand happens in shared memory) and writes oI [N,1,F], oA [N,3,F], oS [N,5,F].
 * W_hi/W_lo: TF32 split of the stacked weights [3F,F].  F % 64 == 0, <= 256. */
int tmd_irreps_linear_tc(const float* I, const float* A, const float* S, int N, int F, const float* W_hi, const float* W_lo,
                         float* oI, float* oA, float* oS, void* stream);
/* layer-0 operand of tmd_pair_chain_tc: rows (value[, d/dr]) of ExpNormalSmearing(r_p), TF32 split, K padded to a multiple
 * of 32.  Computed once per frame and shared by the embedding chain and every interaction layer (rbf_hi / rbf_lo). */
int tmd_pair_rbf_split(const float* pair_r, const int* num_unique, int max_pairs, int R, int rows_per_pair,
                       const float* means, const float* betas, double alpha, double lo, double hi, float* a_hi,
                       float* a_lo, int use_f16, void* stream);
int tmd_mp_pair_bwd(const float* YI, const float* YA, const float* YS, const float* dMI, const float* dMA,
                    const float* dMS, const float* f, const float* fp, const int* pid, const int* nbr_j,
                    const int* indptr, int n, int F, float* dYI, float* dYA, float* dYS, float* d_pair_r, void* stream);
/* single-pass variant: the radial gradient is emitted per DIRECTED slot, g_dist[k] = <f'_p, dM_j * Y_b> for k = (b, j); the two
 * slots of a pair together carry dE/dr_p (the position-gradient kernel combines them).  Gathers only dM_j. */
int tmd_mp_edge_bwd(const float* YI, const float* YA, const float* YS, const float* dMI, const float* dMA,
                    const float* dMS, const float* f, const float* fp, const int* pid, const int* nbr_j,
                    const int* indptr, int n, int F, float* dYI, float* dYA, float* dYS, float* g_dist, void* stream);
/* ---- a7, cooperative form (csrc/mp_coop.cu): the same message passing (tensornet_mp.py:42-131) with the neighbour rows of a
 * GROUP of up to tmd_mp_group_size() (15) spatially adjacent atoms staged once per group in shared memory (bulk async copies into an mbarrier ring) and the
 * per-pair radial rows streamed through per-warp async rings.  Bit-identical to tmd_mp_fwd / tmd_mp_edge_bwd.
 * tmd_mp_plan builds the per-frame plan shared by all launches of a frame: rows[n] (atom of slot r; group g = rows[gs g .. gs g + gs),
 * gs = rows per group, 1 .. tmd_mp_group_size(); tmd_mp_pick_group(n) returns the size that fills the 148 one-CTA-per-SM slots best:
 * `order` when given and *order_flag != 0, else 0..n-1), union_j / ucount [max_pairs] (sorted distinct neighbours of each group
 * and the number of its rows containing them, at goff[g] .. goff[g] + gcnt[g]), upos[max_pairs] (position of nbr_j[k] in its
 * group's union, bit 15 = self loop), goff / gcnt [ceil(n / gs)], gcnt[g] = -1 for a group over the staged capacity (it is gathered
 * directly).  cursor: one int of scratch.  F % 4 == 0, F <= 128.  The same gs goes to the plan and to the two kernels. */
int tmd_mp_group_size(void);
int tmd_mp_pick_group(int n);
int tmd_mp_plan(const int* indptr, const int* nbr_j, const int* num_pairs, int max_pairs, int n, int gs, const int* order,
                const int* order_flag, int* rows, int* cursor, int* union_j, unsigned char* ucount, unsigned short* upos,
                int* goff, int* gcnt, void* stream);
int tmd_mp_coop_fwd(const float* YI, const float* YA, const float* YS, const float* f, const int* pid, const int* nbr_j,
                    const int* indptr, const int* num_pairs, int max_pairs, int n, int F, int gs, const int* rows, const int* union_j,
                    const unsigned char* ucount, const unsigned short* upos, const int* goff, const int* gcnt, float* MI,
                    float* MA, float* MS, void* stream);
int tmd_mp_coop_bwd(const float* YI, const float* YA, const float* YS, const float* dMI, const float* dMA, const float* dMS,
                    const float* f, const float* fp, const int* pid, const int* nbr_j, const int* indptr, const int* num_pairs,
                    int max_pairs, int n, int F, int gs, const int* rows, const int* union_j, const unsigned char* ucount,
                    const unsigned short* upos, const int* goff, const int* gcnt, const float* addI, const float* addA,
                    const float* addS, float* dYI, float* dYA, float* dYS, float* g_dist, void* stream);
/* (addI, addA, addS): optional (all NULL or all given) gradient that reached Y along another path (the P2 stage); the kernel stores
 * dY + add, which replaces the separate accumulation launches of an autograd chain. */
/* embedding message passing on pair-indexed radial projections dp[P_u,3F] (and their d/dr dpp): the unit vector is formed in the
 * kernel from (vec, dist) of the neighbour list (0 for the self loop, tensornet.py:363-366); U, V are per-atom rows [n,F] when
 * zidx == NULL, else per-element tables [max_z,F] indexed through the atomic numbers zidx[n] (U_i = U[z_i], V_j = V[z_j]).
 * Backward w.r.t. geometry: g_dist[E] and g_vec[E,3]. */
int tmd_embed_fwd(const float* vec, const float* dist, const float* dp, const int* pid, const int* nbr_j,
                  const int* indptr, const float* U, const float* V, const int64_t* zidx, double lo, double hi, int n, int F,
                  float* oI, float* oA, float* oS, void* stream);
int tmd_embed_bwd(const float* vec, const float* dist, const float* dp, const float* dpp, const int* pid,
                  const int* nbr_j, const int* indptr, const float* U, const float* V, const int64_t* zidx, double lo, double hi,
                  int n, int F, const float* dI, const float* dA, const float* dS, float* g_dist, float* g_vec, void* stream);

/* ---- a5 by radial moments (csrc/embed_moments.cu) ---------------------------------------------------------------------------
 * The embedding's edge weight C(r)(U[z_i]+V[z_j])(W_g rbf(r) + b_g) (models/tensornet.py:554-585) is linear in the radial basis,
 * so the neighbour sum is taken before the linear map: per atom and neighbour element v the moments
 * m[v][k][c] = sum_e C phi_c(rhat) rbf_k(r) (c = 1 + 3 + 5 angular factors, k = R is the bias) and then
 * X[c][h] = sum_v (U+V_v)[h] sum_k W~[h][k] m[v][k][c].  No per-pair radial rows are read or written.
 * zslot[n]: slot (0..Z-1) of every atom's element; Uz, Vz [Z, F]: element-table rows of the Z species; Wk [3, R+1, F] (k-major,
 * row R = bias), Wh [3, F/4, R, 4] (Wh[g][h/4][k][h%4] = W_g[h][k]); means/betas [R], alpha, lo, hi: ExpNormalSmearing / CosineCutoff parameters
 * (models/utils.py:356-407, 500-528).  bwd returns dE/d dist[e] and dE/d vec[e,3] (buffers zero-initialised by the caller). */
int tmd_embed_moments_fwd(const float* vec, const float* dist, const int* nbr_j, const int* indptr, const int* zslot, const float* Uz,
                          const float* Vz, const float* Wk, const float* means, const float* betas, double alpha, double lo, double hi,
                          int R, int n, int F, int Z, float* oI, float* oA, float* oS, void* stream);
int tmd_embed_moments_bwd(const float* vec, const float* dist, const int* nbr_j, const int* indptr, const int* zslot, const float* Uz,
                          const float* Vz, const float* Wk, const float* Wh, const float* means, const float* betas, double alpha,
                          double lo, double hi, int R, int n, int F, int Z, const float* dI, const float* dA, const float* dS,
                          float* g_dist, float* g_vec, void* stream);

/* ---- per-atom dense layers as weight-tiled GEMMs and the row-wise stages around them (csrc/rows_gemm.cu) ---------------------
 * Rows a10 / a11 (TensorNet tail + Scalar head, models/tensornet.py:394-402, models/output_modules.py:79-117) and the scalar gate of
 * the embedding (row a5, models/tensornet.py:588-619) are chains  norm -> LayerNorm -> Linear -> SiLU -> Linear -> SiLU.
 * tmd_rows_gemm: out[M,N] = epilogue(A[M,K] . W[K,N]), fp32 operands, 3xTF32 warp-level tensor-core MMAs with fp32 parity; W is the
 * TRANSPOSED nn.Linear weight in the forward and the nn.Linear weight itself in the backward.  lda/ldw/ldo/ldaux: row strides (floats).
 *   epi 0: z = acc + vec[col]; zout (may be NULL) = z; out = SiLU(z)        (forward layer, saves the pre-activation)
 *   epi 1: out = acc * SiLU'(aux[row,col])                                  (backward through an activation)
 *   epi 2: out = acc * vec[col]                                             (backward into a LayerNorm: x gamma)
 *   epi 3: out = acc (+ vec[col] when vec != NULL)                          (plain Linear / its transposed backward)
 *   epi 4: out = acc + aux[row,col]                                         (gradient of a tensor consumed by two Linears)
 * K % 16 == 0, N % 32 == 0.
 * tmd_norm_ln_fwd: xn = LayerNorm(norm(X)), mode 3: norm = (3 I^2, |A|^2, |S|^2) in R^{3F} (tensor_norm3), mode 2: |X|^2 per channel
 * in R^F; saves (mu, rstd) [N,2].  tmd_norm_ln_bwd: g [N,K] = d xn * gamma -> gX = add + d/dX (add: all NULL or all given).
 * tmd_head_energy: e = (w3 . h2 + b3) * std + atomref[z];  tmd_head_seed: dz2 = gy[batch] * std * w3 * SiLU'(z2).
 * tmd_gate_bwd_preact (tensor_ops.cu): gate backward with the gate given as its pre-activation z: g = SiLU(z); returns the gradient of
 * the gated tensors (gI, gA, gS) and dg [N,3F] = gradient of the gates (chain == 0) or of z (chain != 0). */
int tmd_rows_gemm(const float* A, int lda, const float* W, int ldw, int M, int K, int N, int epi, const float* vec, const float* aux,
                  int ldaux, float* zout, float* out, int ldo, void* stream);
int tmd_norm_ln_fwd(const float* I, const float* A, const float* S, int N, int F, int mode, const float* ln_w, const float* ln_b, double eps,
                    float* xn, float* mu_rstd, void* stream);
int tmd_norm_ln_bwd(const float* I, const float* A, const float* S, int N, int F, int mode, const float* g, const float* mu_rstd,
                    const float* addI, const float* addA, const float* addS, float* gI, float* gA, float* gS, void* stream);
int tmd_head_energy(const float* h2, int N, int H, const float* w3, const float* b3, const float* std_, const int64_t* z, const float* atomref,
                    int n_ref, float* e, void* stream);
int tmd_head_seed(const float* gy, const int64_t* batch, int M, const float* std_, const float* w3, const float* z2, int N, int H, float* dz2,
                  void* stream);
int tmd_gate_bwd_preact(const float* I, const float* A, const float* S, const float* z, const float* dI, const float* dA, const float* dS,
                        float* gI, float* gA, float* gS, float* dg, int chain, int N, int F, void* stream);
/* ---- a15: EquivariantScalar head of the ET = two GatedEquivariantBlocks (reference models/utils.py:583-655, output_modules.py:159-163).
 * The projections and the update MLP run on tmd_rows_gemm (epilogue 3 = plain Linear, 4 = acc + aux for a tensor consumed by two
 * Linears) or, for layers with <= 8 outputs (the last block's Linear(F/2 -> 2) and vec2_proj(F/2 -> 1)), on tmd_rows_narrow_fwd/bwd;
 * around them: tmd_geb_cat_fwd: cat [N,2H] = [x | |P1|] with |P1| the norm over the 3 vector components of P1 = vec1_proj(v)
 * (0 for an all-zero row, like the reference's masked torch.norm); tmd_geb_cat_bwd: g_P1 = g_vec1 * P1 / |P1|;
 * tmd_geb_out_fwd: hh [N,2O] = [xo | gate] -> xo_out = act ? SiLU(xo) : xo, vo = gate * V2; tmd_geb_out_bwd: its backward
 * (g_xo with row stride ldgx, g_xo / g_vo may be NULL = zero). */
int tmd_geb_cat_fwd(const float* x, const float* P1, int N, int H, float* cat, void* stream);
int tmd_geb_cat_bwd(const float* P1, const float* cat, const float* g_cat, int N, int H, float* g_P1, void* stream);
int tmd_geb_out_fwd(const float* hh, const float* V2, int N, int O, int act, float* xo, float* vo, void* stream);
int tmd_geb_out_bwd(const float* hh, const float* V2, const float* g_xo, int ldgx, const float* g_vo, int N, int O, int act, float* g_hh,
                    float* g_V2, void* stream);
int tmd_rows_narrow_fwd(const float* A, int lda, const float* W, int M, int K, int Nn, const float* b, float* out, void* stream);
int tmd_rows_narrow_bwd(const float* dOut, const float* Wn, int M, int K, int Nn, int accumulate, const float* aux, float* dA, void* stream);
/* P1 backward with the incoming gradient given as two tensors (d1 + d2), and P3 backward for the dX branch only (the Xn branch of
 * X' = Xn + dX + qf dX.dX is the identity): what the hand-sequenced backward of one interaction layer needs
 * (extensions/fused.py::_InteractionLayer; models/tensornet.py:745, :812). */
int tmd_p1_bwd2(const float* I, const float* A, const float* S, const float* d1I, const float* d1A, const float* d1S, const float* d2I,
                const float* d2A, const float* d2S, float* gI, float* gA, float* gS, int N, int F, void* stream);
int tmd_p3_bwd_dx(const float* DI, const float* DA, const float* DS, const float* qf, const float* dI, const float* dA, const float* dS,
                  float* gDI, float* gDA, float* gDS, int N, int F, void* stream);

/* ---- a10 + a11: TensorNet tail + Scalar head + per-molecule reduction (csrc/head.cu) -------------------------------------
 * Replaces tensor_norm3 -> LayerNorm(3F) -> Linear(3F->F) -> SiLU (models/tensornet.py:394-402), Scalar.pre_reduce = MLP
 * Linear(F->H) -> SiLU -> Linear(H->1) (models/output_modules.py:79-117, models/utils.py:531-580), * std, + atomref[z]
 * (priors/atomref.py:66), OutputModel.reduce = scatter-sum over batch (output_modules.py:43-73) and + mean (model.py:606).
 * tmd_head_fwd: e[N] per-atom energies; saves (mu, rstd)[N,2] of the LayerNorm and the pre-activations z1[N,F], z2[N,H].
 * W1T [3F,F], W2T [F,H] are the transposed nn.Linear weights; atomref [n_ref] may be NULL.  F, H multiples of 32, <= 256.
 * tmd_segment_sum: y[m] = mean + sum of e over the atoms of molecule m (deterministic; `flag` is one int of scratch).
 * tmd_head_bwd: dI, dA, dS for the upstream gradient gy[M] (forces: ones); W1 [F,3F], W2 [H,F] untransposed. */
int tmd_head_fwd(const float* I, const float* A, const float* S, int N, int F, int H, const float* ln_w, const float* ln_b,
                 double eps, const float* W1T, const float* b1, const float* W2T, const float* b2, const float* w3,
                 const float* b3, const float* std_, const int64_t* z, const float* atomref, int n_ref, float* e,
                 float* mu_rstd, float* z1, float* z2, void* stream);
int tmd_segment_sum(const float* e, const int64_t* batch, int N, int M, const float* mean, int* flag, float* y, void* stream);
int tmd_head_bwd(const float* I, const float* A, const float* S, int N, int F, int H, const float* ln_w, double eps,
                 const float* W1, const float* W2, const float* w3, const float* std_, const float* gy, const int64_t* batch,
                 int M, const float* mu_rstd, const float* z1, const float* z2, float* dI, float* dA, float* dS, void* stream);

/* ---- a14: Equivariant-Transformer attention message --------------------------------------------------------------------
 * Replaces EquivariantMultiHeadAttention.propagate/message/aggregate (models/torchmd_et.py:355-426: 4 index_select gathers,
 * ~10 elementwise launches on [E,H,3hd] tensors, 2 scatter-adds).  One warp per target atom of the sorted symmetric CSR;
 * q, kk [N,F], v [N,3F] (per head: a|b|c), vec [N,3,F]; dkv [P,4F] = SiLU(dk_proj(rbf)) [F] | SiLU(dv_proj(rbf)) [3F] per
 * unordered pair (from tmd_pair_chain_tc), dkvp its d/dr.  Outputs xagg [N,F], vagg [N,3,F]; backward returns the node
 * gradients and dE/d dist[E], dE/d rhat[E,3].  F <= 128, head_dim % 4 == 0, head_dim/4 a power of two. */
int tmd_et_msg_fwd(const float* q, const float* kk, const float* v, const float* vec, const float* dkv, const int* pid,
                   const int* nbr_j, const int* indptr, const float* dist, const float* rhat, double lo, double hi,
                   int vector_cutoff, int n, int F, int head_dim, float* xagg, float* vagg, void* stream);
int tmd_et_msg_bwd(const float* q, const float* kk, const float* v, const float* vec, const float* dkv,
                   const float* dkvp, const float* gx, const float* gv, const int* pid, const int* nbr_j,
                   const int* indptr, const float* dist, const float* rhat, double lo, double hi, int vector_cutoff, int n,
                   int F, int head_dim, float* gq, float* gk, float* gv_out, float* gvec, float* g_dist, float* g_rhat,
                   void* stream);

/* ---- a13: ET NeighborEmbedding (models/utils.py:45-117): out[i] = sum over the neighbours j != i of C(r)(W_d rbf(r) + b) * T[z_j].
 * wp / wpp [P_u,F]: the radial factor per unordered pair and its d/dr (tmd_pair_chain_tc / tmd_radial_table); table [max_z,F] the
 * neighbour embedding; z int64 [n].  Backward w.r.t. geometry: g_dist[E] (self-loop slots get 0). */
int tmd_nbr_embed_fwd(const float* wp, const float* table, const int64_t* z, const int* pid, const int* nbr_j, const int* indptr,
                      int n, int F, float* out, void* stream);
int tmd_nbr_embed_bwd(const float* wpp, const float* table, const int64_t* z, const int* pid, const int* nbr_j, const int* indptr,
                      int n, int F, const float* gout, float* g_dist, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TMDNET_B200_H */
