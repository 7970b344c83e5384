"""Fused inference+force path of the TensorNet edge pipeline (C ABI: tmd_pair_*, tmd_mp_pair_bwd, tmd_embed_*).

Used by `models/tensornet.py` when the radial weights need no gradient (evaluation / MD): the per-edge
rbf -> MLP -> cutoff chain runs once per unordered pair with its d/dr tangent, and the backward of the
message passing is a single gather kernel that emits dE/dr per pair.  Training (weight gradients) keeps
using the autograd-composed ops of tensornet_ops.py.
"""
import ctypes
import math
import os
import warnings
from typing import NamedTuple, Optional, Sequence

import torch
from torch import Tensor
from torch.autograd.function import once_differentiable

from .. import _lib
from .neighbors import NeighborGraph

p = _lib.ptr

# tcgen05 (3xTF32) radial chain; False selects the fp32 CUDA-core chain kernel (kept as the arbiter in tests)
USE_TENSOR_CORES = True
# fp16 (hi, lo) split terms instead of tf32 ones when a host-side bound proves every hidden activation / tangent fits fp16
USE_F16_SPLIT = True
# read the radial functions of a frozen model from a validated per-model table instead of re-evaluating the
# chain for every pair of every frame (TMD_RADIAL_TABLE=0 or this flag off -> always the exact tcgen05 / FFMA chain)
USE_RADIAL_TABLE = os.environ.get("TMD_RADIAL_TABLE", "1") != "0"
TABLE_TOL = 2e-6          # interpolation vs the fp64 evaluation, relative to max|f| resp. max|f'|
TABLE_TOL_EXACT = 1e-5    # interpolation vs the exact GPU chain (SURVEY's gate)


class PairIndex(NamedTuple):
    pid: Tensor         # [E_cap] int32  directed edge slot -> pair id
    pair_edge: Tensor   # [P_cap] int32  pair id -> its canonical edge slot (j <= i)
    pair_r: Tensor      # [P_cap] fp32   |r| of the pair (no autograd)
    num_unique: Tensor  # [1] int32
    pair_ptr: Tensor    # [N+1] int32


def build_pair_index(g: NeighborGraph) -> PairIndex:
    n = g.indptr.shape[0] - 1
    cap = g.dist.shape[0]
    dev = g.dist.device
    i32 = dict(device=dev, dtype=torch.int32)
    pid = torch.zeros(cap, **i32)
    pair_edge = torch.zeros(cap, **i32)
    pair_r = torch.zeros(cap, device=dev, dtype=torch.float32)
    num_unique = torch.zeros(1, **i32)
    pair_ptr = torch.empty(n + 1, **i32)
    ws = torch.empty(max(n, 1), **i32)
    _lib.call("tmd_pair_index", p(g.indptr), p(g.nbr_i), p(g.nbr_j), p(g.rev), p(g.num_pairs), cap, n, p(g.dist.detach()),
              p(pair_ptr), p(pid), p(pair_edge), p(pair_r), p(num_unique), p(ws), device=dev)
    return PairIndex(pid, pair_edge, pair_r, num_unique, pair_ptr)


class _PairDist(torch.autograd.Function):
    """pair_r[p] = dist[pair_edge[p]]; backward splits dE/dr_p evenly over the two directed edges (the position
    gradient kernel turns g_e - g_rev(e) back into the full pair force)."""

    @staticmethod
    def forward(ctx, dist, pair_edge, pid, P):
        ctx.save_for_backward(pid)
        ctx.E = dist.shape[0]
        return dist.index_select(0, pair_edge[:P].long())

    @staticmethod
    @once_differentiable
    def backward(ctx, g_pair):
        (pid,) = ctx.saved_tensors
        return 0.5 * g_pair.index_select(0, pid[:ctx.E].long().clamp_(max=g_pair.shape[0] - 1)), None, None, None


def pair_distances(dist: Tensor, pidx: PairIndex, P: int) -> Tensor:
    return _PairDist.apply(dist, pidx.pair_edge, pidx.pid, P)


class RadialChain:
    """Host-side cache of the transposed weights of a Linear stack (rebuilt when a weight is modified in place).
    Each layer is a list of nn.Linear sharing the same input, concatenated along the output axis (the embedding's
    three radial projections form one [R,3F] layer)."""

    def __init__(self, layers: Sequence[Sequence[torch.nn.Linear]], act: bool, mul_cutoff: bool):
        self.layers, self.act, self.mul_cutoff = [list(l) for l in layers], act, mul_cutoff
        self._key, self._wt, self._b, self._dims = None, None, None, None

    def supported_ffma(self) -> bool:
        """fp32 CUDA-core chain kernel: widths % 32 == 0 and <= 384."""
        widths = [sum(m.out_features for m in l) for l in self.layers]
        R = self.layers[0][0].in_features
        return R % 4 == 0 and len(widths) <= 3 and all(w % 32 == 0 and w <= 384 for w in widths)

    def supported(self) -> bool:
        return (USE_TENSOR_CORES and self.supported_tc()) or self.supported_ffma()

    def supported_tc(self) -> bool:
        """tcgen05 path: hidden widths % 32 == 0, every output width % 64 == 0."""
        widths = [sum(m.out_features for m in l) for l in self.layers]
        return len(widths) <= 3 and all(w % 64 == 0 and w <= 1024 for w in widths) and all(w % 32 == 0 for w in widths[:-1])

    def tc_tensors(self, f16: bool = False):
        """hi/lo split of every weight -- TF32 (round-to-nearest-away, = cvt.rna.tf32.f32) or fp16 -- with K of layer 0
        zero-padded to a multiple of 32 (64 for fp16), plus the pointer tables the C ABI takes."""
        self.tensors()
        if f16:
            if getattr(self, "_tc16_key", None) != self._key:
                his, los = [], []
                for wt in self._wt:
                    w = wt.t().contiguous()
                    kpad = (-w.shape[1]) % 64
                    if kpad:
                        w = torch.nn.functional.pad(w, (0, kpad))
                    hi = w.half()
                    his.append(hi.contiguous()); los.append((w - hi.float()).half().contiguous())
                n = len(his)
                mk = lambda ts: (ctypes.c_void_p * 3)(*([t.data_ptr() for t in ts] + [None] * (3 - n)))
                self._tc16 = (his, los, mk(his), mk(los), mk(self._b))
                self._tc16_key = self._key
            return self._tc16
        if getattr(self, "_tc_key", None) != self._key:
            def rna(x):
                b = x.contiguous().view(torch.int32)
                return ((b + 0x1000) & -8192).view(torch.float32)
            his, los = [], []
            for wt in self._wt:               # wt is [K, N]; the tensor core wants W = [N, K], K contiguous
                w = wt.t().contiguous()
                kpad = (-w.shape[1]) % 32
                if kpad:
                    w = torch.nn.functional.pad(w, (0, kpad))
                hi = rna(w)
                his.append(hi.contiguous()); los.append(rna(w - hi).contiguous())
            n = len(his)
            mk = lambda ts: (ctypes.c_void_p * 3)(*([t.data_ptr() for t in ts] + [None] * (3 - n)))
            self._tc = (his, los, mk(his), mk(los), mk(self._b))
            self._tc_key = self._key
        return self._tc

    # ---- tabulated radial function ------------------------------------------------------------------
    def _eval_fp64(self, rbf_mod, lo: float, hi: float, r: Tensor):
        """f(r), f'(r) of this chain in fp64 (torch, forward-mode): table nodes and the reference of the gate."""
        means, betas = rbf_mod.means.detach().double(), rbf_mod.betas.detach().double()
        alpha = float(rbf_mod.alpha)

        def cos_cut(d, a, b):
            if a > 0:
                return 0.5 * (torch.cos(math.pi * (2 * (d - a) / (b - a) + 1.0)) + 1.0) * (d < b) * (d > a)
            return 0.5 * (torch.cos(d * math.pi / b) + 1.0) * (d < b)

        def fn(rr):
            d = rr.unsqueeze(-1)
            x = cos_cut(d, 0.0, hi) * torch.exp(-betas * (torch.exp(alpha * (lo - d)) - means) ** 2)
            for layer in self.layers:
                W = torch.cat([m.weight.detach() for m in layer], 0).double()
                b = torch.cat([m.bias.detach() if m.bias is not None else m.weight.new_zeros(m.out_features) for m in layer], 0).double()
                x = torch.nn.functional.linear(x, W, b)
                if self.act:
                    x = torch.nn.functional.silu(x)
            if self.mul_cutoff:
                x = x * cos_cut(d, lo, hi)
            return x

        return torch.func.jvp(fn, (r,), (torch.ones_like(r),))

    def table(self, rbf_mod, lo: float, hi: float):
        """-> (table [n_int,4,N] fp32, n_int, h) or None.  Built once per weight version: nodes at i*2^-k evaluated in fp64,
        accepted only if the CUDA interpolation reproduces the fp64 function (TABLE_TOL) and the exact GPU chain
        (TABLE_TOL_EXACT) at 3 interior points of every interval; refined (k -> k+1) until it does, else None."""
        from ..models.utils import ExpNormalSmearing
        if not isinstance(rbf_mod, ExpNormalSmearing):
            return None
        self.tensors()
        key = (self._key, float(lo), float(hi), rbf_mod.means.data_ptr(), rbf_mod.means._version, rbf_mod.betas._version)
        if getattr(self, "_tab_key", None) == key:
            return self._tab
        if torch.cuda.is_current_stream_capturing():
            return None                       # cannot validate (host sync) inside a capture; the warm-up pass builds it
        dev = self._wt[0].device
        N = self._wt[-1].shape[1]
        result = None
        with torch.no_grad():
            for k in (9, 10, 11, 12):
                h = 2.0 ** -k
                n_int = int(math.ceil(hi / h))
                nodes = torch.arange(n_int + 1, device=dev, dtype=torch.float64) * h
                f, fp = self._eval_fp64(rbf_mod, lo, hi, nodes)
                m = fp * h
                tab = torch.stack([f[:-1], f[1:] - f[:-1], m[:-1], m[1:]], 1).float().contiguous()
                # gate: three interior points per interval, through the CUDA kernel
                rt = ((torch.arange(n_int, device=dev, dtype=torch.float64).unsqueeze(1) +
                       torch.tensor([0.25, 0.5, 0.75], device=dev, dtype=torch.float64)) * h).reshape(-1)
                rt = rt[rt < hi]
                f64, fp64 = self._eval_fp64(rbf_mod, lo, hi, rt)
                r32 = rt.float()
                nu = torch.tensor([r32.numel()], device=dev, dtype=torch.int32)
                fi = torch.empty(r32.numel(), N, device=dev, dtype=torch.float32)
                fpi = torch.empty_like(fi)
                _lib.call("tmd_radial_table", p(r32), p(nu), r32.numel(), p(tab), n_int, h, N, p(fi), p(fpi), device=dev)
                sf, sfp = f64.abs().max().clamp_min(1e-30), fp64.abs().max().clamp_min(1e-30)
                e_f = ((fi.double() - f64).abs().max() / sf).item()
                e_fp = ((fpi.double() - fp64).abs().max() / sfp).item()
                if e_f < TABLE_TOL and e_fp < TABLE_TOL:
                    fe, fpe = run_chain(self, r32, nu, rbf_mod, lo, hi, True, exact=True)
                    x_f = ((fi - fe).abs().max() / sf).item()
                    x_fp = ((fpi - fpe).abs().max() / sfp).item()
                    if x_f < TABLE_TOL_EXACT and x_fp < TABLE_TOL_EXACT:
                        result = (tab, n_int, h)
                        self.table_report = {"k": k, "intervals": n_int, "bytes": tab.numel() * 4, "err_f_vs_fp64": e_f,
                                             "err_fp_vs_fp64": e_fp, "err_f_vs_exact": x_f, "err_fp_vs_exact": x_fp}
                    else:
                        warnings.warn(f"radial table disagrees with the exact chain ({x_f:.1e}, {x_fp:.1e}); using the exact path")
                    break
        self._tab_key, self._tab = key, result
        return result

    def f16_safe(self, rbf_mod, lo: float, hi: float) -> bool:
        """Conservative bound on |hidden activations| and |d/dr| of the chain (infinity-norm of each weight, |SiLU'| <= 1.1):
        fp16 operands are used only if everything stays below 3e4 and hidden widths are multiples of 64."""
        if not USE_F16_SPLIT:
            return False
        self.tensors()
        key = (self._key, float(lo), float(hi), rbf_mod.means.data_ptr(), rbf_mod.means._version, rbf_mod.betas._version)
        if getattr(self, "_f16_key", None) != key:
            widths = [w.shape[1] for w in self._wt]
            ok = all(w % 64 == 0 for w in widths[:-1])
            if ok:
                with torch.no_grad():
                    r = torch.linspace(0.0, float(hi), 4097, device=self._wt[0].device, dtype=torch.float64)
                    alpha = float(rbf_mod.alpha)
                    c0 = 0.5 * (torch.cos(r * 3.141592653589793 / float(hi)) + 1.0)
                    xd = c0[:, None] * torch.exp(-rbf_mod.betas.double() * (torch.exp(alpha * (float(lo) - r))[:, None] - rbf_mod.means.double()) ** 2)
                    # |d rbf_k / dr| bound: finite differences on the fine grid are enough for a 2x-margin check
                    t = float(((xd[1:] - xd[:-1]).abs().max() / (float(hi) / 4096)).item()) * 2.0
                    v = 1.0
                    ok = ok and t < 3e4
                    for wt, b in zip(self._wt[:-1], self._b[:-1]):
                        nrm = float(wt.abs().sum(0).max().item())        # wt is [K, N]: max over outputs of sum_k |W|
                        v = nrm * v + float(b.abs().max().item())
                        t = 1.1 * nrm * t
                        ok = ok and v < 3e4 and t < 3e4
            self._f16_ok, self._f16_key = bool(ok), key
        return self._f16_ok

    def frozen(self) -> bool:
        return not any(t.requires_grad for l in self.layers for m in l for t in (m.weight, m.bias))

    def tensors(self):
        key = tuple((m.weight.data_ptr(), m.weight._version, m.bias.data_ptr(), m.bias._version) for l in self.layers for m in l)
        if key != self._key:
            self._wt = [torch.cat([m.weight.detach() for m in l], 0).t().contiguous() for l in self.layers]
            self._b = [torch.cat([m.bias.detach() for m in l], 0).contiguous() for l in self.layers]
            dims = [self.layers[0][0].in_features] + [w.shape[1] for w in self._wt]
            self._dims = (ctypes.c_int * 4)(*(dims + [0] * 3)[:4])
            self._key = key
        return self._wt, self._b, self._dims


class RbfOperand(NamedTuple):
    hi: Tensor
    lo: Tensor
    rows_per_pair: int
    f16: bool


def precompute_rbf(pair_r: Tensor, num_unique: Tensor, rbf_mod, lo: float, hi: float, want_tangent: bool,
                   f16: bool = False) -> Optional[RbfOperand]:
    """Layer-0 tensor-core operand (rbf and d rbf/dr, split into tf32 or fp16 terms), shared by all radial chains of a frame."""
    P = pair_r.shape[0]
    if not USE_TENSOR_CORES or P == 0:
        return None
    rpp = 2 if want_tangent else 1
    R = rbf_mod.means.shape[0]
    rows_pad = -(-(P * rpp) // 128) * 128
    K0 = -(-R // 32) * 32
    buf = torch.empty(2, rows_pad, K0, device=pair_r.device, dtype=torch.float32)
    _lib.call("tmd_pair_rbf_split", p(pair_r), p(num_unique), P, R, rpp, p(rbf_mod.means.detach()), p(rbf_mod.betas.detach()),
              float(rbf_mod.alpha), float(lo), float(hi), p(buf[0]), p(buf[1]), int(f16), device=pair_r.device)
    return RbfOperand(buf[0], buf[1], rpp, bool(f16))


def shared_rbf_operand(chains, pair_r: Tensor, num_unique: Tensor, rbf_mod, lo: float, hi: float, want_tangent: bool):
    """One layer-0 operand for all radial chains of a frame when they agree on the operand type (fp16 / tf32); none when every
    chain is read from its table."""
    if not USE_TENSOR_CORES or not all(c.supported_tc() for c in chains):
        return None
    if USE_RADIAL_TABLE and all(c.table(rbf_mod, lo, hi) is not None for c in chains):
        return None
    modes = {c.f16_safe(rbf_mod, lo, hi) for c in chains}
    if len(modes) != 1:
        return None
    return precompute_rbf(pair_r, num_unique, rbf_mod, lo, hi, want_tangent, f16=modes.pop())


def _chain_launches(dims, n_layers, f16, tangent, act, have_rbf):
    """kernels one tmd_pair_chain_tc call launches: rbf operand (unless shared) + one per layer, layers 2+3 fused when the
    tail kernel applies (pair_chain_tc.cu: tail_ok)."""
    n = n_layers + (0 if have_rbf else 1)
    fused_tail = (os.environ.get("TMD_TC_FUSE", "1") != "0" and f16 and n_layers == 3 and tangent and act and
                  dims[1] in (64, 128) and dims[2] == 2 * dims[1] and dims[3] == 3 * dims[1])
    return n - 1 if fused_tail else n


def run_chain(chain: RadialChain, pair_r: Tensor, num_unique: Tensor, rbf_mod, lo: float, hi: float, want_tangent: bool,
              rbf_op: Optional[RbfOperand] = None, exact: bool = False):
    """-> (f [P,N_last], f' or None).  `rbf_mod` is the ExpNormalSmearing module (means/betas/alpha).  Unless `exact`, a frozen
    chain is read from its validated table; the exact tcgen05 / FFMA chain is the fallback and the gate."""
    wt, b, dims = chain.tensors()
    P = pair_r.shape[0]
    n_last = wt[-1].shape[1]
    f = torch.empty(P, n_last, device=pair_r.device, dtype=torch.float32)
    fp = torch.empty_like(f) if want_tangent else None
    if USE_RADIAL_TABLE and not exact and P > 0 and n_last % 4 == 0:
        tab = chain.table(rbf_mod, lo, hi)
        if tab is not None:
            _lib.call("tmd_radial_table", p(pair_r), p(num_unique), P, p(tab[0]), tab[1], tab[2], n_last, p(f), p(fp),
                      device=pair_r.device)
            return f, fp
    if USE_TENSOR_CORES and chain.supported_tc() and P > 0:
        f16 = chain.f16_safe(rbf_mod, lo, hi)
        his, los, p_hi, p_lo, p_b = chain.tc_tensors(f16)
        rows_pad = -(-(P * (2 if want_tangent else 1)) // 128) * 128
        width = max([his[0].shape[1]] + [w.shape[1] for w in wt[:-1]])
        ws = torch.empty(4, rows_pad, width, device=pair_r.device, dtype=torch.float32)
        if rbf_op is not None and (rbf_op.rows_per_pair != (2 if want_tangent else 1) or rbf_op.f16 != f16):
            rbf_op = None
        _lib.call("tmd_pair_chain_tc", p(pair_r), p(num_unique), P, ctypes.cast(p_hi, ctypes.c_void_p),
                  ctypes.cast(p_lo, ctypes.c_void_p), ctypes.cast(p_b, ctypes.c_void_p), ctypes.cast(dims, ctypes.c_void_p),
                  len(wt), int(chain.act), int(chain.mul_cutoff), p(rbf_mod.means.detach()), p(rbf_mod.betas.detach()),
                  float(rbf_mod.alpha), float(lo), float(hi), p(rbf_op.hi) if rbf_op else None, p(rbf_op.lo) if rbf_op else None,
                  p(ws[0]), p(ws[1]), p(ws[2]), p(ws[3]), p(f), p(fp), int(f16), device=pair_r.device,
                  launches=_chain_launches(list(dims), len(wt), f16, want_tangent, chain.act, rbf_op is not None))
        return f, fp
    if not chain.supported_ffma():
        raise RuntimeError("radial chain widths are supported by neither the tcgen05 nor the fp32 chain kernel")
    wt3 = wt + [None] * (3 - len(wt))
    b3 = b + [None] * (3 - len(b))
    _lib.call("tmd_pair_chain", p(pair_r), p(num_unique), P, p(wt3[0]), p(b3[0]), p(wt3[1]), p(b3[1]), p(wt3[2]), p(b3[2]),
              ctypes.cast(dims, ctypes.c_void_p), len(wt), int(chain.act), int(chain.mul_cutoff), p(rbf_mod.means.detach()),
              p(rbf_mod.betas.detach()), float(rbf_mod.alpha), float(lo), float(hi), p(f), p(fp), device=pair_r.device)
    return f, fp


def _triple(N, F, dev):
    return (torch.empty(N, 1, F, device=dev), torch.empty(N, 3, F, device=dev), torch.empty(N, 5, F, device=dev))


# cooperative (group-staged) message passing kernels, csrc/mp_coop.cu; TMD_MP_COOP=0 selects the warp-per-atom kernels
USE_MP_COOP = os.environ.get("TMD_MP_COOP", "1") != "0"


class MessagePassingPlan:
    """The interaction message passing of one frame in the form the model calls it (pair-indexed radial weights f[P_u,3F] read
    through `pid`): `forward(Y, f) -> M`, `backward(Y, dM, f, fp) -> (dY, g_dist)`.  bench.py times these two entry points for
    the `roofline` / `roofline_secondary` keys, so the benchmark and the model share one launch path.

    Cooperative path (F <= 128): one `tmd_mp_plan` launch per frame groups the rows (12-15 per group, the size that fills the 148 SMs best) in the neighbour search's cell order
    (spatially adjacent atoms share most neighbours) and lists each group's distinct neighbours; the forward / backward kernels
    stage every distinct neighbour row once per group in shared memory.  Results are bit-identical to the warp-per-atom kernels."""

    def __init__(self, g: NeighborGraph, pidx: PairIndex, F: Optional[int] = None):
        self.g, self.pidx = g, pidx
        self.N = g.indptr.shape[0] - 1
        self.coop = bool(USE_MP_COOP and self.N > 0 and (F is None or (F % 4 == 0 and F <= 128)))
        if self.coop:
            dev, cap, n = g.dist.device, g.dist.shape[0], self.N
            forced = int(os.environ.get("TMD_MP_GROUP", "0"))       # rows per group; default: the size that fills the SM slots best
            self.group = forced if forced > 0 else int(_lib.library().tmd_mp_pick_group(n))
            ng = -(-n // self.group)
            i32 = dict(device=dev, dtype=torch.int32)
            self.rows = torch.empty(n, **i32)
            self.union_j = torch.empty(cap, **i32)
            self.ucount = torch.empty(cap, device=dev, dtype=torch.uint8)
            self.upos = torch.empty(cap, device=dev, dtype=torch.int16)
            self.goff, self.gcnt = torch.empty(ng, **i32), torch.empty(ng, **i32)
            self.cursor = torch.empty(1, **i32)
            has_order = g.order is not None and g.order.numel() == n
            _lib.call("tmd_mp_plan", p(g.indptr), p(g.nbr_j), p(g.num_pairs), cap, n, self.group, p(g.order) if has_order else None,
                      p(g.order_flag) if has_order else None, p(self.rows), p(self.cursor), p(self.union_j), p(self.ucount),
                      p(self.upos), p(self.goff), p(self.gcnt), device=dev)

    @property
    def fwd_kernel_name(self):
        return "mp_coop_fwd_kernel" if self.coop else "mp_fwd_kernel"

    @property
    def bwd_kernel_name(self):
        return "mp_coop_bwd_kernel" if self.coop else "mp_edge_bwd_kernel"

    # dram bytes per launch on the C4 graph from the committed ncu captures (bench.py `roofline.traffic`)
    fwd_traffic_bytes = 0.605e9
    bwd_traffic_bytes = 1.455e9
    traffic_source = "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch on the C4 graph (profiles/r02_mp_coop_notes.md)"

    def _ok(self, F):
        return self.coop and F % 4 == 0 and F <= 128

    def forward(self, Y, f):
        YI, YA, YS = Y
        N, F = self.N, YI.shape[-1]
        M = _triple(N, F, YI.device)
        g, x = self.g, self.pidx
        if self._ok(F):
            _lib.call("tmd_mp_coop_fwd", p(YI), p(YA), p(YS), p(f), p(x.pid), p(g.nbr_j), p(g.indptr), p(g.num_pairs), g.dist.shape[0],
                      N, F, self.group, p(self.rows), p(self.union_j), p(self.ucount), p(self.upos), p(self.goff), p(self.gcnt), p(M[0]), p(M[1]),
                      p(M[2]), device=YI.device)
        else:
            _lib.call("tmd_mp_fwd", p(YI), p(YA), p(YS), p(f), p(x.pid), p(g.nbr_j), p(g.indptr), N, F, p(M[0]), p(M[1]), p(M[2]),
                      device=YI.device)
        return M

    def backward(self, Y, dM, f, fp, add=None):
        """-> (dY, g_dist); `add` (optional triple): a gradient that reached Y along another path, summed into dY by the kernel"""
        YI, YA, YS = Y
        N, F = self.N, YI.shape[-1]
        dY = _triple(N, F, YI.device)
        g, x = self.g, self.pidx
        ad = (None, None, None) if add is None else add
        if self._ok(F):
            # every valid slot is written by the kernel; padded slots (static shapes) are never read by the position gradient
            g_dist = torch.empty(g.dist.shape[0], device=YI.device, dtype=torch.float32) if fp is not None else None
            _lib.call("tmd_mp_coop_bwd", p(YI), p(YA), p(YS), p(dM[0]), p(dM[1]), p(dM[2]), p(f), p(fp), p(x.pid), p(g.nbr_j),
                      p(g.indptr), p(g.num_pairs), g.dist.shape[0], N, F, self.group, p(self.rows), p(self.union_j), p(self.ucount), p(self.upos),
                      p(self.goff), p(self.gcnt), p(ad[0]), p(ad[1]), p(ad[2]), p(dY[0]), p(dY[1]), p(dY[2]), p(g_dist), device=YI.device)
        else:
            g_dist = torch.zeros(g.dist.shape[0], device=YI.device, dtype=torch.float32) if fp is not None else None
            _lib.call("tmd_mp_edge_bwd", p(YI), p(YA), p(YS), p(dM[0]), p(dM[1]), p(dM[2]), p(f), p(fp), p(x.pid), p(g.nbr_j),
                      p(g.indptr), N, F, p(dY[0]), p(dY[1]), p(dY[2]), p(g_dist), device=YI.device)
            if add is not None:
                dY = tuple(a + b for a, b in zip(dY, add))
        return dY, g_dist


class _InteractionMP(torch.autograd.Function):
    """M = sum_e f(r_e) * Y_j with f = cutoff * MLP(rbf(r)) evaluated per pair (rows a6 + a7).  `dist` is the per-slot
    distance (the differentiable input); `pair_r` its per-pair gather (no autograd)."""

    @staticmethod
    def forward(ctx, YI, YA, YS, dist, pair_r, chain, rbf_mod, lo, hi, g, pidx, rbf_op, plan):
        YI, YA, YS = YI.contiguous(), YA.contiguous(), YS.contiguous()
        need = dist.requires_grad
        f, fp = run_chain(chain, pair_r.detach(), pidx.num_unique, rbf_mod, lo, hi, need, rbf_op)
        if plan is None:
            plan = MessagePassingPlan(g, pidx, YI.shape[-1])
        M = plan.forward((YI, YA, YS), f)
        ctx.save_for_backward(YI, YA, YS, f, fp if fp is not None else f)
        ctx.has_fp, ctx.plan, ctx.E = fp is not None, plan, dist.shape[0]
        return M

    @staticmethod
    @once_differentiable
    def backward(ctx, dI, dA, dS):
        YI, YA, YS, f, fp = ctx.saved_tensors
        dM = (dI.contiguous(), dA.contiguous(), dS.contiguous())   # bound to a local: the copies outlive the launch
        dY, g_dist = ctx.plan.backward((YI, YA, YS), dM, f, fp if ctx.has_fp else None)
        if g_dist is not None and g_dist.shape[0] != ctx.E:     # dynamic shapes: `dist` is the exact-size prefix of the capacity buffer
            g_dist = g_dist[: ctx.E]
        return dY[0], dY[1], dY[2], g_dist, None, None, None, None, None, None, None, None, None


def interaction_message_passing(Y, dist, pair_r, chain, rbf_mod, lo, hi, g, pidx, rbf_op=None, plan=None):
    return _InteractionMP.apply(Y[0], Y[1], Y[2], dist, pair_r, chain, rbf_mod, lo, hi, g, pidx, rbf_op, plan)


class _EmbeddingMP(torch.autograd.Function):
    """I/A/S of the TensorNet embedding from the neighbour list's (dist, vec) with w = C(r)(U_i+V_j)(W_k rbf + b_k) built in-kernel
    (row a5).  U, V: per-atom rows [N,F] (zidx None) or per-element tables [max_z,F] read through the atomic numbers zidx [N]."""

    @staticmethod
    def forward(ctx, dist, vec, U, V, zidx, pair_r, chain, rbf_mod, lo, hi, g, pidx, rbf_op):
        need = dist.requires_grad or vec.requires_grad
        dp, dpp = run_chain(chain, pair_r.detach(), pidx.num_unique, rbf_mod, lo, hi, need, rbf_op)
        dist_c, vec_c, U, V = dist.detach().contiguous(), vec.detach().contiguous(), U.contiguous(), V.contiguous()
        zidx = None if zidx is None else zidx.contiguous()
        N, F = g.indptr.shape[0] - 1, U.shape[-1]
        out = _triple(N, F, U.device)
        _lib.call("tmd_embed_fwd", p(vec_c), p(dist_c), p(dp), p(pidx.pid), p(g.nbr_j), p(g.indptr), p(U), p(V), p(zidx), float(lo),
                  float(hi), N, F, p(out[0]), p(out[1]), p(out[2]), device=U.device)
        ctx.save_for_backward(dist_c, vec_c, dp, dpp if dpp is not None else dp, pidx.pid, g.nbr_j, g.indptr, U, V,
                              zidx if zidx is not None else pidx.pid)
        ctx.has_tangent, ctx.has_z, ctx.lo, ctx.hi, ctx.N = dpp is not None, zidx is not None, float(lo), float(hi), N
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dI, dA, dS):
        dist, vec, dp, dpp, pid, nbr_j, indptr, U, V, zidx = ctx.saved_tensors
        if not ctx.has_tangent:
            return (None,) * 13
        N, F = ctx.N, U.shape[-1]
        g_dist, g_vec = torch.zeros_like(dist), torch.zeros_like(vec)
        dI, dA, dS = dI.contiguous(), dA.contiguous(), dS.contiguous()
        _lib.call("tmd_embed_bwd", p(vec), p(dist), p(dp), p(dpp), p(pid), p(nbr_j), p(indptr), p(U), p(V), p(zidx) if ctx.has_z else None,
                  ctx.lo, ctx.hi, N, F, p(dI), p(dA), p(dS), p(g_dist), p(g_vec), device=U.device)
        return (g_dist, g_vec) + (None,) * 11


def embedding_message_passing(dist, vec, U, V, pair_r, chain, rbf_mod, lo, hi, g, pidx, rbf_op=None, zidx=None):
    return _EmbeddingMP.apply(dist, vec, U, V, zidx, pair_r, chain, rbf_mod, lo, hi, g, pidx, rbf_op)


# whole-stage autograd nodes (hand-sequenced backward, fused gradient sums): TMD_FUSED_LAYER=0 / TMD_EMBED_TAIL=0 -> one node per kernel
USE_FUSED_LAYER = os.environ.get("TMD_FUSED_LAYER", "1") != "0"
USE_EMBED_TAIL = os.environ.get("TMD_EMBED_TAIL", "1") != "0"
# per-atom dense layers (head, embedding gate MLP) as weight-tiled GEMM launches (csrc/rows_gemm.cu); TMD_ROWS_GEMM=0 -> the head runs
# the per-32-atom-tile kernels of csrc/head.cu (whole chain in one launch, but every tile re-streams the weights: latency-bound)
USE_ROWS_GEMM = os.environ.get("TMD_ROWS_GEMM", "1") != "0"
# embedding by radial moments (csrc/embed_moments.cu): no per-pair radial rows at all; TMD_EMBED_MOMENTS=0 -> chain + embed kernels
USE_EMBED_MOMENTS = os.environ.get("TMD_EMBED_MOMENTS", "1") != "0"
MAX_SPECIES = 16


class SpeciesSlots:
    """zslot[N] (int32 slot of every atom's element among the Z elements present) for the LAST atomic-number tensor seen.  Finding
    the distinct elements needs one host read (torch.unique), so the result is cached against the identity of `z` (the tensor is
    kept alive, so its storage cannot be recycled under the key); inside a CUDA-graph capture only a cached entry can be used."""

    def __init__(self):
        self._z, self._ver, self._val = None, None, None

    def get(self, z: Tensor):
        c = self._z
        if c is not None and c.data_ptr() == z.data_ptr() and c.shape == z.shape and self._ver == z._version and c.device == z.device:
            return self._val
        if z.is_cuda and torch.cuda.is_current_stream_capturing():
            return None
        species = torch.unique(z)                         # host sync: output size
        if species.numel() == 0 or species.numel() > MAX_SPECIES:
            val = None
        else:
            zslot = torch.searchsorted(species, z).to(torch.int32).contiguous()
            val = (zslot, species)
        self._z, self._ver, self._val = z, z._version, val
        return val


class EmbedMomentWeights:
    """Radial projections of the embedding as Wk [3, R+1, F] (k-major, row R = bias) and Wh [3, F/4, R, 4]; Uz / Vz rows of the element
    tables for the species present.  Cached per weight version / species set."""

    def __init__(self, projs: Sequence[torch.nn.Linear]):
        self.projs = list(projs)
        self._key, self._uvkey = None, None

    def frozen(self) -> bool:
        return not any(t.requires_grad for m in self.projs for t in (m.weight, m.bias))

    def supported(self, rbf_mod) -> bool:
        from ..models.utils import ExpNormalSmearing
        R, F = self.projs[0].in_features, self.projs[0].out_features
        return (isinstance(rbf_mod, ExpNormalSmearing) and len(self.projs) == 3 and R <= 64 and F % 4 == 0 and
                all(m.in_features == R and m.out_features == F and m.bias is not None and m.weight.dtype == torch.float32 for m in self.projs))

    def tensors(self):
        key = tuple((t.data_ptr(), t._version) for m in self.projs for t in (m.weight, m.bias))
        if key != self._key:
            W = torch.stack([m.weight.detach() for m in self.projs], 0)                    # [3, F, R]
            b = torch.stack([m.bias.detach() for m in self.projs], 0)                      # [3, F]
            self._Wk = torch.cat([W.transpose(1, 2), b.unsqueeze(1)], 1).contiguous()      # [3, R+1, F]
            F_, R_ = W.shape[1], W.shape[2]
            self._Wh = W.view(3, F_ // 4, 4, R_).permute(0, 1, 3, 2).contiguous()             # [3, F/4, R, 4]: one float4 per (4 channels, k)
            self._key = key
        return self._Wk, self._Wh

    def species_rows(self, U: Tensor, V: Tensor, species: Tensor):
        key = (U.data_ptr(), V.data_ptr(), species.data_ptr())
        if key != self._uvkey:
            self._uv = (U.index_select(0, species).contiguous(), V.index_select(0, species).contiguous(), U, V, species)
            self._uvkey = key
        return self._uv[0], self._uv[1]


class _EmbeddingMoments(torch.autograd.Function):
    """I/A/S of the TensorNet embedding (row a5) from per-atom radial moments: see csrc/embed_moments.cu."""

    @staticmethod
    def forward(ctx, dist, vec, zslot, Uz, Vz, Wk, Wh, rbf_mod, lo, hi, g):
        dist_c, vec_c = dist.detach().contiguous(), vec.detach().contiguous()
        means, betas = rbf_mod.means.detach().contiguous(), rbf_mod.betas.detach().contiguous()
        N, F, R, Z = g.indptr.shape[0] - 1, Uz.shape[-1], Wh.shape[2], Uz.shape[0]
        out = _triple(N, F, Uz.device)
        _lib.call("tmd_embed_moments_fwd", p(vec_c), p(dist_c), p(g.nbr_j), p(g.indptr), p(zslot), p(Uz), p(Vz), p(Wk), p(means), p(betas),
                  float(rbf_mod.alpha), float(lo), float(hi), R, N, F, Z, p(out[0]), p(out[1]), p(out[2]), device=Uz.device)
        ctx.save_for_backward(dist_c, vec_c, zslot, Uz, Vz, Wk, Wh, means, betas, g.nbr_j, g.indptr)
        ctx.cfg = (float(rbf_mod.alpha), float(lo), float(hi), N, dist.requires_grad or vec.requires_grad)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dI, dA, dS):
        dist, vec, zslot, Uz, Vz, Wk, Wh, means, betas, nbr_j, indptr = ctx.saved_tensors
        alpha, lo, hi, N, need = ctx.cfg
        if not need:
            return (None,) * 11
        F, R, Z = Uz.shape[-1], Wh.shape[2], Uz.shape[0]
        g_dist, g_vec = torch.zeros_like(dist), torch.zeros_like(vec)
        dI, dA, dS = dI.contiguous(), dA.contiguous(), dS.contiguous()
        _lib.call("tmd_embed_moments_bwd", p(vec), p(dist), p(nbr_j), p(indptr), p(zslot), p(Uz), p(Vz), p(Wk), p(Wh), p(means), p(betas),
                  alpha, lo, hi, R, N, F, Z, p(dI), p(dA), p(dS), p(g_dist), p(g_vec), device=Uz.device)
        return (g_dist, g_vec) + (None,) * 9


def embedding_moments(dist, vec, zslot, Uz, Vz, Wk, Wh, rbf_mod, lo, hi, g):
    return _EmbeddingMoments.apply(dist, vec, zslot, Uz, Vz, Wk, Wh, rbf_mod, lo, hi, g)


class _ETMessage(torch.autograd.Function):
    """Fused ET attention message + aggregation (row a14): see csrc/et_attention.cu."""

    @staticmethod
    def forward(ctx, q, kk, v, vec, dist, rhat, pair_r, chain, rbf_mod, lo, hi, vcut, head_dim, g, pidx, rbf_op):
        need = dist.requires_grad or rhat.requires_grad
        dkv, dkvp = run_chain(chain, pair_r.detach(), pidx.num_unique, rbf_mod, lo, hi, need, rbf_op)
        q, kk, v, vec = q.contiguous(), kk.contiguous(), v.contiguous(), vec.contiguous()
        dist_c, rhat_c = dist.detach().contiguous(), rhat.detach().contiguous()
        N, F = q.shape[0], q.shape[-1]
        xagg = torch.empty(N, F, device=q.device, dtype=torch.float32)
        vagg = torch.empty(N, 3, F, device=q.device, dtype=torch.float32)
        _lib.call("tmd_et_msg_fwd", p(q), p(kk), p(v), p(vec), p(dkv), p(pidx.pid), p(g.nbr_j), p(g.indptr), p(dist_c), p(rhat_c),
                  float(lo), float(hi), int(vcut), N, F, int(head_dim), p(xagg), p(vagg), device=q.device)
        ctx.save_for_backward(q, kk, v, vec, dkv, dkvp if dkvp is not None else dkv, pidx.pid, g.nbr_j, g.indptr, dist_c, rhat_c)
        ctx.cfg = (float(lo), float(hi), int(vcut), int(head_dim), dkvp is not None)
        return xagg, vagg

    @staticmethod
    @once_differentiable
    def backward(ctx, gx, gv):
        q, kk, v, vec, dkv, dkvp, pid, nbr_j, indptr, dist, rhat = ctx.saved_tensors
        lo, hi, vcut, hd, has_t = ctx.cfg
        if not has_t:
            raise RuntimeError("ET fused message was evaluated without tangents (positions did not require grad)")
        N, F = q.shape[0], q.shape[-1]
        gq, gk = torch.empty_like(q), torch.empty_like(kk)
        gvo, gvec = torch.empty_like(v), torch.empty_like(vec)
        g_dist, g_rhat = torch.zeros_like(dist), torch.zeros_like(rhat)
        gx, gv = gx.contiguous(), gv.contiguous()
        _lib.call("tmd_et_msg_bwd", p(q), p(kk), p(v), p(vec), p(dkv), p(dkvp), p(gx), p(gv), p(pid),
                  p(nbr_j), p(indptr), p(dist), p(rhat), lo, hi, vcut, N, F, hd, p(gq), p(gk), p(gvo), p(gvec), p(g_dist),
                  p(g_rhat), device=q.device)
        return (gq, gk, gvo, gvec, g_dist, g_rhat) + (None,) * 10


def et_message(q, kk, v, vec, dist, rhat, pair_r, chain, rbf_mod, lo, hi, vector_cutoff, head_dim, g, pidx, rbf_op=None):
    """(x_agg [N,F], vec_agg [N,3,F]) of one ET attention layer; q, kk [N,F], v [N,3F], vec [N,3,F]."""
    return _ETMessage.apply(q, kk, v, vec, dist, rhat, pair_r, chain, rbf_mod, lo, hi, vector_cutoff, head_dim, g, pidx, rbf_op)


class _NbrEmbed(torch.autograd.Function):
    """ET NeighborEmbedding sum (row a13): agg[i] = sum_{j != i} C(r)(W_d rbf(r) + b) * T[z_j], radial factor per unordered pair."""

    @staticmethod
    def forward(ctx, dist, pair_r, chain, rbf_mod, lo, hi, table, z, g, pidx, rbf_op):
        need = dist.requires_grad
        wp, wpp = run_chain(chain, pair_r.detach(), pidx.num_unique, rbf_mod, lo, hi, need, rbf_op)
        table, z = table.detach().contiguous(), z.contiguous()
        N, F = g.indptr.shape[0] - 1, table.shape[-1]
        out = torch.empty(N, F, device=table.device, dtype=torch.float32)
        _lib.call("tmd_nbr_embed_fwd", p(wp), p(table), p(z), p(pidx.pid), p(g.nbr_j), p(g.indptr), N, F, p(out), device=table.device)
        ctx.save_for_backward(wpp if wpp is not None else wp, table, z, pidx.pid, g.nbr_j, g.indptr)
        ctx.has_t, ctx.E, ctx.N = wpp is not None, dist.shape[0], N
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        wpp, table, z, pid, nbr_j, indptr = ctx.saved_tensors
        if not ctx.has_t:
            return (None,) * 11
        gout = gout.contiguous()
        g_dist = torch.zeros(ctx.E, device=gout.device, dtype=torch.float32)
        _lib.call("tmd_nbr_embed_bwd", p(wpp), p(table), p(z), p(pid), p(nbr_j), p(indptr), ctx.N, table.shape[-1], p(gout), p(g_dist),
                  device=gout.device)
        return (g_dist,) + (None,) * 10


def neighbor_embedding_sum(dist, pair_r, chain, rbf_mod, lo, hi, table, z, g, pidx, rbf_op=None):
    return _NbrEmbed.apply(dist, pair_r, chain, rbf_mod, lo, hi, table, z, g, pidx, rbf_op)


# ------------------------------------------------------------------------------------------ node linears on tensor cores
def _tf32_split(w: Tensor):
    b = w.contiguous().view(torch.int32)
    hi = ((b + 0x1000) & -8192).view(torch.float32)
    lo = w - hi
    lo = ((lo.contiguous().view(torch.int32) + 0x1000) & -8192).view(torch.float32)
    return hi.contiguous(), lo.contiguous()


class IrrepsLinearWeights:
    """TF32-split, stacked copies of three bias-free nn.Linear(F, F) (forward: W, backward: W^T), cached per weight version."""

    def __init__(self, linears: Sequence[torch.nn.Linear]):
        self.linears = list(linears)
        self._key = None

    def supported(self) -> bool:
        F = self.linears[0].in_features
        return (USE_TENSOR_CORES and F % 64 == 0 and F <= 256 and
                all(l.bias is None and l.in_features == F and l.out_features == F for l in self.linears))

    def frozen(self) -> bool:
        return not any(l.weight.requires_grad for l in self.linears)

    def tensors(self):
        key = tuple((l.weight.data_ptr(), l.weight._version) for l in self.linears)
        if key != self._key:
            W = torch.cat([l.weight.detach() for l in self.linears], 0)              # [3F, F]: out = x W_k^T  ->  B[n, k] = W_k
            Wt = torch.cat([l.weight.detach().t() for l in self.linears], 0)        # backward: dx = dy W_k   ->  B[n, k] = W_k^T
            self._fwd, self._bwd = _tf32_split(W), _tf32_split(Wt.contiguous())
            self._key = key
        return self._fwd, self._bwd


def _irreps_gemm(I, A, S, w_hi, w_lo):
    I, A, S = I.contiguous(), A.contiguous(), S.contiguous()
    N, F = I.shape[0], I.shape[-1]
    oI, oA, oS = _triple(N, F, I.device)
    _lib.call("tmd_irreps_linear_tc", p(I), p(A), p(S), N, F, p(w_hi), p(w_lo), p(oI), p(oA), p(oS), device=I.device)
    return oI, oA, oS


class _IrrepsLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, I, A, S, weights):
        fwd, bwd = weights.tensors()
        ctx.bwd = bwd
        return _irreps_gemm(I, A, S, fwd[0], fwd[1])

    @staticmethod
    @once_differentiable
    def backward(ctx, dI, dA, dS):
        gI, gA, gS = _irreps_gemm(dI, dA, dS, ctx.bwd[0], ctx.bwd[1])
        return gI, gA, gS, None


def irreps_linear(X, weights: IrrepsLinearWeights):
    """(W0 I, W1 A, W2 S) as one grouped tensor-core GEMM (row a9); gradients flow to I/A/S only (frozen weights)."""
    return _IrrepsLinear.apply(X[0], X[1], X[2], weights)


# ------------------------------------------------------------------------------------------ whole stages as ONE autograd node
# The frozen-weights frame differentiates w.r.t. geometry only, so a stage's backward is a fixed kernel sequence.  Written out by
# hand (instead of chaining one autograd Function per kernel) the sums that autograd would perform with separate `add` launches
# -- a tensor consumed by two kernels receives two gradients -- become part of a consumer kernel: the message-passing backward
# adds the gradient Y received from P2, the P1 backward takes its two incoming gradients as separate operands, the embedding's MLP
# backward adds what came back through the node linears, and P3's identity branch is not materialised at all.
def rows_gemm(A: Tensor, W: Tensor, epi: int, vec: Optional[Tensor] = None, aux: Optional[Tensor] = None, save_z: bool = False):
    """out [M,N] = epilogue(A [M,K] . W [K,N]) (csrc/rows_gemm.cu); epi 0: SiLU(acc + vec) (+ pre-activation if save_z),
    1: acc * SiLU'(aux), 2: acc * vec.  -> out, or (z, out) when save_z."""
    M, K = A.shape
    N = W.shape[1]
    out = torch.empty(M, N, device=A.device, dtype=torch.float32)
    z = torch.empty_like(out) if save_z else None
    _lib.call("tmd_rows_gemm", p(A), A.stride(0), p(W), W.stride(0), M, K, N, int(epi), p(vec), p(aux), 0 if aux is None else aux.stride(0),
              p(z), p(out), N, device=A.device)
    return (z, out) if save_z else out


# radial functions of all interaction layers ahead of time on a forked stream (TMD_TABLE_OVERLAP=0: inside each layer, in order)
USE_TABLE_OVERLAP = os.environ.get("TMD_TABLE_OVERLAP", "1") != "0"
_side_streams = {}


class LayerRadials:
    """(f, f') of every interaction layer, evaluated on a side stream forked from the caller's stream.

    With frozen weights the radial functions depend on the pair distances only, so all layers' table lookups (HBM-write-bound:
    0.85 GB per layer at C4) can run while the main stream does the issue- / latency-bound start of the frame (message-passing plan,
    moments embedding, embedding tail).  `get(i)` joins the side stream into the current one on first use.  Fork and join are plain
    stream waits, so the pattern is CUDA-graph capturable."""

    def __init__(self, chains, pair_r, num_unique, rbf_mod, lo, hi, want_tangent):
        cur = torch.cuda.current_stream(pair_r.device)
        side = _side_streams.get(pair_r.device)
        if side is None:
            side = _side_streams[pair_r.device] = torch.cuda.Stream(device=pair_r.device)
        self._side, self._joined, self.out = side, False, []
        P = pair_r.shape[0]
        tabs = [c.table(rbf_mod, lo, hi) for c in chains]
        for c, tab in zip(chains, tabs):     # outputs are allocated on the CALLER's stream (they are consumed and freed there);
            n_last = c.tensors()[0][-1].shape[1]   # the side stream only fills them between the fork and the join
            f = torch.empty(P, n_last, device=pair_r.device, dtype=torch.float32)
            self.out.append((f, torch.empty_like(f) if want_tangent else None))
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for tab, (f, fp) in zip(tabs, self.out):
                _lib.call("tmd_radial_table", p(pair_r), p(num_unique), P, p(tab[0]), tab[1], tab[2], f.shape[1], p(f), p(fp),
                          device=pair_r.device)

    def get(self, i):
        if not self._joined:
            torch.cuda.current_stream(self.out[0][0].device).wait_stream(self._side)
            self._joined = True
        return self.out[i]


def layer_radials(chains, pair_r, num_unique, rbf_mod, lo, hi, want_tangent):
    """LayerRadials when every chain is read from its table (the overlap pays for the HBM-bound lookup), else None"""
    if not (USE_TABLE_OVERLAP and USE_RADIAL_TABLE and len(chains) > 0 and pair_r.shape[0] > 0):
        return None
    if not all(c.table(rbf_mod, lo, hi) is not None for c in chains):
        return None
    return LayerRadials(chains, pair_r.detach(), num_unique, rbf_mod, lo, hi, want_tangent)


class _InteractionLayer(torch.autograd.Function):
    """One TensorNet interaction (reference tensornet.py:723-814): P1 -> node linears -> radial chain / table -> message passing ->
    P2 -> node linears -> P3.  Inputs X = (I, A, S) [N,{1,3,5},F] and the per-slot distances; output X'."""

    @staticmethod
    def forward(ctx, XI, XA, XS, dist, qf, o3, pair_r, chain, rbf_mod, lo, hi, g, pidx, rbf_op, plan, nl0, nl1, radial):
        XI, XA, XS, qf = XI.contiguous(), XA.contiguous(), XS.contiguous(), qf.contiguous()
        N, F, dev = XI.shape[0], XI.shape[-1], XI.device
        (w0, _), (w1, _) = nl0.tensors(), nl1.tensors()
        Xn = _triple(N, F, dev)
        _lib.call("tmd_p1_fwd", p(XI), p(XA), p(XS), p(Xn[0]), p(Xn[1]), p(Xn[2]), N, F, device=dev)
        Y = _irreps_gemm(Xn[0], Xn[1], Xn[2], w0[0], w0[1])
        if radial is not None:       # evaluated ahead of time on the forked stream (LayerRadials)
            f, fp = radial
        else:
            f, fp = run_chain(chain, pair_r.detach(), pidx.num_unique, rbf_mod, lo, hi, dist.requires_grad, rbf_op)
        if plan is None:
            plan = MessagePassingPlan(g, pidx, F)
        M = plan.forward(Y, f)
        Cn = _triple(N, F, dev)
        _lib.call("tmd_p2_fwd", p(Y[0]), p(Y[1]), p(Y[2]), p(M[0]), p(M[1]), p(M[2]), p(qf), int(o3), p(Cn[0]), p(Cn[1]), p(Cn[2]), N, F, device=dev)
        dX = _irreps_gemm(Cn[0], Cn[1], Cn[2], w1[0], w1[1])
        out = _triple(N, F, dev)
        _lib.call("tmd_p3_fwd", p(Xn[0]), p(Xn[1]), p(Xn[2]), p(dX[0]), p(dX[1]), p(dX[2]), p(qf), p(out[0]), p(out[1]), p(out[2]), N, F, device=dev)
        ctx.save_for_backward(XI, XA, XS, *Y, *M, *dX, f, fp if fp is not None else f, qf)
        ctx.cfg = (int(o3), plan, fp is not None, dist.shape[0], nl0, nl1)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dI, dA, dS):
        XI, XA, XS, YI, YA, YS, MI, MA, MS, DI, DA, DS, f, fp, qf = ctx.saved_tensors
        o3, plan, has_fp, E, nl0, nl1 = ctx.cfg
        N, F, dev = XI.shape[0], XI.shape[-1], XI.device
        dO = (dI.contiguous(), dA.contiguous(), dS.contiguous())
        (_, b0), (_, b1) = nl0.tensors(), nl1.tensors()
        gD = _triple(N, F, dev)
        _lib.call("tmd_p3_bwd_dx", p(DI), p(DA), p(DS), p(qf), p(dO[0]), p(dO[1]), p(dO[2]), p(gD[0]), p(gD[1]), p(gD[2]), N, F, device=dev)
        gC = _irreps_gemm(gD[0], gD[1], gD[2], b1[0], b1[1])
        gYa, gM = _triple(N, F, dev), _triple(N, F, dev)
        _lib.call("tmd_p2_bwd", p(YI), p(YA), p(YS), p(MI), p(MA), p(MS), p(qf), o3, p(gC[0]), p(gC[1]), p(gC[2]), p(gYa[0]), p(gYa[1]), p(gYa[2]),
                  p(gM[0]), p(gM[1]), p(gM[2]), N, F, device=dev)
        gY, g_dist = plan.backward((YI, YA, YS), gM, f, fp if has_fp else None, add=gYa)
        if g_dist is not None and g_dist.shape[0] != E:     # dynamic shapes: `dist` is the exact-size prefix of the capacity buffer
            g_dist = g_dist[:E]
        gXn = _irreps_gemm(gY[0], gY[1], gY[2], b0[0], b0[1])
        gX = _triple(N, F, dev)
        _lib.call("tmd_p1_bwd2", p(XI), p(XA), p(XS), p(dO[0]), p(dO[1]), p(dO[2]), p(gXn[0]), p(gXn[1]), p(gXn[2]), p(gX[0]), p(gX[1]), p(gX[2]),
                  N, F, device=dev)
        return (gX[0], gX[1], gX[2], g_dist) + (None,) * 14


def interaction_layer(X, dist, qf, o3, pair_r, chain, rbf_mod, lo, hi, g, pidx, rbf_op, plan, nl0, nl1, radial=None):
    return _InteractionLayer.apply(X[0], X[1], X[2], dist, qf, bool(o3), pair_r, chain, rbf_mod, lo, hi, g, pidx, rbf_op, plan, nl0, nl1,
                                   radial)


class EmbedMlpWeights:
    """init_norm + linears_scalar[0..1] of TensorEmbedding in the layouts of csrc/rows_gemm.cu, cached per weight version."""

    def __init__(self, init_norm: torch.nn.LayerNorm, lin0: torch.nn.Linear, lin1: torch.nn.Linear):
        self.mods = (init_norm, lin0, lin1)
        self._key = None

    def params(self):
        return [t for m in self.mods for t in (m.weight, m.bias)]

    def frozen(self) -> bool:
        return not any(t.requires_grad for t in self.params())

    def supported(self) -> bool:
        ln, l0, l1 = self.mods
        F = l0.in_features
        return (F % 32 == 0 and F <= 256 and l0.out_features == 2 * F and l1.in_features == 2 * F and l1.out_features == 3 * F and
                ln.elementwise_affine and tuple(ln.normalized_shape) == (F,) and l0.bias is not None and l1.bias is not None and
                all(t.dtype == torch.float32 for t in self.params()))

    def tensors(self):
        key = tuple((t.data_ptr(), t._version) for t in self.params())
        if key != self._key:
            ln, l0, l1 = self.mods
            d = lambda t: t.detach().contiguous()
            self._t = dict(ln_w=d(ln.weight), ln_b=d(ln.bias), eps=float(ln.eps), W1T=d(l0.weight.detach().t()), b1=d(l0.bias),
                           W2T=d(l1.weight.detach().t()), b2=d(l1.bias), W1=d(l0.weight), W2=d(l1.weight))
            self._key = key
        return self._t


class _EmbeddingTail(torch.autograd.Function):
    """Per-atom part of TensorEmbedding (tensornet.py:588-619): X' = (W_t X) * SiLU(MLP(LayerNorm(|X|^2))) as one autograd node:
    node linears (tcgen05), norm + LayerNorm, two GEMM launches with fused bias / SiLU, gate; the backward adds the gradient that
    came back through the node linears inside the norm backward (csrc/rows_gemm.cu, tensor_ops.cu)."""

    @staticmethod
    def forward(ctx, I, A, S, nl, mw):
        I, A, S = I.contiguous(), A.contiguous(), S.contiguous()
        N, F, dev = I.shape[0], I.shape[-1], I.device
        (wf, _), t = nl.tensors(), mw.tensors()
        C = _irreps_gemm(I, A, S, wf[0], wf[1])
        xn = torch.empty(N, F, device=dev, dtype=torch.float32)
        mu_rstd = torch.empty(N, 2, device=dev, dtype=torch.float32)
        _lib.call("tmd_norm_ln_fwd", p(I), p(A), p(S), N, F, 2, p(t["ln_w"]), p(t["ln_b"]), t["eps"], p(xn), p(mu_rstd), device=dev)
        z1, h1 = rows_gemm(xn, t["W1T"], 0, vec=t["b1"], save_z=True)
        z2, gact = rows_gemm(h1, t["W2T"], 0, vec=t["b2"], save_z=True)
        out = _triple(N, F, dev)
        _lib.call("tmd_gate_fwd", p(C[0]), p(C[1]), p(C[2]), p(gact), p(out[0]), p(out[1]), p(out[2]), N, F, device=dev)
        ctx.save_for_backward(I, A, S, *C, mu_rstd, z1, z2)
        ctx.w = (nl, t)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dI, dA, dS):
        I, A, S, CI, CA, CS, mu_rstd, z1, z2 = ctx.saved_tensors
        nl, t = ctx.w
        N, F, dev = I.shape[0], I.shape[-1], I.device
        dO = (dI.contiguous(), dA.contiguous(), dS.contiguous())
        (_, wb) = nl.tensors()
        dC = _triple(N, F, dev)
        dz2 = torch.empty(N, 3 * F, device=dev, dtype=torch.float32)
        _lib.call("tmd_gate_bwd_preact", p(CI), p(CA), p(CS), p(z2), p(dO[0]), p(dO[1]), p(dO[2]), p(dC[0]), p(dC[1]), p(dC[2]), p(dz2), 1, N, F,
                  device=dev)
        gXa = _irreps_gemm(dC[0], dC[1], dC[2], wb[0], wb[1])
        dz1 = rows_gemm(dz2, t["W2"], 1, aux=z1)
        g = rows_gemm(dz1, t["W1"], 2, vec=t["ln_w"])
        gX = _triple(N, F, dev)
        _lib.call("tmd_norm_ln_bwd", p(I), p(A), p(S), N, F, 2, p(g), p(mu_rstd), p(gXa[0]), p(gXa[1]), p(gXa[2]), p(gX[0]), p(gX[1]), p(gX[2]),
                  device=dev)
        return gX[0], gX[1], gX[2], None, None


def embedding_tail(X, nl: IrrepsLinearWeights, mw: EmbedMlpWeights):
    return _EmbeddingTail.apply(X[0], X[1], X[2], nl, mw)


# ------------------------------------------------------------------------------------------ frozen nn.Linear on the GEMM launcher
class _RowsLinear(torch.autograd.Function):
    """y = x W^T + b for a frozen nn.Linear on tmd_rows_gemm (3xTF32 tensor-core tiles, fp32 parity) instead of a cuBLAS SIMT SGEMM;
    backward w.r.t. x only."""

    @staticmethod
    def forward(ctx, x2d, Wt, W, bias):
        ctx.W = W
        return rows_gemm(x2d.contiguous(), Wt, 3, vec=bias)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        return rows_gemm(g.contiguous(), ctx.W, 3), None, None, None


def rows_linear(x: Tensor, lin: torch.nn.Linear) -> Tensor:
    """lin(x) over the last dimension through _RowsLinear when the layer is frozen and its widths fit the tiles, else lin(x)."""
    K, N = lin.in_features, lin.out_features
    if lin.weight.requires_grad or K % 16 or N % 32 or x.dtype != torch.float32 or not x.is_cuda or lin.weight.dtype != torch.float32:
        return lin(x)
    key = (lin.weight.data_ptr(), lin.weight._version, None if lin.bias is None else lin.bias._version)
    cache = getattr(lin, "_rows_w", None)
    if cache is None or cache[0] != key:
        cache = (key, lin.weight.detach().t().contiguous(), lin.weight.detach().contiguous(),
                 None if lin.bias is None else lin.bias.detach().contiguous())
        lin._rows_w = cache
    y = _RowsLinear.apply(x.reshape(-1, K), cache[1], cache[2], cache[3])
    return y.view(*x.shape[:-1], N)


# ------------------------------------------------------------------------------------------ ET output head (row a15)
def _linear_rows(A: Tensor, Wt: Tensor, bias: Optional[Tensor] = None) -> Tensor:
    """A [M,K] . Wt [K,Nout] (+ bias): the weight-tiled GEMM, or the warp-per-row kernel for layers with <= 8 outputs"""
    if Wt.shape[1] % 32 == 0:
        return rows_gemm(A, Wt, 3, vec=bias)
    M, K = A.shape
    out = torch.empty(M, Wt.shape[1], device=A.device, dtype=torch.float32)
    _lib.call("tmd_rows_narrow_fwd", p(A), A.stride(0), p(Wt), M, K, Wt.shape[1], p(bias), p(out), device=A.device)
    return out


class GatedBlockWeights:
    """One GatedEquivariantBlock (reference models/utils.py:583-655) in the layouts of csrc/rows_gemm.cu, cached per weight version."""

    def __init__(self, block):
        self.block = block
        self._key = None

    def params(self):
        b = self.block
        lin_a, lin_b = b.update_net.layers[0], b.update_net.layers[-1]
        return [b.vec1_proj.weight, b.vec2_proj.weight, lin_a.weight, lin_a.bias, lin_b.weight, lin_b.bias]

    def frozen(self) -> bool:
        return not any(t.requires_grad for t in self.params())

    def supported(self) -> bool:
        b = self.block
        layers = b.update_net.layers
        if len(layers) != 3 or not isinstance(layers[1], torch.nn.SiLU) or (b.act is not None and not isinstance(b.act, torch.nn.SiLU)):
            return False
        Hh, O, I = b.vec1_proj.in_features, b.out_channels, layers[0].out_features
        wide = lambda n: n % 32 == 0
        return (Hh % 32 == 0 and I % 32 == 0 and layers[0].in_features == 2 * Hh and layers[-1].out_features == 2 * O and
                (wide(O) and O % 16 == 0 or O <= 4) and all(t.dtype == torch.float32 for t in self.params()))

    def tensors(self):
        key = tuple((t.data_ptr(), t._version) for t in self.params())
        if key != self._key:
            b = self.block
            lin_a, lin_b = b.update_net.layers[0], b.update_net.layers[-1]
            d = lambda t: t.detach().contiguous()
            self._t = dict(W1T=d(b.vec1_proj.weight.detach().t()), W1=d(b.vec1_proj.weight), W2T=d(b.vec2_proj.weight.detach().t()),
                           W2=d(b.vec2_proj.weight), WaT=d(lin_a.weight.detach().t()), Wa=d(lin_a.weight), ba=d(lin_a.bias),
                           WbT=d(lin_b.weight.detach().t()), Wb=d(lin_b.weight), bb=d(lin_b.bias), O=b.out_channels,
                           act=int(b.act is not None))
            self._key = key
        return self._t


def _geb_forward(x, v, t):
    """one gated block: x [N,Hh], v [N,3,Hh] -> (xo [N,O], vo [N,3,O], saved tensors)"""
    N, Hh, O, dev = x.shape[0], x.shape[1], t["O"], x.device
    v2d = v.reshape(3 * N, Hh)
    P1 = _linear_rows(v2d, t["W1T"])
    V2 = _linear_rows(v2d, t["W2T"])
    cat = torch.empty(N, 2 * Hh, device=dev, dtype=torch.float32)
    _lib.call("tmd_geb_cat_fwd", p(x), p(P1), N, Hh, p(cat), device=dev)
    z1, h1 = rows_gemm(cat, t["WaT"], 0, vec=t["ba"], save_z=True)
    hh = _linear_rows(h1, t["WbT"], t["bb"])
    xo = torch.empty(N, O, device=dev, dtype=torch.float32)
    vo = torch.empty(N, 3, O, device=dev, dtype=torch.float32)
    _lib.call("tmd_geb_out_fwd", p(hh), p(V2), N, O, t["act"], p(xo), p(vo), device=dev)
    return xo, vo, (P1, V2, cat, z1, hh)


def _geb_backward(saved, t, g_xo, g_vo):
    """-> (g_x as a strided [N,Hh] view of g_cat, g_v [N,3,Hh]); g_xo may be a strided view (row stride passed on), g_vo contiguous"""
    P1, V2, cat, z1, hh = saved
    N, Hh, O, I, dev = cat.shape[0], cat.shape[1] // 2, t["O"], z1.shape[1], cat.device
    g_hh = torch.empty(N, 2 * O, device=dev, dtype=torch.float32)
    g_V2 = torch.empty(3 * N, O, device=dev, dtype=torch.float32)
    _lib.call("tmd_geb_out_bwd", p(hh), p(V2), p(g_xo), 0 if g_xo is None else g_xo.stride(0), p(g_vo), N, O, t["act"], p(g_hh), p(g_V2),
              device=dev)
    if (2 * O) % 32 == 0:
        g_h1 = rows_gemm(g_hh, t["Wb"], 1, aux=z1)
    else:
        g_h1 = torch.empty(N, I, device=dev, dtype=torch.float32)
        _lib.call("tmd_rows_narrow_bwd", p(g_hh), p(t["Wb"]), N, I, 2 * O, 0, p(z1), p(g_h1), device=dev)
    g_cat = rows_gemm(g_h1, t["Wa"], 3)
    g_P1 = torch.empty(3 * N, Hh, device=dev, dtype=torch.float32)
    _lib.call("tmd_geb_cat_bwd", p(P1), p(cat), p(g_cat), N, Hh, p(g_P1), device=dev)
    if O % 32 == 0:
        g_v = rows_gemm(g_P1, t["W1"], 3)
        g_v = rows_gemm(g_V2, t["W2"], 4, aux=g_v)
    else:
        g_v = rows_gemm(g_P1, t["W1"], 3)
        _lib.call("tmd_rows_narrow_bwd", p(g_V2), p(t["W2"]), 3 * N, Hh, O, 1, None, p(g_v), device=dev)
    return g_cat[:, :Hh], g_v.view(N, 3, Hh)


class _ETHead(torch.autograd.Function):
    """EquivariantScalar.pre_reduce (two gated equivariant blocks) -> * std -> + atomref[z] -> per-molecule sum -> + mean, as one
    autograd node with a hand-sequenced backward w.r.t. (x, vec) -- frozen weights (row a15 + the reduce of a11)."""

    @staticmethod
    def forward(ctx, x, v, z, batch, M, blocks, std, mean, atomref):
        x, v = x.contiguous(), v.contiguous()
        dev, N = x.device, x.shape[0]
        ts = [b.tensors() for b in blocks]
        saved = []
        for t in ts:
            x, v, sv = _geb_forward(x, v, t)
            saved.append(sv)
        std_c, mean_c = std.detach().reshape(1).contiguous(), mean.detach().reshape(1).contiguous()
        aref = None if atomref is None else atomref.detach().contiguous()
        one, zero = torch.ones(1, device=dev), torch.zeros(1, device=dev)
        e = torch.empty(N, device=dev, dtype=torch.float32)
        z_c, batch_c = z.contiguous(), batch.contiguous()
        _lib.call("tmd_head_energy", p(x), N, 1, p(one), p(zero), p(std_c), p(z_c), p(aref), 0 if aref is None else aref.shape[0], p(e), device=dev)
        y = torch.empty(M, 1, device=dev, dtype=torch.float32)
        flag = torch.empty(1, device=dev, dtype=torch.int32)
        _lib.call("tmd_segment_sum", p(e), p(batch_c), N, M, p(mean_c), p(flag), p(y), device=dev)
        flat = [tt for sv in saved for tt in sv]
        ctx.save_for_backward(batch_c, std_c, *flat)
        ctx.ts, ctx.nsv = ts, len(saved[0])
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        batch, std_c, *flat = ctx.saved_tensors
        ts, k = ctx.ts, ctx.nsv
        saved = [tuple(flat[i * k:(i + 1) * k]) for i in range(len(ts))]
        g_x = (gy.reshape(-1).index_select(0, batch) * std_c).reshape(-1, 1).contiguous()     # d y / d x_final
        g_v = None
        for t, sv in zip(reversed(ts), reversed(saved)):
            g_x, g_v = _geb_backward(sv, t, g_x, None if g_v is None else g_v.contiguous())
        return g_x, g_v, None, None, None, None, None, None, None


def et_head_energy(x, v, z, batch, num_systems: int, blocks, std, mean, atomref=None):
    return _ETHead.apply(x, v, z, batch, int(num_systems), blocks, std, mean, atomref)


# ------------------------------------------------------------------------------------------ tail + head + reduce (rows a10, a11)
class HeadWeights:
    """Contiguous (and transposed) copies of the weights behind norm3 -> LayerNorm -> Linear -> SiLU -> MLP(F -> F/2 -> 1), cached
    per weight version (csrc/head.cu reads W k-major in the forward and untransposed in the backward)."""

    def __init__(self, out_norm: torch.nn.LayerNorm, linear: torch.nn.Linear, lin_a: torch.nn.Linear, lin_b: torch.nn.Linear):
        self.mods = (out_norm, linear, lin_a, lin_b)
        self._key = None

    def params(self):
        return [t for m in self.mods for t in (m.weight, m.bias)]

    def frozen(self) -> bool:
        return not any(t.requires_grad for t in self.params())

    def supported(self) -> bool:
        out_norm, linear, lin_a, lin_b = self.mods
        F, H = linear.out_features, lin_a.out_features
        smem_bwd = (32 * ((3 * F + 4) + (F + 4) + (H + 4)) + 3 * 16 * (max(F, H) + 8)) * 4      # tile + weight-staging buffers of the head kernels (csrc/head.cu, mlp_tile.cuh)
        return (linear.in_features == 3 * F and lin_a.in_features == F and lin_b.in_features == H and lin_b.out_features == 1 and
                F % 32 == 0 and H % 32 == 0 and F <= 256 and H <= 256 and smem_bwd <= 227 * 1024 and out_norm.elementwise_affine and
                tuple(out_norm.normalized_shape) == (3 * F,) and all(t.dtype == torch.float32 for t in self.params()))

    def tensors(self):
        key = tuple((t.data_ptr(), t._version) for t in self.params())
        if key != self._key:
            out_norm, linear, lin_a, lin_b = self.mods
            d = lambda t: t.detach().contiguous()
            self._t = dict(ln_w=d(out_norm.weight), ln_b=d(out_norm.bias), W1=d(linear.weight), W1T=d(linear.weight.detach().t()),
                           b1=d(linear.bias), W2=d(lin_a.weight), W2T=d(lin_a.weight.detach().t()), b2=d(lin_a.bias),
                           w3=d(lin_b.weight).view(-1), b3=d(lin_b.bias).view(-1), eps=float(out_norm.eps),
                           F=linear.out_features, H=lin_a.out_features)
            self._key = key
        return self._t


class _TailHead(torch.autograd.Function):
    """y[m] = mean + sum_{i in m} ((w3 . SiLU(W2 SiLU(W1 LN(norm3(X_i)) + b1) + b2) + b3) * std + atomref[z_i]); backward w.r.t.
    the node tensors only (frozen weights)."""

    @staticmethod
    def forward(ctx, I, A, S, z, batch, M, hw, std, mean, atomref):
        I, A, S = I.contiguous(), A.contiguous(), S.contiguous()
        t = hw.tensors()
        N, F, H, dev = I.shape[0], t["F"], t["H"], I.device
        e = torch.empty(N, device=dev, dtype=torch.float32)
        mu_rstd = torch.empty(N, 2, device=dev, dtype=torch.float32)
        z1 = torch.empty(N, F, device=dev, dtype=torch.float32)
        z2 = torch.empty(N, H, device=dev, dtype=torch.float32)
        y = torch.empty(M, 1, device=dev, dtype=torch.float32)
        flag = torch.empty(1, device=dev, dtype=torch.int32)
        batch_c, z_c = batch.contiguous(), z.contiguous()
        std_c, mean_c = std.detach().reshape(1).contiguous(), mean.detach().reshape(1).contiguous()
        aref = None if atomref is None else atomref.detach().contiguous()
        if USE_ROWS_GEMM:   # norm3 + LayerNorm, two GEMM launches (bias + SiLU fused, pre-activations saved), energy per atom
            xn = torch.empty(N, 3 * F, device=dev, dtype=torch.float32)
            _lib.call("tmd_norm_ln_fwd", p(I), p(A), p(S), N, F, 3, p(t["ln_w"]), p(t["ln_b"]), t["eps"], p(xn), p(mu_rstd), device=dev)
            z1, h1 = rows_gemm(xn, t["W1T"], 0, vec=t["b1"], save_z=True)
            z2, h2 = rows_gemm(h1, t["W2T"], 0, vec=t["b2"], save_z=True)
            _lib.call("tmd_head_energy", p(h2), N, H, p(t["w3"]), p(t["b3"]), p(std_c), p(z_c), p(aref), 0 if aref is None else aref.shape[0],
                      p(e), device=dev)
        else:
            _lib.call("tmd_head_fwd", p(I), p(A), p(S), N, F, H, p(t["ln_w"]), p(t["ln_b"]), t["eps"], p(t["W1T"]), p(t["b1"]), p(t["W2T"]),
                      p(t["b2"]), p(t["w3"]), p(t["b3"]), p(std_c), p(z_c), p(aref), 0 if aref is None else aref.shape[0], p(e), p(mu_rstd),
                      p(z1), p(z2), device=dev)
        _lib.call("tmd_segment_sum", p(e), p(batch_c), N, M, p(mean_c), p(flag), p(y), device=dev)
        ctx.save_for_backward(I, A, S, batch_c, mu_rstd, z1, z2, std_c)
        ctx.hw, ctx.M = t, M
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        I, A, S, batch, mu_rstd, z1, z2, std_c = ctx.saved_tensors
        t = ctx.hw
        N, F, H = I.shape[0], t["F"], t["H"]
        gy = gy.contiguous()
        dX = _triple(N, F, I.device)
        if USE_ROWS_GEMM:
            dz2 = torch.empty(N, H, device=I.device, dtype=torch.float32)
            _lib.call("tmd_head_seed", p(gy), p(batch), ctx.M, p(std_c), p(t["w3"]), p(z2), N, H, p(dz2), device=I.device)
            dz1 = rows_gemm(dz2, t["W2"], 1, aux=z1)
            g = rows_gemm(dz1, t["W1"], 2, vec=t["ln_w"])
            _lib.call("tmd_norm_ln_bwd", p(I), p(A), p(S), N, F, 3, p(g), p(mu_rstd), None, None, None, p(dX[0]), p(dX[1]), p(dX[2]),
                      device=I.device)
        else:
            _lib.call("tmd_head_bwd", p(I), p(A), p(S), N, F, H, p(t["ln_w"]), t["eps"], p(t["W1"]), p(t["W2"]), p(t["w3"]), p(std_c), p(gy),
                      p(batch), ctx.M, p(mu_rstd), p(z1), p(z2), p(dX[0]), p(dX[1]), p(dX[2]), device=I.device)
        return dX[0], dX[1], dX[2], None, None, None, None, None, None, None


def tail_head_energy(X, z, batch, num_systems: int, hw: HeadWeights, std, mean, atomref=None):
    """per-molecule energies [M,1] from the final node tensors X = (I, A, S): rows a10 + a11 in three launches."""
    return _TailHead.apply(X[0], X[1], X[2], z, batch, int(num_systems), hw, std, mean, atomref)
