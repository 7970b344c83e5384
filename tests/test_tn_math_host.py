"""CPU check of the hand-derived per-(atom,channel) arithmetic in torchmdnet_b200/csrc/tn_math.h.

The header is compiled with g++ (double) into a throw-away test library and every forward /
backward is compared with the reference formulas (oracle/tensornet_oracle.py, which follows
models/tensornet.py) differentiated by torch.autograd.  This is how kernel math is validated on
the GPU-less build box; the same header is what the sm_100a kernels inline.
"""
import ctypes
import os
import subprocess
import tempfile

import pytest
import torch

import tensornet_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
P = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def hc():
    out = os.path.join(tempfile.mkdtemp(prefix="hostcheck"), "libhostcheck.so")
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-o", out, os.path.join(HERE, "hostcheck", "hostcheck.cpp")])
    return ctypes.CDLL(out)


def ptr(t):
    assert t.dtype == torch.float64 and t.is_contiguous()
    return ctypes.cast(t.data_ptr(), P)


def full_from_compact(c):  # c [n,9] -> X [n,3,3,1] with the oracle's conventions
    I = c[:, 0:1]
    A = orc.skew(torch.stack([-c[:, 3], c[:, 2], -c[:, 1]], 1)[:, :, None])  # v with A01=-vz, A02=vy, A12=-vx
    S = torch.zeros(c.shape[0], 3, 3, 1, dtype=c.dtype)
    S[:, 0, 0, 0] = c[:, 4]; S[:, 0, 1, 0] = c[:, 5]; S[:, 1, 0, 0] = c[:, 5]
    S[:, 0, 2, 0] = c[:, 6]; S[:, 2, 0, 0] = c[:, 6]; S[:, 1, 1, 0] = c[:, 7]
    S[:, 1, 2, 0] = c[:, 8]; S[:, 2, 1, 0] = c[:, 8]; S[:, 2, 2, 0] = -(c[:, 4] + c[:, 7])
    return orc.compose(I, A, S)


def compact_from_full(X):  # X [n,3,3,1] -> [n,9]
    I, A, S = orc.decompose(X)
    return torch.stack([I[:, 0], A[:, 0, 1, 0], A[:, 0, 2, 0], A[:, 1, 2, 0], S[:, 0, 0, 0], S[:, 0, 1, 0],
                        S[:, 0, 2, 0], S[:, 1, 1, 0], S[:, 1, 2, 0]], 1)


N = 64
g = torch.Generator().manual_seed(0)


def rnd(*s):
    return torch.randn(*s, generator=g, dtype=torch.float64)


def close(a, b, tol=1e-12):
    assert (a - b).abs().max().item() <= tol * (1 + b.abs().max().item()), (a - b).abs().max().item()


def test_compose_decompose(hc):
    c = rnd(N, 9); X = torch.empty(N, 9, dtype=torch.float64)
    hc.hc_compose(ptr(c), ptr(X), N)
    close(X, full_from_compact(c).reshape(N, 9))
    c2 = torch.empty_like(c); hc.hc_decompose(ptr(X), ptr(c2), N); close(c2, c)
    Xr = rnd(N, 9); cr = torch.empty(N, 9, dtype=torch.float64); hc.hc_decompose(ptr(Xr), ptr(cr), N)
    close(cr, compact_from_full(Xr.view(N, 3, 3, 1)))
    # backward operators are the transposes
    dX = rnd(N, 9); dc = torch.empty_like(c); hc.hc_compose_bwd(ptr(dX), ptr(dc), N)
    cc = c.clone().requires_grad_(True); (full_from_compact(cc).reshape(N, 9) * dX).sum().backward(); close(dc, cc.grad)
    dcc = rnd(N, 9); dXo = torch.empty_like(dX); hc.hc_decompose_bwd(ptr(dcc), ptr(dXo), N)
    Xg = Xr.clone().requires_grad_(True); (compact_from_full(Xg.view(N, 3, 3, 1)) * dcc).sum().backward(); close(dXo, Xg.grad)


@pytest.mark.parametrize("o3", [True, False])
def test_matmul(hc, o3):
    A, B, dC = rnd(N, 9), rnd(N, 9), rnd(N, 9)
    C = torch.empty_like(A); dA = torch.empty_like(A); dB = torch.empty_like(A)
    (hc.hc_mm_o3 if o3 else hc.hc_mm_so3)(ptr(A), ptr(B), ptr(C), N)
    (hc.hc_mm_o3_bwd if o3 else hc.hc_mm_so3_bwd)(ptr(A), ptr(B), ptr(dC), ptr(dA), ptr(dB), N)
    Ag, Bg = A.clone().requires_grad_(True), B.clone().requires_grad_(True)
    Af, Bf = Ag.view(N, 3, 3, 1), Bg.view(N, 3, 3, 1)
    ref = orc.mm3(Af, Bf) + orc.mm3(Bf, Af) if o3 else orc.mm3(Af, Bf)
    close(C, ref.reshape(N, 9).detach())
    (ref.reshape(N, 9) * dC).sum().backward(); close(dA, Ag.grad); close(dB, Bg.grad)


def test_norms(hc):
    c = rnd(N, 9); o = torch.empty(N, 3, dtype=torch.float64); hc.hc_norm3(ptr(c), ptr(o), N)
    cg = c.clone().requires_grad_(True)
    I, A, S = orc.decompose(full_from_compact(cg))
    ref = torch.cat([3 * I ** 2, orc.tnorm(A), orc.tnorm(S)], -1)
    close(o, ref.detach())
    dn = rnd(N, 3); dc = torch.empty_like(c); hc.hc_norm3_bwd(ptr(c), ptr(dn), ptr(dc), N)
    (ref * dn).sum().backward(); close(dc, cg.grad)
    n2 = torch.empty(N, dtype=torch.float64); hc.hc_norm2(ptr(c), ptr(n2), N)
    close(n2, orc.tnorm(full_from_compact(c))[:, 0])


def test_p1(hc):
    c, d = rnd(N, 9), rnd(N, 9); o = torch.empty_like(c); dc = torch.empty_like(c)
    hc.hc_p1(ptr(c), ptr(o), N); hc.hc_p1_bwd(ptr(c), ptr(d), ptr(dc), N)
    cg = c.clone().requires_grad_(True); X = full_from_compact(cg)
    ref = compact_from_full(X / (orc.tnorm(X) + 1)[:, None, None, :])
    close(o, ref.detach()); (ref * d).sum().backward(); close(dc, cg.grad)


@pytest.mark.parametrize("o3", [True, False])
def test_p2(hc, o3):
    Y, M, d = rnd(N, 9), rnd(N, 9), rnd(N, 9); qf = 1 + 0.1 * rnd(N)
    o = torch.empty_like(Y); dY = torch.empty_like(Y); dM = torch.empty_like(Y)
    hc.hc_p2(ptr(Y), ptr(M), ptr(qf), int(o3), ptr(o), N)
    hc.hc_p2_bwd(ptr(Y), ptr(M), ptr(qf), int(o3), ptr(d), ptr(dY), ptr(dM), N)
    Yg, Mg = Y.clone().requires_grad_(True), M.clone().requires_grad_(True)
    Yf, Mf = full_from_compact(Yg), full_from_compact(Mg)
    C = qf[:, None, None, None] * (orc.mm3(Mf, Yf) + orc.mm3(Yf, Mf)) if o3 else 2 * orc.mm3(Yf, Mf)
    ref = compact_from_full(C / (orc.tnorm(C) + 1)[:, None, None, :])
    close(o, ref.detach()); (ref * d).sum().backward(); close(dY, Yg.grad); close(dM, Mg.grad)


def test_p3(hc):
    Xn, dX, d = rnd(N, 9), rnd(N, 9), rnd(N, 9); qf = 1 + 0.1 * rnd(N)
    o = torch.empty_like(Xn); g1 = torch.empty_like(Xn); g2 = torch.empty_like(Xn)
    hc.hc_p3(ptr(Xn), ptr(dX), ptr(qf), ptr(o), N); hc.hc_p3_bwd(ptr(dX), ptr(qf), ptr(d), ptr(g1), ptr(g2), N)
    a, b = Xn.clone().requires_grad_(True), dX.clone().requires_grad_(True)
    D = full_from_compact(b)
    ref = compact_from_full(full_from_compact(a) + D + qf[:, None, None, None] * orc.mm3(D, D))
    close(o, ref.detach()); (ref * d).sum().backward(); close(g1, a.grad); close(g2, b.grad)


def test_radial_functions(hc):
    r = torch.rand(200, generator=g, dtype=torch.float64) * 6.0
    v = torch.empty_like(r); dv = torch.empty_like(r)
    xs = (r - 3).contiguous()
    hc.hc_silu(ptr(xs), ptr(v), ptr(dv), 200)
    x = xs.clone().requires_grad_(True); y = torch.nn.functional.silu(x); y.sum().backward()
    close(v, y.detach()); close(dv, x.grad)
    for lo, hi in ((0.0, 5.0), (1.0, 5.5)):
        c = torch.empty_like(r); dc = torch.empty_like(r)
        hc.hc_cutoff(ptr(r), ctypes.c_double(lo), ctypes.c_double(hi), ptr(c), ptr(dc), 200)
        rg = r.clone().requires_grad_(True); ref = orc.cosine_cutoff(rg, lo, hi); ref.sum().backward()
        close(c, ref.detach()); close(dc, rg.grad)
        R = 16
        mu, beta = orc.expnorm_initial_params(lo, hi, R, torch.float64)
        val = torch.empty(200, R, dtype=torch.float64); dval = torch.empty_like(val)
        hc.hc_expnorm(ptr(r), ctypes.c_double(lo), ctypes.c_double(hi), ptr(mu), ptr(beta), R, ptr(val), ptr(dval), 200)
        rg = r.clone().requires_grad_(True); ref = orc.expnorm_rbf(rg, mu, beta, lo, hi)
        close(val, ref.detach())
        w = rnd(200, R); (ref * w).sum().backward()
        close((dval * w).sum(1), rg.grad, 1e-11)


def test_sym_basis(hc):
    r = rnd(N, 3); s = torch.empty(N, 5, dtype=torch.float64); hc.hc_sym_basis(ptr(r), ptr(s), N)
    rg = r.clone().requires_grad_(True)
    T = rg[:, :, None] * rg[:, None, :]
    S = 0.5 * (T + T.transpose(1, 2)) - T.diagonal(dim1=1, dim2=2).sum(-1)[:, None, None] / 3 * torch.eye(3, dtype=torch.float64)
    ref = torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2]], 1)
    close(s, ref.detach())
    dS = rnd(N, 5); w = rnd(N); gr = torch.empty(N, 3, dtype=torch.float64)
    hc.hc_sym_basis_bwd(ptr(r), ptr(dS), ptr(w), ptr(gr), N)
    (ref * dS * w[:, None]).sum().backward(); close(gr, rg.grad)
