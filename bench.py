#!/usr/bin/env python
"""Headline benchmark: frames/s (energy + forces) of TensorNet on a 10k-atom periodic water box.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU (oracle port)

Workload C4 (BASELINE.json configs[3], the one the metric is quoted on): TensorNet 3 layers, F=128, 32 expnorm RBFs,
cutoff 5 A, 10 125-atom water box (15^3 H2O, L = 46.42 A, rho = 0.101 /A^3, ~5.5e5 directed edges incl. self loops),
random-init weights (seed 1), synthetic positions (seed = rank).  One step = one frame = energy + forces for the
whole box.  N > 1: one box per GPU (weak scaling, independent boxes, seeds differ), one NCCL all-reduce of the scalar
total energy per step and nothing else; value = N frames / max-over-ranks device time.

Timing: W untimed warm-up steps, then exactly K steps between CUDA events on the launching stream, bracketed by
barrier + synchronize; max over ranks.  The per-frame working set (~4 GB of edge tensors) is >> the 126 MB L2, so no
explicit L2 flush is needed between steps (stated in config.l2).  Clocks are sampled with nvidia-smi during the
timed region.  `e2e` repeats the measurement through the public API with HOST inputs (pinned) -> H2D -> model ->
D2H of energy+forces every step.
"""
import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import systems  # noqa: E402  (synthetic inputs; shared with tests and the golden generator)

METRIC = "frames_per_sec_energy_forces_tensornet_10k_atom_box"
UNIT = "frames/s"


def workload(name, seed=0):
    if name == "c4":
        hp = systems.tensornet_hparams(num_layers=3, F=128, R=32, cutoff=5.0, max_nb=64)
        z, pos, batch, box = systems.water_box(n_side=15, L=46.42, seed=seed)
        desc = "C4: TensorNet 3-layer F=128 R=32 rc=5.0, 10125-atom water box L=46.42A (15^3 H2O), energy+forces"
    elif name == "c2":
        hp = systems.tensornet_hparams(num_layers=2, F=128, R=32, cutoff=5.0, max_nb=128)
        z, pos, batch, box = systems.random_box(1000, seed=seed)
        desc = "C2: TensorNet 2-layer F=128 R=32 rc=5.0, 1000-atom random periodic box L=21.54A, energy+forces"
    elif name == "c3":
        hp = systems.et_hparams()
        z, pos, batch, box = systems.molecule_batch(n_mol=256, seed=seed)
        desc = "C3: ET-SPICE (5 layers, F=128, H=8, R=64, rc=10), 256 molecules x 20-40 atoms, energy+forces"
    elif name == "c5":
        hp = systems.tensornet_hparams(num_layers=2, F=128, R=32, cutoff=5.0, max_nb=128)
        z, pos, batch, box = systems.random_box(5000, seed=seed, n_boxes=8)
        desc = "C5 (per GPU): TensorNet-ANI2X config (2 layers, F=128), 8 x 5000-atom periodic boxes, energy+forces"
    else:
        raise ValueError(name)
    return hp, z, pos, batch, box, desc


def seeded_model(hp, seed=1):
    from torchmdnet_b200.models.model import create_model
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    return create_model(dict(hp)).eval()


class ClockSampler:
    """Samples SM clock / power / throttle reasons with ONE streaming nvidia-smi process (-lms 100) during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.index, self.proc = [], index, None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.35)  # first sample lands before the timed region starts
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill(); out = ""
            for line in out.strip().splitlines():
                c = [x.strip() for x in line.split(",")]
                if len(c) >= 7:
                    self.rows.append(c)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows if len(r) > 3 + k)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows), "reasons": reasons}


def measured_peaks():
    """(HBM GB/s, dense bf16 TFLOP/s burst, source) from the driver-written MEASURED_PEAKS.json, else the recipe's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def cpu_oracle_frames_per_s(hp, n_side, L, steps, threads, full_atoms):
    """Times the CPU oracle (port of the reference algorithm) on a bounded sub-box of the same workload and scales
    frames/s by the atom ratio (cost is linear in atoms at fixed density)."""
    import tensornet_oracle as orc
    sd = {k: v.detach().clone() for k, v in seeded_model(hp).state_dict().items()}
    # pick the thread count that is fastest on this host (many-core hosts oversubscribe on these tensor sizes)
    zc, pc, bc, xc = systems.water_box(n_side=4, L=46.42 * 4 / 15, seed=0)
    best = (None, 1e30)
    for th in sorted({threads, min(threads, 64), min(threads, 32), min(threads, 16), min(threads, 8)}, reverse=True):
        torch.set_num_threads(th)
        orc.energy_and_forces(sd, hp, zc, pc, bc, xc)
        t0 = time.perf_counter(); orc.energy_and_forces(sd, hp, zc, pc, bc, xc); dt = time.perf_counter() - t0
        if dt < best[1]:
            best = (th, dt)
    threads = best[0]
    torch.set_num_threads(threads)
    z, pos, batch, box = systems.water_box(n_side=n_side, L=L, seed=0)
    orc.energy_and_forces(sd, hp, z, pos, batch, box)  # warm-up
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        orc.energy_and_forces(sd, hp, z, pos, batch, box)
        ts.append(time.perf_counter() - t0)
    t = statistics.median(ts)
    return (1.0 / t) * (z.numel() / full_atoms), z.numel(), t, threads


def run_reference(args):
    """--impl reference: the reference's algorithm on the box's host cores (CPU oracle port), same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    hp, z, *_rest, desc = workload(args.workload)
    cores = os.cpu_count() or 1
    n_side = 6 if args.workload == "c4" else 4
    L = 46.42 * n_side / 15
    t0 = time.time()
    fps, n_sample, t, cores = cpu_oracle_frames_per_s(hp, n_side, L, max(1, min(args.steps, 5)), cores, z.numel())
    out = {"metric": METRIC, "value": fps, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 / fps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": {"workload": desc},
           "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{n_sample}-atom sub-box of the same water system (n_side={n_side}), {t:.2f} s/frame, scaled by atoms ratio"},
           "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "wall_s": time.time() - t0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c4", "c2", "c3", "c5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="launch the ~250 kernels of a frame from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist_on = world > 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if dist_on:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if dist_on:
        dist.barrier()
    from torchmdnet_b200 import _lib
    _lib.library()  # fail loudly if the extension is missing

    hp, z, pos, batch, box, desc = workload(args.workload, seed=rank)
    if args.workload == "c5":
        args.eager = True  # 40k atoms at max_num_neighbors=128: exact-size (dynamic) buffers instead of 2.4x capacity ones
    hp["static_shapes"] = (not args.eager) and hp["model"] == "tensornet"   # capacity-sized buffers + device counters: the frame is CUDA-graph capturable
    model = seeded_model(hp).to(dev)
    for p_ in model.parameters():
        p_.requires_grad_(False)
    zd, bd, boxd = z.to(dev), batch.to(dev), (None if box is None else box.to(dev))
    n_sys = int(batch.max()) + 1
    is_tn = hp["model"] == "tensornet"
    pos_d = pos.to(dev)

    def barrier():
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()

    ff = None
    if not args.eager and is_tn:  # (the ET module reads the pair counter on the host: eager only)
        from torchmdnet_b200.calculators import CUDAGraphForceField
        ff = CUDAGraphForceField(model, zd, bd, box=boxd, num_systems=n_sys).capture(pos_d)

    def frame(p):
        y, f = ff(p) if ff is not None else model(zd, p, bd, box=boxd, num_systems=n_sys)
        if dist_on:  # the single collective of the path: total energy over all boxes
            tot = y.sum()
            dist.all_reduce(tot)
        return y, f

    for _ in range(max(args.warmup, 3)):
        frame(pos_d.clone())
    barrier()
    launches_per_frame = None
    if ff is not None:  # count this library's launches in one frame with an eager evaluation (the graph replays the same nodes)
        l_a = _lib.launches(); model(zd, pos_d.clone(), bd, box=boxd, num_systems=n_sys); launches_per_frame = _lib.launches() - l_a
        barrier()
    l0 = _lib.launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            frame(pos_d.clone())
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launches() - l0 if launches_per_frame is None else launches_per_frame * args.steps
    t = torch.tensor([ms], device=dev)
    if dist_on:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()

    # ---- e2e: host buffers, H2D of the inputs + D2H of energy and forces inside the timed region ----
    pin = lambda x: x.clone().pin_memory()
    h_z, h_pos, h_batch, h_box = pin(z), pin(pos), pin(batch), pin(box if box is not None else torch.zeros(3, 3))
    h_f = torch.empty(pos.shape, dtype=torch.float32).pin_memory()
    h_e = torch.empty(n_sys, 1, dtype=torch.float32).pin_memory()
    h2d = sum(x.numel() * x.element_size() for x in (h_z, h_pos, h_batch, h_box))
    d2h = h_f.numel() * 4 + h_e.numel() * 4

    def frame_e2e():
        zz, pp, bb, xx = (h_z.to(dev, non_blocking=True), h_pos.to(dev, non_blocking=True),
                          h_batch.to(dev, non_blocking=True), h_box.to(dev, non_blocking=True))
        y, f = frame(pp)
        h_e.copy_(y, non_blocking=True); h_f.copy_(f, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller reads the forces before the next step

    for _ in range(3):
        frame_e2e()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        frame_e2e()
    ev1.record()
    barrier()
    t2 = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if dist_on:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    ms_e2e = t2.item()

    # ---- the same frame with the radial tables off: every pair of every frame goes through the tcgen05 chain ----
    exact_arm = None
    from torchmdnet_b200.extensions import fused as _fused
    table_on = bool(_fused.USE_RADIAL_TABLE)
    if table_on and is_tn and args.workload != "c5":
        _fused.USE_RADIAL_TABLE = False
        try:
            ff2 = None
            if ff is not None:
                from torchmdnet_b200.calculators import CUDAGraphForceField
                ff2 = CUDAGraphForceField(model, zd, bd, box=boxd, num_systems=n_sys).capture(pos_d)
            frame2 = (lambda p: ff2(p)) if ff2 is not None else (lambda p: model(zd, p, bd, box=boxd, num_systems=n_sys))
            for _ in range(3):
                frame2(pos_d.clone())
            barrier()
            ev0.record()
            for _ in range(args.steps):
                frame2(pos_d.clone())
            ev1.record()
            barrier()
            t3 = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
            if dist_on:
                dist.all_reduce(t3, op=dist.ReduceOp.MAX)
            exact_arm = {"value": world * args.steps / (t3.item() * 1e-3), "unit": UNIT, "ms_per_step": t3.item() / args.steps,
                         "note": "radial tables off (TMD_RADIAL_TABLE=0): the radial MLP of every pair is evaluated on the tcgen05 chain each frame"}
            del ff2
        finally:
            _fused.USE_RADIAL_TABLE = True

    if rank != 0:
        if dist_on:
            dist.destroy_process_group()
        return

    roofline = roofline2 = roofline3 = None
    N = z.numel()
    if is_tn:
        # ---- roofline of the dominant hand-written kernel: the tcgen05 radial chain of one interaction layer (a6) ----
        from torchmdnet_b200.extensions import tensornet_ops as ops
        from torchmdnet_b200.extensions import fused
        rep = model.representation_model
        g = rep.distance.graph(pos_d, bd, boxd)
        pidx = fused.build_pair_index(g)
        E, Pn = torch.cat([g.num_pairs, pidx.num_unique]).tolist()
        N = z.numel(); F = hp["embedding_dimension"]; R = hp["num_rbf"]
        pair_r = pidx.pair_r[:Pn].contiguous()
        chain = rep.layers[0].radial_chain()
        f16_mode = bool(chain.supported_tc() and chain.f16_safe(rep.distance_expansion, 0.0, hp["cutoff_upper"]))
        # the rbf operand is built once per frame and shared by the chains of all layers (as in the model's forward)
        rbf_op = fused.precompute_rbf(pair_r, pidx.num_unique, rep.distance_expansion, 0.0, hp["cutoff_upper"], True, f16=f16_mode) \
            if chain.supported_tc() else None
        run = lambda: fused.run_chain(chain, pair_r, pidx.num_unique, rep.distance_expansion, 0.0, hp["cutoff_upper"], True, rbf_op, exact=True)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        reps = 10
        ev0.record()
        for _ in range(reps):
            run()
        ev1.record(); torch.cuda.synchronize()
        c_ms = ev0.elapsed_time(ev1) / reps
        hbm_peak, tc_peak, how = measured_peaks()
        # algorithmic flops of one chain call: value + d/dr rows (2 per pair) through R->F->2F->3F (2(RF+2F^2+6F^2)
        # = 270 336 flop per evaluated row at F=128, R=32); the 3xTF32 split executes 3x this on the tensor pipe
        flops = 2 * Pn * 2 * (R * F + F * 2 * F + 2 * F * 3 * F)
        ach_tf = flops / (c_ms * 1e-3) / 1e12
        split_div = 3.0 if f16_mode else 6.0   # 3 split MMAs per product; tf32 additionally runs at half the 16-bit rate
        fused_tail = f16_mode and os.environ.get("TMD_TC_FUSE", "1") != "0" and F in (64, 128)
        kname = ("gemm3x_tf32_kernel (layer 1) + chain_tail_f16_kernel (layers 2+3 fused, h2 in shared memory)" if fused_tail
                 else "gemm3x_tf32_kernel x3")
        roofline3 = {"kernel": kname + ": tcgen05 radial chain of one interaction layer (3 split-term MMAs per product), value + d/dr on unique pairs",
                    "bound": "tensor", "achieved": ach_tf, "peak": tc_peak, "peak_source": how + " dense bf16 burst", "unit": "TFLOP/s",
                    "frac": ach_tf / tc_peak, "frac_of_split_ceiling": ach_tf / (tc_peak / split_div),
                    "operands": "fp16 (hi, lo) split terms" if f16_mode else "tf32 (hi, lo) split terms",
                    "note": f"fp32-parity path: every product needs 3 split MMAs (lo.hi + hi.lo + hi.hi), so the tensor ceiling of this kernel is peak/{split_div:.0f}",
                    "traffic": (1.45e9 if fused_tail else (2.53e9 if f16_mode else 4.240e9)),
                    "traffic_source": "ncu --set full r01, dram read+write summed over the launches of one chain call on the C4 pair count (profiles/r01_chain_tail_ncu.md, profiles/r01_gemm3x_ncu.md)",
                    "algorithmic_flops": flops, "kernel_ms": c_ms, "pairs": Pn}
        roofline3["role"] = ("exact evaluation of the radial MLP: every frame when the tables are off or rejected by their gate; "
                             "with tables on it runs once per model, at table build")
        # ---- dominant kernel of the frame: interaction message-passing backward (a7 + dE/dr), HBM/L2-gather bound ----
        from torchmdnet_b200 import _lib as L
        from torchmdnet_b200._lib import ptr as _p
        Y = [torch.randn(N, k, F, device=dev) for k in (1, 3, 5)]
        dM = [torch.randn(N, k, F, device=dev) for k in (1, 3, 5)]
        dY = [torch.empty(N, k, F, device=dev) for k in (1, 3, 5)]
        f_ = torch.randn(Pn, 3, F, device=dev); fp_ = torch.randn(Pn, 3, F, device=dev)
        gd = torch.zeros(E, device=dev)
        eb = lambda: L.call("tmd_mp_edge_bwd", _p(Y[0]), _p(Y[1]), _p(Y[2]), _p(dM[0]), _p(dM[1]), _p(dM[2]), _p(f_), _p(fp_), _p(pidx.pid),
                            _p(g.nbr_j), _p(g.indptr), N, F, _p(dY[0]), _p(dY[1]), _p(dY[2]), _p(gd), device=dev)
        for _ in range(3):
            eb()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(reps):
            eb()
        ev1.record(); torch.cuda.synchronize()
        b_ms = ev0.elapsed_time(ev1) / reps
        # algorithmic bytes: f and f' once per unique pair, Y_b / dM / dY once per node, CSR (nbr_j, pid) + g_dist per slot
        eb_bytes = 4 * (2 * 3 * Pn * F + 3 * 9 * N * F) + 4 * (3 * E + N + 1)
        eb_ach = eb_bytes / (b_ms * 1e-3) / 1e9
        roofline = {"kernel": "mp_edge_bwd_kernel (interaction message passing backward fused with dE/dr, one layer): dY_b = sum f_p dM_j, g_dist = <f'_p, dM_j * Y_b>",
                    "bound": "hbm", "achieved": eb_ach, "peak": hbm_peak, "peak_source": how, "unit": "GB/s", "frac": eb_ach / hbm_peak,
                    "traffic": 1.53e9, "traffic_source": "ncu --set full r01 (dram read 1.48 GB + write 0.05 GB per launch, profiles/r01_mp_ncu.md): f and f' are fetched once per direction when the two uses of a pair are further apart than L2",
                    "algorithmic_bytes": eb_bytes, "kernel_ms": b_ms, "edges": E, "pairs": Pn,
                    "share_of_frame": "largest single kernel of the frame (3 launches), profiles/r01_launches_final.md"}
        # ---- secondary: interaction message passing forward (a7), HBM/L2-gather bound ----
        I_, A_, S_ = (torch.randn(N, k, F, device=dev) for k in (1, 3, 5))
        w = torch.randn(E, 3, F, device=dev)
        for _ in range(3):
            ops.fn_message_passing(I_, A_, S_, w, None, g.nbr_j, g.indptr, g.rev, g.nbr_j, g.indptr)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(reps):
            ops.fn_message_passing(I_, A_, S_, w, None, g.nbr_j, g.indptr, g.rev, g.nbr_j, g.indptr)
        ev1.record(); torch.cuda.synchronize()
        k_ms = ev0.elapsed_time(ev1) / reps
        alg_bytes = 4 * (3 * E * F + 2 * 9 * N * F) + 4 * (2 * E + N + 1)  # edge weights + node in/out + CSR
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        roofline2 = {"kernel": "mp_fwd_kernel (interaction message passing fwd, one layer)", "bound": "hbm", "achieved": achieved,
                     "peak": hbm_peak, "peak_source": how, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                     "algorithmic_bytes": alg_bytes, "kernel_ms": k_ms, "edges": E}

    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_side = 6 if args.workload == "c4" else 4
        if not (is_tn and box is not None and box.dim() == 2):
            raise SystemExit("cpu baseline sampling is defined for the single-box TensorNet workloads; use --no-cpu-baseline")
        fps, n_sample, tt, cores = cpu_oracle_frames_per_s(hp, n_side, 46.42 * n_side / 15, 2, cores, N)
        cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_sample}-atom sub-box of the same water system, {tt:.2f} s/frame on {cores} threads, scaled by atoms ratio"}

    value = world * args.steps / (ms * 1e-3)
    metric = METRIC if args.workload == "c4" else f"frames_per_sec_energy_forces_{args.workload}"
    out = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": desc, "atoms_per_gpu": N, "edges_per_gpu": int(model.representation_model.distance.graph(pos_d, bd, boxd).num_pairs.item()), "boxes": world, "parallelism": f"dp{world} (one box per GPU, one scalar all-reduce)",
                      "l2": "per-frame working set ~4 GB >> 126 MB L2; no explicit flush",
                      "cuda_graph": ff is not None,
                      "radial_table": table_on and is_tn},
           "clocks": clocks.summary(),
           "e2e": {"value": world * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": ms_e2e / args.steps},
           "gpu_launches": launches, "roofline": roofline, "roofline_secondary": roofline2, "roofline_exact_chain": roofline3,
           "exact_chain": exact_arm, "cpu_baseline": cpu}
    print(json.dumps(out), flush=True)
    if dist_on:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
