#!/usr/bin/env python
"""Benchmark of the NeRF-tiny per-ray hot path on B200 (see BASELINE.json / SURVEY.md §8(d)).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--precision bf16|fp32]

One "step" = one pass of the hot path over one batch of synthetic input.  Workload at N=1 = BASELINE.json
configs[1]: lego.ini coarse64+fine128 render of a full 400x400 synthetic Blender-shape view (160 000 rays,
random-init weights of the reference architecture, synthetic pose).  With N>1 every rank renders its own view
(rays shard with no data-path collective: weak scaling); value = all ranks' rays / max-over-ranks time.

Printed JSON (one line, rank 0): metric/value/unit (device-resident inputs), ms_per_step, e2e (public API,
host buffers, H2D+D2H inside the timed region), roofline (dominant kernel: the fused tcgen05 encode+MLP),
cpu_baseline (the oracle port of the reference timed on the host cores), train (the training step, extra),
clocks, gpu_launches.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 400
N_RAYS = H * W
FLOP_PER_SAMPLE = 1182976            # 591 488 MAC
SAMPLES_PER_RAY = 192
FLOP_PER_RAY_RENDER = FLOP_PER_SAMPLE * SAMPLES_PER_RAY          # 227.131 MFLOP
FLOP_PER_RAY_TRAIN = 676.282e6
WORKLOAD = "lego.ini coarse64+fine128 render of a 400x400 synthetic Blender-shape view (160000 rays)"
TRAIN_BATCH = 1024
# arithmetic type of the MLP contraction per --precision (accumulation is fp32 everywhere)
DTYPE = {"fp16": "fp16", "bf16": "bf16", "tc32": "fp16x3 (split fp16, fp32-tolerance mode)", "fp32": "f32"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained"), src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def view_inputs(rank, height=H, width=W):
    from nerf_tiny_b200 import synth
    f = synth.focal_of(width)
    rows17 = synth.pose_rows(8, height, width, f)
    row, col, pix, pb, pic = synth.view_batch(rows17, rank % 8, height, width)
    return row, col, pix, pb, synth.k_inv_of(height, width, f), rows17


# --------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's CPU PyTorch path on the host cores
# --------------------------------------------------------------------------------------------------------
def cpu_render_rays_per_s(n_rays, iters, threads):
    from oracle import nerf_oracle as O
    torch.set_num_threads(threads)
    row, col, pix, pb, k_inv, _ = view_inputs(0)
    sd = O.init_state_dict(624)
    sel = torch.linspace(0, N_RAYS - 1, n_rays).long()
    row, col, pb = row[sel], col[sel], pb[sel]
    times = []
    with torch.no_grad():
        for i in range(iters + 1):
            t0 = time.perf_counter()
            O.forward(sd, row.numpy(), col.numpy(), pb, k_inv)
            if i > 0:
                times.append(time.perf_counter() - t0)
    return n_rays / float(np.mean(times)), float(np.mean(times))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = 2048
    warm = max(0, min(args.warmup, 1))
    from oracle import nerf_oracle as O
    torch.set_num_threads(threads)
    row, col, pix, pb, k_inv, _ = view_inputs(0)
    sd = O.init_state_dict(624)
    sel = torch.linspace(0, N_RAYS - 1, n).long()
    row, col, pb = row[sel], col[sel], pb[sel]
    with torch.no_grad():
        for _ in range(warm):
            O.forward(sd, row.numpy(), col.numpy(), pb, k_inv)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            O.forward(sd, row.numpy(), col.numpy(), pb, k_inv)
        dt = time.perf_counter() - t0
    v = n * args.steps / dt
    sample = f"{n} rays of the view per step (evenly strided pixels), oracle port (torch CPU fp32), {threads} threads"
    out = {"impl": "reference", "metric": "rays/sec (render, coarse64+fine128)", "value": v, "unit": "rays/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": warm, "ms_per_step": 1e3 * dt / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "sample": sample},
           "cpu_baseline": {"value": v, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# --------------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on STDOUT when NCCL_DEBUG>=VERSION; the contract is one JSON line there
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    from nerf_tiny_b200 import nerf, ops, _lib, build
    build.build()
    nerf.seed_everything(624)
    model = nerf.NeRFModel(64, 128, batch_ray=N_RAYS, precision=args.precision).to(dev)
    model.check_range = False          # status flag is read once after the timed region (no per-step host sync)
    model.eval()
    row, col, pix, pb, k_inv, rows17 = view_inputs(rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident inputs: C-ABI nt_render_forward on tensors already in HBM ------------------
    model._ensure_ctx()
    d_row, d_col = row.to(dev), col.to(dev)
    d_pb = pb.to(dev).float().contiguous()
    d_kinv = k_inv.to(dev)
    d_near, d_far = d_pb[:, 15].contiguous(), d_pb[:, 16].contiguous()
    flat = model.network.flat_params()

    def step_dev():
        with torch.no_grad():
            return model._render_raw(flat, d_row, d_col, d_pb, d_kinv, d_near, d_far, train=False)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        l0 = model.launch_count
        for a, b in ev:
            flush.zero_()                 # L2 flush between timed iterations (outside the events)
            a.record()
            fn()
            b.record()
        barrier()
        launches = model.launch_count - l0
        ms = sum(a.elapsed_time(b) for a, b in ev)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, launches

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, launches = timed(step_dev, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    model.check_status()

    # ---- end to end: public API with host (pinned) buffers, H2D + D2H inside the timed region -----------
    h_row, h_col, h_pb = row.pin_memory(), col.pin_memory(), pb.pin_memory()
    out_host = torch.empty(2, N_RAYS, 3, dtype=torch.float32).pin_memory()

    def step_e2e():
        with torch.no_grad():
            cc, cf = model(h_row, h_col, h_pb, k_inv)          # the call a user of the reference makes (nerf.py:516)
            out_host[0].copy_(cc, non_blocking=True)
            out_host[1].copy_(cf, non_blocking=True)
        torch.cuda.current_stream().synchronize()               # result is on the host

    ms_e2e, _ = timed(step_e2e, args.steps, args.warmup)
    h2d = h_row.numel() * 8 + h_col.numel() * 8 + h_pb.numel() * 8 + k_inv.numel() * 4
    d2h = out_host.numel() * 4

    # ---- dominant kernel alone: fused encode+MLP (fine pass, 128 samples/ray), CUDA events on its stream --
    roof, roof_hbm = None, None
    if rank == 0:
        roof = mlp_roofline(model, dev, d_row, d_col, d_pb, d_kinv, flat, flush)
        roof_hbm = hbm_rooflines(model, dev, flush, N_RAYS)

    # ---- training step (extra): reference loop body nerf.py:464-475 through train_step ------------------
    train = None
    if not args.no_train:
        train = bench_train(model, dev, rows17, world, rank, barrier, args)

    cpu = None
    if rank == 0 and not args.no_cpu:
        try:
            threads = os.cpu_count() or 1
            v, sec = cpu_render_rays_per_s(1024, 3, threads)
            cpu = {"value": v, "unit": "rays/s", "cores": threads, "kind": "port",
                   "sample": f"3 x 1024 rays of the same view (strided pixels), oracle port of the reference's CPU PyTorch "
                             f"path, {sec:.2f} s per 1024-ray batch"}
        except Exception as e:  # the baseline must never sink the GPU numbers
            cpu = {"value": None, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        pk = peaks()
        value = world * N_RAYS / (ms_dev * 1e-3)
        out = {
            "metric": "rays/sec (render, coarse64+fine128)", "value": value, "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": DTYPE[args.precision],
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_step_per_gpu": N_RAYS, "samples_per_ray": "64+128",
                       "weights": "random init (nn.Linear default), 593924 params", "precision": args.precision,
                       "l2": "flushed between timed steps (256 MiB memset outside the CUDA events)",
                       "parallelism": f"ray-sharded x{world}, no data-path collective"},
            "mlp_tc_frac_of_peak": value / world * FLOP_PER_RAY_RENDER / (pk["tf"] * 1e12),
            "e2e": {"value": world * N_RAYS / (ms_e2e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "roofline": roof, "roofline_hbm_kernels": roof_hbm, "cpu_baseline": cpu, "train": train, "clocks": clocks,
            "peaks": pk,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def mlp_roofline(model, dev, d_row, d_col, d_pb, d_kinv, flat, flush, iters=5):
    """Average duration of the fused encode+MLP kernel over the fine-pass shape of the workload."""
    from nerf_tiny_b200 import ops, _lib
    import ctypes as C
    L, h = model._lib, model._ctx
    n = d_row.shape[0]
    rays = torch.empty(n, 16, device=dev)
    denc = torch.empty(n, 24, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.nt_raygen(h, n, p(d_row), p(d_col), p(d_pb), 17, p(d_kinv), p(rays), None, p(denc), st))
    t = (torch.rand(n, 128, device=dev) * 4 + 2).contiguous()
    rgb = torch.empty(n, 128, 3, device=dev)
    sig = torch.empty(n, 128, device=dev)
    prec = model._prec
    need = max(256, L.nt_mlp_workspace_bytes(h, prec, n, 128, 0))
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    packed = model._pack(flat, prec)

    def go():
        _lib.check(L.nt_mlp_forward(h, prec, n, 128, p(t), p(rays), p(denc), p(flat), p(packed) if packed is not None else None, p(rgb), p(sig),
                                    p(ws), ws.numel(), 0, st))
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        go()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    dur = float(np.mean(ms)) * 1e-3
    flops = FLOP_PER_SAMPLE * n * 128
    pk = peaks()
    ach = flops / dur / 1e12
    return {"kernel": "mlp_tc7_kernel (fused encode + 8x256 MLP, cta_group::2 schedule, fine pass 128 samples/ray)" if prec in (2, 3) else
            ("mlp_tc32_kernel (3-pass split-fp16)" if prec == 1 else "gemm_f32_kernel chain (fp32 accuracy path)"), "bound": "tensor", "achieved": ach, "peak": pk["tf"],
            "unit": "TFLOP/s", "frac": ach / pk["tf"], "frac_of_sustained": ach / pk["tf_sustained"] if pk["tf_sustained"] else None,
            "peak_source": pk["src"] + " bf16 burst", "launch_ms": dur * 1e3, "algorithmic_flop_per_launch": flops,
            # dram__bytes_read.sum + dram__bytes_write.sum of this launch from the ncu --set full capture summarised in
            # profiles/r1j_mlp_tc7_summary.txt (110.8 MB + 277.2 MB); algorithmic bytes = 4 B t + 16 B rgb/sigma per sample
            "traffic": 388.0e6 if prec in (2, 3) else None, "traffic_unit": "B per launch (ncu, profiles/r1j_mlp_tc7_summary.txt)"}


def hbm_rooflines(model, dev, flush, n, iters=5):
    """The per-ray HBM-bound kernels of the render step alone (CUDA events on their stream, L2 flushed before each): achieved
    GB/s = SURVEY.md §8(d)'s algorithmic bytes per ray x rays / launch time, against the measured HBM peak."""
    from nerf_tiny_b200 import _lib
    import ctypes as C
    L, h = model._lib, model._ctx
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device=dev).manual_seed(7)
    near, far = torch.full((n,), 2.0, device=dev), torch.full((n,), 6.0, device=dev)
    t_c = torch.empty(n, 64, device=dev)
    _lib.check(L.nt_sample_coarse(h, n, p(near), p(far), -1, p(t_c), st))
    rgb_c, sig_c = torch.rand(n, 64, 3, device=dev, generator=g), torch.rand(n, 64, device=dev, generator=g) * 3
    rgb_f, sig_f = torch.rand(n, 128, 3, device=dev, generator=g), torch.rand(n, 128, device=dev, generator=g) * 3
    w_c, c_c, c_f = torch.empty(n, 64, device=dev), torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev)
    t_f = torch.empty(n, 128, device=dev)
    calls = [
        ("composite_coarse_kernel (get_density + color_cum, 64 samples)", 1300,
         lambda: L.nt_composite_coarse(h, n, p(near), p(far), p(rgb_c), p(sig_c), p(w_c), p(c_c), st)),
        ("sample_pdf_kernel (cdf, searchsorted, interpolation)", 776,
         lambda: L.nt_sample_pdf(h, n, p(t_c), p(w_c), None, p(t_f), None, st)),
        ("composite_fine_fwd_kernel (merge, 5 channel sorts, compositing, 192 samples)", 3852,
         lambda: L.nt_composite_fine(h, n, p(t_c), p(rgb_c), p(sig_c), p(t_f), p(rgb_f), p(sig_f), 1e-4, p(c_f), None, None, st)),
    ]
    pk = peaks()
    out = []
    for name, bytes_per_ray, fn in calls:
        for _ in range(2):
            _lib.check(fn())
        ms = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(fn())
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        dur = float(np.mean(ms)) * 1e-3
        ach = n * bytes_per_ray / dur / 1e9
        out.append({"kernel": name, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                    "launch_us": dur * 1e6, "algorithmic_bytes_per_ray": bytes_per_ray})
    return out


def bench_train(model, dev, rows17, world, rank, barrier, args):
    from nerf_tiny_b200 import nerf, synth
    import torch.distributed as dist
    gen = torch.Generator().manual_seed(1000 + rank)
    k_inv = synth.k_inv_of(H, W, synth.focal_of(W))
    batches = [synth.random_batch(rows17, TRAIN_BATCH, H, W, gen) for _ in range(4)]
    opt = nerf.FusedAdam(model, lr=3e-4, betas=(0.9, 0.999), eps=1e-7)
    saved = model.network.flat_params().clone()
    model.train()
    allreduce = (lambda g: dist.all_reduce(g, op=dist.ReduceOp.SUM)) if world > 1 else None
    # N>1: the SUM all-reduce is fused into the Adam kernel over NVLink peer memory (falls back to NCCL if unavailable)
    fused_ar = opt.enable_peer_allreduce() if world > 1 else False
    steps, warm = max(2, min(args.steps, 5)), 2
    pinned = [tuple(x.pin_memory() for x in b) for b in batches]

    # forward / loss / backward replayed from one CUDA graph (NT_TRAIN_GRAPH=0: launch by launch through train_step)
    graphed = nerf.GraphedTrainStep(model, opt, TRAIN_BATCH, k_inv) if os.environ.get("NT_TRAIN_GRAPH", "1") != "0" else None

    def one(i):
        row, col, pix, pb, pic = pinned[i % len(pinned)]
        if graphed is not None:
            loss, _, _ = graphed(row, col, pix, pb, grad_allreduce=allreduce)
        else:
            loss, _, _ = nerf.train_step(model, opt, row, col, pix, pb, k_inv, grad_allreduce=allreduce)
        return loss
    for i in range(warm):
        one(i)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        loss = one(i)
    b.record()
    barrier()
    ms = a.elapsed_time(b) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    model.check_status()
    model.network.flat_params().copy_(saved)
    model.eval()
    v = world * TRAIN_BATCH / (ms * 1e-3)
    return {"metric": "rays/sec (train step: fwd+bwd+Adam, coarse64+fine128)", "value": v, "unit": "rays/s",
            "ms_per_step": ms, "rays_per_step_per_gpu": TRAIN_BATCH, "steps": steps,
            "dtype": "bf16" if args.precision in ("bf16", "fp16") else "f32",
            "note": "fused tcgen05 forward with TMA-stored bf16 stash + fused tcgen05 backward-data chain + one grouped split-K dW launch per pass "
                    "+ fused Adam; forward/loss/backward replayed as one CUDA graph; host batches, H2D inside the timed region; one SUM all-reduce of the 2.4 MB gradient "
                    "per step when N>1",
            "grad_exchange": ("fused all-reduce+Adam kernel over NVLink peer memory" if fused_ar else
                              ("NCCL all-reduce" if world > 1 else "none")),
            "frac_of_tc_peak": v / world * FLOP_PER_RAY_TRAIN / (peaks()["tf"] * 1e12), "last_loss": float(loss)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "tc32", "fp32"])
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    run_b200(args)


if __name__ == "__main__":
    main()
