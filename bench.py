#!/usr/bin/env python
"""Benchmark of the NeRF-tiny per-ray hot path on B200 (see BASELINE.json / SURVEY.md §8(d)).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--precision fp16|bf16|tc32|fp32]

One "step" = one pass of the hot path over one batch of synthetic input.  Workload at N=1 = BASELINE.json
configs[1]: lego.ini coarse64+fine128 render of a full 400x400 synthetic Blender-shape view (160 000 rays,
random-init weights of the reference architecture, synthetic pose).  With N>1 every rank renders its own view
(rays shard with no data-path collective: weak scaling); value = all ranks' rays / max-over-ranks time.

Printed JSON (one line, rank 0): metric/value/unit (device-resident inputs), ms_per_step, e2e (public API,
host buffers, H2D+D2H inside the timed region), roofline (dominant kernel: the fused tcgen05 encode+MLP),
cpu_baseline (the reference's own CPU path - oracle/_ref, else the oracle port - timed on the host cores), clocks,
gpu_launches, and extra blocks: train / train_4096 (training step with its own roofline), train_cfg4 (BASELINE configs[3]:
one 4096-ray fern-shape batch ray-sharded over the N ranks, strong scaling), render_cfg5 (configs[4]: 1 048 576 rays of
800x800 views split over the N ranks INCLUDING the gather to rank 0 and the copy to the host), precisions (the MLP kernel
of every precision mode at the fine-pass shape).
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
DTYPE = {"fp16": "fp16", "bf16": "bf16", "mixed": "fp16 (coarse pass fp16x3)", "tc32": "fp16x3 (split fp16, fp32-tolerance mode)", "fp32": "f32"}


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
# reference arm / cpu_baseline: the reference's own CPU PyTorch path on the host cores.  The UNMODIFIED reference module
# (staged into oracle/_ref/ at build time, see oracle/ref_harness.py) when present, else the oracle port.
# --------------------------------------------------------------------------------------------------------
def cpu_render_fn(n_rays, threads):
    """-> (callable running one no-grad render of `n_rays` strided pixels of the bench view, kind)."""
    from oracle import nerf_oracle as O
    from oracle import ref_harness as RH
    torch.set_num_threads(threads)
    row, col, pix, pb, k_inv, _ = view_inputs(0)
    sel = torch.linspace(0, N_RAYS - 1, n_rays).long()
    row, col, pb = row[sel].contiguous(), col[sel].contiguous(), pb[sel].contiguous()
    if RH.available():
        try:
            ref = RH.import_reference()
            model = RH.make_model(ref, n_rays, O.init_state_dict(624))
            model.eval()

            def go_ref():
                with torch.no_grad():
                    return model(row, col, pb, k_inv)            # nerf.py:333, unmodified
            go_ref()
            return go_ref, "reference"
        except Exception as e:       # the baseline must never sink the run: fall back to the port, say so
            print(f"[bench] reference module unusable ({e!r}); timing the oracle port", file=sys.stderr)
    sd = O.init_state_dict(624)

    def go_port():
        with torch.no_grad():
            return O.forward(sd, row.numpy(), col.numpy(), pb, k_inv)
    return go_port, "port"


def cpu_render_rays_per_s(n_rays, iters, threads):
    go, kind = cpu_render_fn(n_rays, threads)
    times = []
    for i in range(iters + 1):
        t0 = time.perf_counter()
        go()
        if i > 0:
            times.append(time.perf_counter() - t0)
    return n_rays / float(np.mean(times)), float(np.mean(times)), kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = 2048
    go, kind = cpu_render_fn(n, threads)
    for _ in range(args.warmup):
        go()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        go()
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    what = "the unmodified reference module (oracle/_ref/nerf.py NeRFModel.forward)" if kind == "reference" else \
        "oracle port of the reference (torch CPU fp32)"
    sample = f"{n} rays of the view per step (evenly strided pixels), {what}, {threads} threads"
    out = {"impl": "reference", "metric": "rays/sec (render, coarse64+fine128)", "value": v, "unit": "rays/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "sample": sample},
           "cpu_baseline": {"value": v, "unit": "rays/s", "cores": threads, "kind": kind, "sample": sample},
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
    # The contract is ONE JSON line on stdout.  Native libraries (NCCL's version banner, symmetric-memory setup) write to
    # file descriptor 1 directly, so everything but the final line is routed to stderr at the descriptor level.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
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
    train = train4k = cfg4 = cfg5 = precs = None
    if not args.no_train:
        train = bench_train(model, dev, rows17, world, rank, barrier, args)
        train4k = bench_train(model, dev, rows17, world, rank, barrier, args, batch=4096)
        cfg4 = bench_train_cfg4(model, dev, world, rank, barrier, args)
    if not args.no_extra:
        cfg5 = bench_render_cfg5(model, dev, world, rank, barrier, args)
        if rank == 0:
            precs = bench_precisions(model, dev, d_row, d_col, d_pb, d_kinv, flat, flush)

    cpu = None
    if rank == 0 and not args.no_cpu:
        try:
            threads = os.cpu_count() or 1
            v, sec, kind = cpu_render_rays_per_s(1024, 3, threads)
            cpu = {"value": v, "unit": "rays/s", "cores": threads, "kind": kind,
                   "sample": f"3 x 1024 rays of the same view (strided pixels), "
                             f"{'the unmodified reference module' if kind == 'reference' else 'oracle port of the reference'} "
                             f"(CPU PyTorch fp32), {sec:.2f} s per 1024-ray batch"}
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
            "gpu_launches": launches, "roofline": roof, "roofline_hbm_kernels": roof_hbm, "cpu_baseline": cpu, "train": train, "train_4096": train4k, "train_cfg4": cfg4, "render_cfg5": cfg5,
            "precisions": precs, "clocks": clocks,
            "peaks": pk,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from profiles/ncu_traffic.json (written by tools/ncu_traffic.py from an
    `ncu --set full` capture of this bench's own launch); null when no capture of the current kernel is committed."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if kernel and os.path.exists(path):
        try:
            rec = json.load(open(path)).get(kernel)
            if rec:
                return {"traffic": rec["dram_bytes_read"] + rec["dram_bytes_write"],
                        "traffic_unit": "B per launch (dram__bytes_read.sum + dram__bytes_write.sum, %s)" % rec.get("source", "ncu")}
        except Exception:
            pass
    return {"traffic": None, "traffic_unit": "no ncu capture committed for this kernel"}


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
    return {"kernel": "mlp_tc7_kernel (fused encode + 8x256 MLP, cta_group::2 schedule, fine pass 128 samples/ray)" if prec in (2, 3, 4) else
            ("mlp_tc32_kernel (3-pass split-fp16)" if prec == 1 else "gemm_f32_kernel chain (fp32 accuracy path)"), "bound": "tensor", "achieved": ach, "peak": pk["tf"],
            "unit": "TFLOP/s", "frac": ach / pk["tf"], "frac_of_sustained": ach / pk["tf_sustained"] if pk["tf_sustained"] else None,
            "peak_source": pk["src"] + " bf16 burst", "launch_ms": dur * 1e3, "algorithmic_flop_per_launch": flops,
            # dram__bytes_read.sum + dram__bytes_write.sum per launch, read from the committed ncu --set full summary of this
            # kernel (tools/ncu_traffic.py writes it); algorithmic bytes = 4 B t + 16 B rgb/sigma per sample
            **ncu_traffic("mlp_tc7_kernel" if prec in (2, 3, 4) else ("mlp_tc32_kernel" if prec == 1 else None))}


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


def _max_over_ranks(ms, dev, world):
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return ms


def _time_steps(one, steps, warm, barrier, dev, world):
    for i in range(warm):
        one(i)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        out = one(i)
    b.record()
    barrier()
    return _max_over_ranks(a.elapsed_time(b) / steps, dev, world), out


def bench_train(model, dev, rows17, world, rank, barrier, args, batch=TRAIN_BATCH, steps=None):
    """Weak scaling: every rank trains on its own `batch` rays per step (lego shape), one SUM all-reduce per step."""
    from nerf_tiny_b200 import nerf, synth
    import torch.distributed as dist
    gen = torch.Generator().manual_seed(1000 + rank)
    k_inv = synth.k_inv_of(H, W, synth.focal_of(W))
    batches = [synth.random_batch(rows17, batch, H, W, gen) for _ in range(4)]
    opt = nerf.FusedAdam(model, lr=3e-4, betas=(0.9, 0.999), eps=1e-7)
    saved = model.network.flat_params().clone()
    model.train()
    allreduce = (lambda g: dist.all_reduce(g, op=dist.ReduceOp.SUM)) if world > 1 else None
    # N>1: the SUM all-reduce is fused into the Adam kernel over NVLink peer memory (falls back to NCCL if unavailable)
    fused_ar = opt.enable_peer_allreduce() if world > 1 else False
    steps = steps or max(50, args.steps)
    pinned = [tuple(x.pin_memory() for x in b) for b in batches]
    # forward / loss / backward replayed from one CUDA graph (NT_TRAIN_GRAPH=0: launch by launch through train_step)
    graphed = nerf.GraphedTrainStep(model, opt, batch, k_inv) if os.environ.get("NT_TRAIN_GRAPH", "1") != "0" else None

    def one(i):
        row, col, pix, pb, pic = pinned[i % len(pinned)]
        if graphed is not None:
            loss, _, _ = graphed(row, col, pix, pb, grad_allreduce=allreduce)
        else:
            loss, _, _ = nerf.train_step(model, opt, row, col, pix, pb, k_inv, grad_allreduce=allreduce)
        return loss
    ms, loss = _time_steps(one, steps, 5, barrier, dev, world)
    last_loss = float(loss)                                        # the step's result read back on the host
    model.check_status()
    model.network.flat_params().copy_(saved)
    model.eval()
    v = world * batch / (ms * 1e-3)
    pk = peaks()
    tf = v / world * FLOP_PER_RAY_TRAIN / 1e12
    tr = ncu_traffic_train()
    return {"metric": "rays/sec (train step: fwd+bwd+Adam, coarse64+fine128)", "value": v, "unit": "rays/s",
            "ms_per_step": ms, "rays_per_step_per_gpu": batch, "steps": steps, "warmup": 5, "scaling": "weak",
            "dtype": "bf16" if args.precision in ("bf16", "fp16") else "f32",
            "note": "fused tcgen05 forward with TMA-stored bf16 stash + fused tcgen05 backward-data chain + one grouped split-K dW launch per pass "
                    "+ fused Adam; forward/loss/backward replayed as one CUDA graph; host batches in pinned memory, H2D inside the timed "
                    "region, the loss is read back on the host after the last step; one SUM all-reduce of the 2.4 MB gradient per step when N>1",
            "e2e": {"h2d_bytes_per_step": batch * (8 + 8 + 12 + 17 * 8), "d2h_bytes_per_step": 4,
                    "note": "value IS end to end: the public GraphedTrainStep / train_step call with host batches"},
            "grad_exchange": ("fused all-reduce+Adam kernel over NVLink peer memory" if fused_ar else
                              ("NCCL all-reduce" if world > 1 else "none")),
            "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["tf"], "unit": "TFLOP/s", "frac": tf / pk["tf"],
                         "frac_of_sustained": tf / pk["tf_sustained"] if pk["tf_sustained"] else None,
                         "algorithmic_flop_per_ray": FLOP_PER_RAY_TRAIN, "peak_source": pk["src"] + " bf16 burst", **tr},
            "frac_of_tc_peak": tf / pk["tf"], "last_loss": last_loss}


def ncu_traffic_train():
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        try:
            rec = json.load(open(path)).get("train_step")
            if rec:
                return {"traffic": rec["dram_bytes_per_sample"], "traffic_unit": "DRAM B per sample, all kernels of one step (%s)" % rec.get("source", "ncu")}
        except Exception:
            pass
    return {"traffic": None, "traffic_unit": "no ncu capture committed"}


FERN_H, FERN_W, FERN_F, CFG4_BATCH = 378, 504, 407.6, 4096


def bench_train_cfg4(model, dev, world, rank, barrier, args):
    """BASELINE configs[3]: fern.ini LLFF-shape 504x378 rays with per-image near/far, ONE 4096-ray batch per step
    ray-sharded over the N ranks (strong scaling): every rank trains on its contiguous 4096/N slice with the batch-global
    quantities reduced on the device, gradients SUM-all-reduced (fused into the Adam kernel over NVLink peer memory)."""
    from nerf_tiny_b200 import nerf, synth, dist as D
    import torch.distributed as dist
    rows17 = synth.pose_rows(20, FERN_H, FERN_W, FERN_F, llff_bounds=True, seed=3)
    k_inv = synth.k_inv_of(FERN_H, FERN_W, FERN_F)
    gen = torch.Generator().manual_seed(4040)                     # the SAME global batches on every rank
    batches = [synth.random_batch(rows17, CFG4_BATCH, FERN_H, FERN_W, gen) for _ in range(4)]
    sl = D.shard_slice(CFG4_BATCH, rank, world)
    n_loc = sl.stop - sl.start
    pinned = [tuple(x[sl].contiguous().pin_memory() for x in b) for b in batches]
    opt = nerf.FusedAdam(model, lr=3e-4, betas=(0.9, 0.999), eps=1e-7)
    saved = model.network.flat_params().clone()
    model.train()
    allreduce = (lambda g: dist.all_reduce(g, op=dist.ReduceOp.SUM)) if world > 1 else None
    fused_ar = opt.enable_peer_allreduce() if world > 1 else False
    shard = (rank, world) if world > 1 else None
    graphed = nerf.GraphedTrainStep(model, opt, n_loc, k_inv, shard=shard)

    def one(i):
        row, col, pix, pb, pic = pinned[i % len(pinned)]
        return graphed(row, col, pix, pb, grad_allreduce=allreduce)[0]
    steps = max(50, args.steps)
    ms, loss = _time_steps(one, steps, 5, barrier, dev, world)
    last_loss = float(loss)
    # the exchange's share: the optimiser step alone (N>1: two symmetric-memory barriers + the kernel that sums all ranks'
    # gradients over NVLink and applies Adam; N=1: the Adam kernel), and the 16-byte MAX all-reduce of the shard globals
    ms_opt, _ = _time_steps(lambda i: opt.step(), 50, 5, barrier, dev, world)
    ms_glob = None
    if world > 1:
        near, far = graphed.near, graphed.far
        ms_glob, _ = _time_steps(lambda i: model._globals(shard, near, far), 50, 5, barrier, dev, world)
    model.check_status()
    model.network.flat_params().copy_(saved)
    model.eval()
    v = CFG4_BATCH / (ms * 1e-3)
    pk = peaks()
    tf = v * FLOP_PER_RAY_TRAIN / 1e12 / world
    return {"metric": "rays/sec (train step, fern 504x378, one 4096-ray batch ray-sharded over N GPUs)", "value": v,
            "unit": "rays/s", "ms_per_step": ms, "global_batch": CFG4_BATCH, "rays_per_step_per_gpu": n_loc, "steps": steps,
            "scaling": "strong", "n_gpus": world, "dtype": "bf16",
            "optimizer_and_exchange_ms": ms_opt, "shard_globals_allreduce_ms": ms_glob,
            "exchange_share_of_step": (ms_opt + (ms_glob or 0.0)) / ms,
            "grad_exchange": ("fused all-reduce+Adam kernel over NVLink peer memory" if fused_ar else
                              ("NCCL all-reduce" if world > 1 else "none")),
            "frac_of_tc_peak_per_gpu": tf / pk["tf"], "last_loss": last_loss}


CFG5_H = CFG5_W = 800
CFG5_RAYS = 1 << 20


def bench_render_cfg5(model, dev, world, rank, barrier, args):
    """BASELINE configs[4]: 800x800 views, 1 048 576 rays per launch (1.64 views) ray-sharded over the N ranks; the timed
    step INCLUDES the gather of the (N,3) outputs to every rank (NCCL all-gather over NVLink, 12 B/ray) and rank 0's copy
    of the assembled image block to pinned host memory."""
    from nerf_tiny_b200 import synth, dist as D
    f = synth.focal_of(CFG5_W)
    rows17 = synth.pose_rows(8, CFG5_H, CFG5_W, f)
    k_inv = synth.k_inv_of(CFG5_H, CFG5_W, f)
    idx = np.arange(CFG5_RAYS)
    pic, rem = idx // (CFG5_H * CFG5_W), idx % (CFG5_H * CFG5_W)
    sl = D.shard_slice(CFG5_RAYS, rank, world)
    row = torch.from_numpy((rem // CFG5_W)[sl].astype(np.int64)).to(dev)
    col = torch.from_numpy((rem % CFG5_W)[sl].astype(np.int64)).to(dev)
    pb = torch.from_numpy(rows17[pic[sl]]).to(dev).float().contiguous()
    kinv = k_inv.to(dev)
    near, far = pb[:, 15].contiguous(), pb[:, 16].contiguous()
    flat = model.network.flat_params()
    host = torch.empty(CFG5_RAYS, 3, dtype=torch.float32).pin_memory() if rank == 0 else None
    shard = (rank, world) if world > 1 else None

    def one(i):
        with torch.no_grad():
            asz, d0 = model._globals(shard, near, far)
            cc, cf, _ = model._render_raw(flat, row, col, pb, kinv, near, far, train=False, any_step_zero=asz, delta0=d0)
            full = D.gather_rows(cf, CFG5_RAYS)
            if rank == 0:
                host.copy_(full, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return cf
    steps = max(3, min(args.steps, 5))
    ms, _ = _time_steps(one, steps, 2, barrier, dev, world)
    model.check_status()
    v = CFG5_RAYS / (ms * 1e-3)
    pk = peaks()
    return {"metric": "rays/sec (render 800x800, 1 048 576 rays per launch ray-sharded over N GPUs, gather + D2H included)",
            "value": v, "unit": "rays/s", "ms_per_step": ms, "rays_per_launch": CFG5_RAYS, "rays_per_gpu": sl.stop - sl.start,
            "steps": steps, "scaling": "strong", "n_gpus": world, "dtype": DTYPE[args.precision],
            "d2h_bytes_per_step": CFG5_RAYS * 12, "gather_bytes_per_rank": (CFG5_RAYS * 12) if world > 1 else 0,
            "frac_of_tc_peak_per_gpu": v / world * FLOP_PER_RAY_RENDER / (pk["tf"] * 1e12)}


def bench_precisions(model, dev, d_row, d_col, d_pb, d_kinv, flat, flush):
    """The encode+MLP kernel of every precision mode alone at the fine-pass shape of the workload (160 000 x 128 samples):
    the fp32-tolerance mode on the tensor cores (tc32, 3 MMAs per product: its roofline is a third of the 16-bit peak) next to
    the CUDA-core fp32 path it replaces, and bf16 next to fp16 (same kernel, same speed)."""
    from nerf_tiny_b200 import _lib, nerf
    import ctypes as C
    L, h = model._lib, model._ctx
    n = d_row.shape[0]
    rays, denc = torch.empty(n, 16, device=dev), torch.empty(n, 24, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.nt_raygen(h, n, p(d_row), p(d_col), p(d_pb), 17, p(d_kinv), p(rays), None, p(denc), st))
    t = (torch.rand(n, 128, device=dev, generator=torch.Generator(device=dev).manual_seed(3)) * 4 + 2).contiguous()
    rgb, sig = torch.empty(n, 128, 3, device=dev), torch.empty(n, 128, device=dev)
    pk = peaks()
    out, ref = [], None
    for name in ("fp32", "tc32", "fp16", "bf16"):
        prec = nerf.PRECISION[name]
        packed = model._pack(flat, prec)
        ws = torch.empty(max(256, L.nt_mlp_workspace_bytes(h, prec, n, 128, 0)), dtype=torch.uint8, device=dev)
        go = lambda: _lib.check(L.nt_mlp_forward(h, prec, n, 128, p(t), p(rays), p(denc), p(flat), p(packed), p(rgb), p(sig),
                                                 p(ws), ws.numel(), 0, st))
        iters = 1 if name == "fp32" else 3
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
        if name == "fp32":
            ref = (rgb.clone(), sig.clone())
        err_rgb = float((rgb - ref[0]).abs().max())
        err_sig = float((sig - ref[1]).abs().max() / ref[1].abs().max())
        ach = FLOP_PER_SAMPLE * n * 128 / dur / 1e12
        mma_per_product = {"fp32": None, "tc32": 3, "fp16": 1, "bf16": 1}[name]
        peak = pk["tf"] / mma_per_product if mma_per_product else None
        out.append({"precision": name, "kernel": {"fp32": "gemm_f32_kernel chain (CUDA cores)", "tc32": "mlp_tc32_kernel (tcgen05, 3-pass split fp16)",
                                                  "fp16": "mlp_tc7_kernel<fp16> (tcgen05 cta_group::2)", "bf16": "mlp_tc7_kernel<bf16> (tcgen05 cta_group::2)"}[name],
                    "launch_ms": dur * 1e3, "samples_per_s": n * 128 / dur, "useful_tflops": ach,
                    "roofline_peak_tflops": peak, "frac": ach / peak if peak else None,
                    "max_abs_rgb_vs_fp32": err_rgb, "max_rel_sigma_vs_fp32": err_sig})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "mixed", "tc32", "fp32"])
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg5 render sweep and the precision table")
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
