#!/usr/bin/env python
"""bench.py -- headline benchmark of the rasterizer hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--config cfg2]

Metric: Gaussians/s rasterised (forward + fused L1/D-SSIM loss + backward) at 1920x1080, SH degree 3
(BASELINE.json configs[1]: 500 k Gaussians, one view per GPU).  One JSON line on stdout (rank 0).

  value     : whole-job Gaussians/s with every input already resident in HBM (C ABI, persistent workspaces)
  e2e       : the same step through the reference-facing LibTorch symbols (RasterizeGaussiansCUDA ...,
              driven by the autograd mirror of rasterizer.cpp) with the per-iteration HOST inputs of the
              reference loop (gaussian.cpp:674-699): pinned ground-truth image + camera H2D every step,
              loss scalar D2H every step
  roofline  : dominant kernel (per-stage cudaEvent timing inside the library, on the launching stream)
  cpu_baseline : the CPU oracle (port of the reference algorithm; the reference has no CPU path) on the
              host cores, bounded sample
  --impl reference : the reference's own CUDA sources compiled for sm_100a (oracle/_ref), same scene,
              same metric; falls back to the CPU oracle port when that build is absent.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Gaussians/sec rasterized (fwd+bwd) @1920x1080"
LAMBDA_DSSIM = 0.2


def load_synthetic():
    """gaussian_lic_b200/synthetic.py (pure numpy) loaded BY PATH: importing the package would dlopen libglic_b200.so,
    and the reference arm must not have the product library mapped at all."""
    import importlib.util
    name = "_glic_synthetic"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "gaussian_lic_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(P, V, R, HW, M):
    """compulsory HBM bytes of one view, forward + loss + backward (A1)."""
    S = 44 + 12 * (M + 1)
    T = 8160 if HW == 1920 * 1080 else None
    bit = 13 if T == 8160 else 12
    passes = (32 + bit + 7) // 8
    A1 = 44 * P + (12 * (M + 1) + 112 + 2 * S) * V + (100 + 24 * passes) * R + 84 * HW
    per_stage = {
        "preprocess": 44 * P + 12 * (M + 1) * V + 40 * V,
        "emit": 40 * V + 12 * R,
        "sort": (8 + 24 * passes) * R,
        "render_fwd": 40 * R + 16 * HW,
        "loss_fwd": 24 * HW, "loss_bwd": 12 * HW,
        "render_bwd": 40 * R + 32 * HW + 36 * V,
        "preprocess_bwd": 36 * V + S * V + S * V,
    }
    return A1, per_stage


def physical_cores():
    """Physical cores available to this process (SMT siblings slow the OpenMP oracle down, measured)."""
    try:
        allowed = os.sched_getaffinity(0)
        cores, cur = set(), {}
        for line in open("/proc/cpuinfo"):
            if ":" not in line:
                if "processor" in cur and int(cur["processor"]) in allowed:
                    cores.add((cur.get("physical id", "0"), cur.get("core id", cur["processor"])))
                cur = {}
                continue
            k, v = line.split(":", 1)
            cur[k.strip()] = v.strip()
        return max(1, len(cores))
    except Exception:
        return max(1, (os.cpu_count() or 2) // 2)


def cpu_baseline(cfg, sample_P, threads_note=True):
    """CPU oracle (port) timed on the host cores on a bounded sample: same recipe, fewer Gaussians."""
    from oracle.oracle import Oracle
    syn = load_synthetic()
    o = Oracle(np.float32)
    o.set_threads(physical_cores())                        # torchrun exports OMP_NUM_THREADS=1; use the host cores we have
    g, cam = syn.make_scene(cfg, P=sample_P)
    gt = syn.make_gt_image(cam["W"], cam["H"])
    t0 = time.perf_counter()
    f = o.forward(g, cam)
    L, dl = o.loss(f["color"], gt, LAMBDA_DSSIM)
    o.backward(f, dl)
    dt = time.perf_counter() - t0
    o.free(f)
    return {"value": sample_P / dt, "unit": "Gaussians/s", "cores": o.max_threads(), "kind": "port",
            "sample": "%d Gaussians of the %s recipe at full %dx%d, 1 step fwd+loss+bwd (%.1f s), OpenMP oracle "
                      "(gcc -O2 -fopenmp -ffp-contract=off)" % (sample_P, cfg, cam["W"], cam["H"], dt)}, dt


def dist_setup(n_gpus):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def max_over_ranks(ms, world):
    if world == 1:
        return ms
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_ranks(ms, world):
    if world == 1:
        return [ms]
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(x.item()) for x in out]


def timed_windows(run_step, steps, world, dev, min_windows=5, min_total_ms=500.0, max_windows=60):
    """Device-timed windows of EXACTLY `steps` steps each.  Every window is bracketed by a barrier +
    torch.cuda.synchronize() on both sides and timed with a CUDA event pair on the launching stream; a window's time is
    the MAX over ranks.  The reported figure is the MEDIAN window (>= 5 windows and >= 0.5 s of timed work in total),
    so one host-side hiccup on one rank (a 100 ms stall inside a 30 ms window was the round-1 N = 8 collapse) cannot
    set the number; every window is returned so the spread is visible.  -> (median ms/step, [window ms/step], per-rank
    ms/step of the median window)."""
    import torch
    if world > 1:
        import torch.distributed as dist
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wins, ranks = [], []
    n = min_windows
    i = 0
    while i < n:
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(steps):
            run_step()
        e1.record()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        mine = e0.elapsed_time(e1) / steps
        wins.append(max_over_ranks(mine, world))
        ranks.append(mine)
        if i == 0:                                           # identical on every rank (max-reduced): same window count
            n = int(min(max_windows, max(min_windows, -(-min_total_ms // max(wins[0] * steps, 1e-3)))))
        i += 1
    order = sorted(range(len(wins)), key=lambda k: wins[k])
    mid = order[len(order) // 2]
    return wins[mid], [round(w, 4) for w in wins], [round(x, 4) for x in gather_ranks(ranks[mid], world)]


def run_ours(args):
    import torch
    from gaussian_lic_b200 import capi, ops, synthetic as syn
    rank, world, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    cfg = args.config
    P, W, H, fx, fy, cx, cy, deg, zmax = syn.CONFIGS[cfg]
    g, _ = syn.make_scene(cfg)
    # one training view per rank: an 8-view ring of nearby keyframes around view 0, as a sliding-window
    # mapper would batch them; the small radius keeps the per-GPU work of the weak-scaling run within a few percent
    view_id = (rank % 8) if args.view is None else args.view
    R_wc, t_wc = syn.orbit_pose(view_id, radius=args.rig_radius)
    cam = syn.make_camera(W, H, fx, fy, cx, cy, R_wc, t_wc)
    gt_host = torch.as_tensor(syn.make_gt_image(W, H)).pin_memory()
    gd = ops.scene_to_device(g, dev)
    gt = gt_host.to(dev)
    M = gd["sh"].shape[1]
    r = ops.CRasterizer(W, H, dev)
    view = r.make_view(cam)
    f32 = dict(dtype=torch.float32, device=dev)
    color, T = torch.empty(3, H, W, **f32), torch.empty(H, W, **f32)
    radii = torch.empty(P, dtype=torch.int32, device=dev)
    loss_out, dL = torch.empty(1, **f32), torch.empty(3, H, W, **f32)
    grads = r.alloc_grads(P, M)
    allreduce = None
    if world > 1:
        import torch.distributed as dist
        from gaussian_lic_b200 import dist as gdist
        # the exchange step: our own two-shot all-reduce over NVLink peer memory (csrc/p2p.cu); --exchange nccl keeps
        # the library collective as the comparison
        if args.exchange == "p2p":
            try:
                allreduce = gdist.P2PGradAllReduce(P, M, dev)
            except RuntimeError as e:                          # raised on every rank together (dist.py): IPC not permitted here
                print("[bench] %s -- falling back to the NCCL all-reduce" % (e,), file=sys.stderr)
                args.exchange = "nccl"
        if allreduce is None:
            allreduce = gdist.GradAllReduce(P, M, dev)
        grads = allreduce.grads                           # backward writes straight into the collective's buffer

    def step():
        r.forward(gd, view, out_color=color, out_T=T, radii=radii, sync=False)
        r.loss(color, gt, LAMBDA_DSSIM, loss_out, dL)
        r.backward(gd, view, radii, dL, grads)
        if allreduce is not None:
            allreduce(radii)

    r.forward(gd, view, out_color=color, out_T=T, radii=radii, sync=True)      # settles the binning capacity once
    for _ in range(max(args.warmup, 3)):
        step()
    assert not r.finish(), "binning capacity overflow during warm-up"
    V = int((radii > 0).sum().item())
    Rn, Bn = r.R, r.B

    # ---- timed region: K steps, barrier + sync on both sides, CUDA events, max over ranks -------------
    # CUDA-graph the step when it has no collective: forward, loss and backward contain no host synchronisation
    # (capacity-sized binning), so ~25 launches + memsets replay as one graph launch.
    graph = None
    if (allreduce is None or args.exchange == "p2p") and not args.no_graph:
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                r.stream = side.cuda_stream
                step()                                       # warm the side stream
                side.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    r.stream = torch.cuda.current_stream(dev).cuda_stream
                    step()
            r.stream = None
            torch.cuda.current_stream(dev).wait_stream(side)
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize(dev)
            assert not r.finish()
        except Exception as e:                               # capture is an optimisation, never a requirement
            print("[bench] CUDA graph capture unavailable: %r" % (e,), file=sys.stderr)
            graph, r.stream = None, None
            torch.cuda.synchronize(dev)
    run_step = graph.replay if graph is not None else step
    # clock sampler first (its process start must not land between the barrier and the first event), then the windows
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    l0 = capi.launch_count()
    ms_step, windows, per_rank = timed_windows(run_step, args.steps, world, dev)
    launches = capi.launch_count() - l0
    if graph is not None:                                  # replays do not pass through the launch counter
        l0 = capi.launch_count()
        step(); torch.cuda.synchronize(dev)
        launches = (capi.launch_count() - l0) * args.steps
    else:
        launches //= len(windows)
    assert not r.finish(), "binning capacity overflow inside the timed region"
    clocks = sampler.stop() if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    value = P * world / (ms_step * 1e-3)

    # ---- per-stage timing for the roofline (separate pass; events on the launching stream) ---------------
    capi.profile_enable(True)
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize(dev)
    prof = capi.profile_read()
    capi.profile_enable(False)
    stage_ms = {k: (v[0] / v[1] if v[1] else 0.0) for k, v in prof.items()}
    A1, per_stage = algorithmic_bytes(P, V, Rn, H * W, M)
    dom = max((k for k in stage_ms if k in per_stage), key=lambda k: stage_ms[k])
    peak, peak_src = peaks()
    ach = per_stage[dom] / (stage_ms[dom] * 1e-3) / 1e9 if stage_ms[dom] > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                "frac": round(ach / peak, 4), "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(per_stage[dom]), "kernel_ms": round(stage_ms[dom], 4),
                "step_algorithmic_bytes": int(A1), "step_frac": round(A1 / (ms_step * 1e-3) / 1e9 / peak, 4),
                "stage_ms": {k: round(v, 4) for k, v in stage_ms.items() if v > 0}}
    tf = os.path.join(ROOT, "profiles", "traffic_%s.json" % dom)
    if os.path.isfile(tf):
        roofline["traffic"] = json.load(open(tf)).get("dram_bytes_per_launch")

    if args.kernel_only:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "ms_per_step": ms_step, "roofline": roofline, "kernel_only": True,
                              "config": {"P": P, "V": V, "R": Rn, "B": Bn}}))
        return
    # ---- e2e: reference-facing symbols, host inputs every step -------------------------------------------
    t = lambda a: torch.as_tensor(a, dtype=torch.float32)
    params = dict(means=gd["means"].clone().requires_grad_(True),
                  log_s=t(g["log_scales"]).to(dev).requires_grad_(True),
                  rot=gd["rots"].clone().requires_grad_(True),
                  op=t(g["opacity_logits"]).view(-1, 1).to(dev).requires_grad_(True),
                  dc=gd["dc"].view(P, 1, 3).clone().requires_grad_(True), sh=gd["sh"].clone().requires_grad_(True))
    cam_host = torch.cat([t(cam["view"]), t(cam["proj"]), t(cam["campos"])]).pin_memory()
    lims = [float(x) for x in cam["lims"]]
    bg = torch.zeros(3, device=dev)

    copy_stream = torch.cuda.Stream(dev)

    def e2e_step():
        # per-iteration host inputs (gaussian.cpp:678): the 24.9 MB ground-truth image goes up on a copy stream and
        # overlaps the forward; the camera block (140 B) is needed first and stays on the compute stream
        with torch.cuda.stream(copy_stream):
            gt_d = gt_host.to(dev, non_blocking=True)
        cam_d = cam_host.to(dev, non_blocking=True)
        rs = ops.GaussianRasterizationSettings(H, W, cam["tanfovx"], cam["tanfovy"], lims[0], lims[1], lims[2], lims[3], bg, 1.0,
                                               cam_d[:16].view(4, 4), cam_d[16:32].view(4, 4), deg, cam_d[32:35])
        means2D = torch.zeros_like(params["means"], requires_grad=True)  # renderer.cpp:29
        col, rad, _ = ops.GaussianRasterizer(rs)(params["means"], means2D, torch.sigmoid(params["op"]), params["dc"],
                                                 params["sh"], torch.exp(params["log_s"]),
                                                 torch.nn.functional.normalize(params["rot"]))
        torch.cuda.current_stream(dev).wait_stream(copy_stream)
        gt_d.record_stream(torch.cuda.current_stream(dev))
        loss = (1.0 - LAMBDA_DSSIM) * ops.l1_loss(col, gt_d) + LAMBDA_DSSIM * (1.0 - ops.fused_ssim(col.unsqueeze(0), gt_d.unsqueeze(0)))
        loss.backward()
        for p_ in params.values():
            p_.grad = None
        return float(loss.item())                                        # D2H of the step's result

    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    torch.cuda.synchronize(dev)
    e2e_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps, world)
    e2e = {"value": P * world / (e2e_ms * 1e-3), "unit": "Gaussians/s", "ms_per_step": round(e2e_ms, 4),
           "h2d_bytes_per_step": int(gt_host.numel() * 4 + cam_host.numel() * 4), "d2h_bytes_per_step": 4 + 8,
           "api": "RasterizeGaussiansCUDA/RasterizeGaussiansBackwardCUDA/fusedssim/fusedssim_backward (LibTorch shim) via autograd"}

    # ---- mapping iteration (BASELINE metric, second half): body of optimize()'s loop, gaussian.cpp:674-716 ----------
    # H2D image, render, loss, backward, [all-reduce of the 6 gradient tensors over NCCL], visibility-masked Adam.
    lrs = dict(means=1.6e-4, dc=2.5e-3, sh=2.5e-3 / 20.0, op=0.05, log_s=0.005, rot=0.001)      # config/fastlivo.yaml:18-22
    opt = ops.SparseGaussianAdam([(params[k], lrs[k]) for k in ("means", "dc", "sh", "op", "log_s", "rot")])

    _PACK = (("means", "dL_dmeans3D"), ("log_s", "dL_dscales"), ("rot", "dL_drots"), ("op", "dL_dopacity"),
             ("dc", "dL_ddc"), ("sh", "dL_dsh"))

    def mapping_iter():
        with torch.cuda.stream(copy_stream):
            gt_d = gt_host.to(dev, non_blocking=True)
        cam_d = cam_host.to(dev, non_blocking=True)
        rs = ops.GaussianRasterizationSettings(H, W, cam["tanfovx"], cam["tanfovy"], lims[0], lims[1], lims[2], lims[3], bg, 1.0,
                                               cam_d[:16].view(4, 4), cam_d[16:32].view(4, 4), deg, cam_d[32:35])
        means2D = torch.zeros_like(params["means"], requires_grad=True)
        col, rad, _ = ops.GaussianRasterizer(rs)(params["means"], means2D, torch.sigmoid(params["op"]), params["dc"],
                                                 params["sh"], torch.exp(params["log_s"]),
                                                 torch.nn.functional.normalize(params["rot"]))
        torch.cuda.current_stream(dev).wait_stream(copy_stream)
        gt_d.record_stream(torch.cuda.current_stream(dev))
        loss = (1.0 - LAMBDA_DSSIM) * ops.l1_loss(col, gt_d) + LAMBDA_DSSIM * (1.0 - ops.fused_ssim(col.unsqueeze(0), gt_d.unsqueeze(0)))
        loss.backward()
        visible = rad > 0
        if world > 1 and args.exchange == "p2p":                         # mean of per-view gradients, union of visibility
            for k_, n_ in _PACK:
                allreduce.grads[n_].view_as(params[k_].grad).copy_(params[k_].grad)
            _, vis8 = allreduce(rad)
            for k_, n_ in _PACK:
                params[k_].grad = allreduce.grads[n_].view_as(params[k_])
            visible = vis8.view(torch.bool)
        elif world > 1:
            vis8 = visible.to(torch.uint8)
            works = [dist.all_reduce(p_.grad, async_op=True) for p_ in params.values()]
            works.append(dist.all_reduce(vis8, op=dist.ReduceOp.MAX, async_op=True))
            for w_ in works:
                w_.wait()
            for p_ in params.values():
                p_.grad.mul_(1.0 / world)
            visible = vis8.bool()
        opt.set_visibility_and_N(visible, P)
        opt.step()
        opt.zero_grad()
        return loss

    for _ in range(3):
        mapping_iter()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0.record()
    for _ in range(args.steps):
        mapping_iter()
    e1.record()
    torch.cuda.synchronize(dev)
    map_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps, world)
    mapping = {"ms_per_iter": round(map_ms, 4), "views_per_iter": world,
               "what": "H2D image + render + L1/fused-SSIM loss + backward%s + SparseGaussianAdam (6 groups), LibTorch-shim symbols via autograd"
                       % ((" + %s all-reduce of grads" % ("NVLink-P2P (own kernels)" if args.exchange == "p2p" else "NCCL")) if world > 1 else "")}

    # ---- native mapping iteration (SURVEY 8f rank 1 + 3): packed model, C ABI only, one CUDA graph per rank ---------------
    # pinned GT image H2D (double-buffered on the copy stream) -> [graph: activations, forward, fused loss, backward, chain
    # rule, exchange, one-launch masked Adam].  Same work as `mapping_iter` without the torch autograd / optimizer plumbing.
    native = None
    try:
        from gaussian_lic_b200 import model as gmodel
        raw = dict(means=g["means"], log_scales=g["log_scales"], rots=g["rots"], opacity_logits=g["opacity_logits"],
                   dc=g["dc"], sh=g["sh"], degree=deg)
        mdl = gmodel.PackedModel(raw, dev, exchange=allreduce if (allreduce is not None and args.exchange == "p2p") else None)
        gts = [torch.empty_like(gt), torch.empty_like(gt)]
        graphs, done = [], [torch.cuda.Event(), torch.cuda.Event()]
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            r.stream = side.cuda_stream
            gts[0].copy_(gt_host, non_blocking=True); gts[1].copy_(gt_host, non_blocking=True)
            mdl.iteration(r, view, gts[0], color, T, radii, loss_out, dL)
            side.synchronize()
            for b in range(2):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=side):
                    r.stream = torch.cuda.current_stream(dev).cuda_stream
                    mdl.iteration(r, view, gts[b], color, T, radii, loss_out, dL)
                graphs.append(gr)
        r.stream = None
        torch.cuda.current_stream(dev).wait_stream(side)
        it_no = [0]

        def native_iter():
            b = it_no[0] & 1
            it_no[0] += 1
            copy_stream.wait_event(done[b])                  # the graph that last read this buffer has finished
            with torch.cuda.stream(copy_stream):
                gts[b].copy_(gt_host, non_blocking=True)
            torch.cuda.current_stream(dev).wait_stream(copy_stream)
            graphs[b].replay()
            done[b].record()

        done[0].record(); done[1].record()
        for _ in range(4):
            native_iter()
        torch.cuda.synchronize(dev)
        assert not r.finish()
        if world > 1:
            dist.barrier()
        e0.record()
        for _ in range(args.steps):
            native_iter()
        e1.record()
        torch.cuda.synchronize(dev)
        nat_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps, world)
        assert not r.finish()
        native = {"ms_per_iter": round(nat_ms, 4), "views_per_iter": world, "loss": float(loss_out.item()),
                  "what": "pinned GT H2D (double-buffered) + ONE CUDA graph: fused activations, forward, fused L1/D-SSIM loss, backward, "
                          "in-place chain rule%s, one-launch masked Adam on the packed model (C ABI only)"
                          % (", NVLink-P2P exchange" if mdl.exchange is not None else "")}
    except Exception as e:                                   # an extra metric must never cost the headline line
        print("[bench] native mapping iteration unavailable: %r" % (e,), file=sys.stderr)
        torch.cuda.synchronize(dev)

    if rank != 0:
        return
    cpu, _ = cpu_baseline(cfg, min(P, args.cpu_sample))
    out = {"metric": METRIC, "value": value, "unit": "Gaussians/s", "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
           "timing": {"how": "median of %d device-timed windows of exactly %d steps (barrier + sync both sides, max over ranks per window)"
                             % (len(windows), args.steps), "window_ms_per_step": windows, "per_rank_ms_per_step": per_rank},
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "%s: %d Gaussians, %dx%d, SH degree %d, forward + fused L1/D-SSIM loss + backward, "
                                  "1 view per GPU" % (cfg, P, W, H, deg),
                      "P": P, "V": V, "R": Rn, "B": Bn, "views_per_gpu": 1,
                      "parallelism": "dp%d (view-sharded%s)" % (world, (", in-place all-reduce of packed grads: %s" % ("own NVLink-P2P kernels" if args.exchange == "p2p" else "NCCL")) if world > 1 else ""),
                      "l2": "per-step working set ~%.1f GB > 126 MB L2 (no explicit flush)" % (A1 / 1e9),
                      "cuda_graph": graph is not None},
           "clocks": clocks, "e2e": e2e, "mapping_iter": mapping, "mapping_iter_native": native, "gpu_launches": int(launches), "roofline": roofline,
           "cpu_baseline": cpu}
    print(json.dumps(out))


def run_reference(args):
    """The reference's own implementation of the path on the same scene / metric (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import importlib.util
    syn = load_synthetic()
    cfg = args.config
    P, W, H, fx, fy, cx, cy, deg, zmax = syn.CONFIGS[cfg]
    so = os.path.join(ROOT, "oracle", "_ref", "glic_ref_ext.so")
    have_gpu = False
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        pass
    cpu, cpu_dt = cpu_baseline(cfg, min(P, args.cpu_sample))
    base = {"metric": METRIC, "unit": "Gaussians/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference", "cpu_baseline": cpu,
            "config": {"workload": "%s: %d Gaussians, %dx%d, SH degree %d, forward + fused-SSIM/L1 loss + backward" % (cfg, P, W, H, deg)}}
    if not (os.path.isfile(so) and have_gpu):
        # the reference has no CPU path: the only CPU implementation is the oracle port
        base.update({"value": cpu["value"], "ms_per_step": round(cpu_dt * 1e3 * P / min(P, args.cpu_sample), 2),
                     "reference_kind": "cpu oracle port (oracle/_ref not built or no GPU)",
                     "e2e": {"value": cpu["value"], "unit": "Gaussians/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(base))
        return
    import torch
    spec = importlib.util.spec_from_file_location("glic_ref_ext", so)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    dev = torch.device("cuda", 0)
    g, cam = syn.make_scene(cfg)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32)
    params = dict(means=t(g["means"]).to(dev).requires_grad_(True), log_s=t(g["log_scales"]).to(dev).requires_grad_(True),
                  rot=t(g["rots"]).to(dev).requires_grad_(True), op=t(g["opacity_logits"]).view(-1, 1).to(dev).requires_grad_(True),
                  dc=t(g["dc"]).view(P, 1, 3).to(dev).requires_grad_(True), sh=t(g["sh"]).to(dev).requires_grad_(True))
    gt_host = t(syn.make_gt_image(W, H)).pin_memory()
    cam_host = torch.cat([t(cam["view"]), t(cam["proj"]), t(cam["campos"])]).pin_memory()
    lims = [float(x) for x in cam["lims"]]
    bg = torch.zeros(3, device=dev)

    def step(host_inputs):
        gt_d = gt_host.to(dev, non_blocking=True) if host_inputs else step.gt_d
        cam_d = cam_host.to(dev, non_blocking=True) if host_inputs else step.cam_d
        means2D = torch.zeros_like(params["means"], requires_grad=True)
        col, rad, _ = ref.autograd_rasterize(params["means"], means2D, torch.sigmoid(params["op"]), params["dc"], params["sh"],
                                             torch.exp(params["log_s"]), torch.nn.functional.normalize(params["rot"]), bg,
                                             cam_d[:16].view(4, 4), cam_d[16:32].view(4, 4), cam_d[32:35], H, W,
                                             cam["tanfovx"], cam["tanfovy"], lims[0], lims[1], lims[2], lims[3], deg, False, 0.0)
        loss = (1.0 - LAMBDA_DSSIM) * ref.l1_autograd(col, gt_d) + LAMBDA_DSSIM * (1.0 - ref.fused_ssim_autograd(col.unsqueeze(0), gt_d.unsqueeze(0)))
        loss.backward()
        for p_ in params.values():
            p_.grad = None
        return float(loss.item()) if host_inputs else None

    step.gt_d, step.cam_d = gt_host.to(dev), cam_host.to(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = {}
    for host_inputs in (False, True):
        for _ in range(max(args.warmup, 3)):
            step(host_inputs)
        torch.cuda.synchronize()
        sampler = ClockSampler(0)
        sampler.start()
        e0.record()
        for _ in range(args.steps):
            step(host_inputs)
        e1.record()
        torch.cuda.synchronize()
        res[host_inputs] = (e0.elapsed_time(e1) / args.steps, sampler.stop())
    # mapping iteration with the reference's own optimiser (optim_utils.h SparseGaussianAdam -> adamUpdate)
    order = ("means", "dc", "sh", "op", "log_s", "rot")
    lr = [1.6e-4, 2.5e-3, 2.5e-3 / 20.0, 0.05, 0.005, 0.001]

    def mapping_iter():
        gt_d = gt_host.to(dev, non_blocking=True)
        cam_d = cam_host.to(dev, non_blocking=True)
        means2D = torch.zeros_like(params["means"], requires_grad=True)
        col, rad, _ = ref.autograd_rasterize(params["means"], means2D, torch.sigmoid(params["op"]), params["dc"], params["sh"],
                                             torch.exp(params["log_s"]), torch.nn.functional.normalize(params["rot"]), bg,
                                             cam_d[:16].view(4, 4), cam_d[16:32].view(4, 4), cam_d[32:35], H, W,
                                             cam["tanfovx"], cam["tanfovy"], lims[0], lims[1], lims[2], lims[3], deg, False, 0.0)
        loss = (1.0 - LAMBDA_DSSIM) * ref.l1_autograd(col, gt_d) + LAMBDA_DSSIM * (1.0 - ref.fused_ssim_autograd(col.unsqueeze(0), gt_d.unsqueeze(0)))
        loss.backward()
        ref.sparse_adam_step([params[k] for k in order], lr, rad > 0, P)
        for p_ in params.values():
            p_.grad = None

    map_ms = None
    if hasattr(ref, "sparse_adam_step"):
        for _ in range(3):
            mapping_iter()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            mapping_iter()
        e1.record()
        torch.cuda.synchronize()
        map_ms = e0.elapsed_time(e1) / args.steps
    ms, clocks = res[False]
    e2e_ms, _ = res[True]
    base.update({"value": P / (ms * 1e-3), "ms_per_step": round(ms, 4), "clocks": clocks,
                 "reference_kind": "reference CUDA sources (src/rasterizer, fused-ssim) compiled unmodified for sm_100a, "
                                   "driven through the reference's own autograd op (rasterizer.cpp) and loss_utils.h",
                 "mapping_iter": {"ms_per_iter": None if map_ms is None else round(map_ms, 4), "views_per_iter": 1,
                                  "what": "reference loop body: H2D image + render + loss + backward + reference SparseGaussianAdam"},
                 "e2e": {"value": P / (e2e_ms * 1e-3), "unit": "Gaussians/s", "ms_per_step": round(e2e_ms, 4),
                         "h2d_bytes_per_step": int(gt_host.numel() * 4 + cam_host.numel() * 4), "d2h_bytes_per_step": 4}})
    print(json.dumps(base))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4"])
    ap.add_argument("--cpu-sample", type=int, default=100_000, help="Gaussians in the bounded CPU-baseline sample")
    ap.add_argument("--view", type=int, default=None, help="view of the 8-view rig to render (default: rank % 8)")
    ap.add_argument("--rig-radius", type=float, default=0.25, help="radius [m] of the multi-view rig")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1 gradient exchange: own NVLink peer-memory all-reduce (default) or the NCCL library call")
    ap.add_argument("--no-graph", action="store_true", help="do not CUDA-graph the resident-input step")
    ap.add_argument("--kernel-only", action="store_true", help="skip the e2e and CPU legs (for ncu captures; not a bench value)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
