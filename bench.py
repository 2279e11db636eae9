#!/usr/bin/env python
"""bench.py -- headline benchmark of the rasterizer hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--config cfg2]

Metric: Gaussians/s rasterised (forward + fused L1/D-SSIM loss + backward) at 1920x1080, SH degree 3
(BASELINE.json configs[1]: 500 k Gaussians, one view per GPU).  One JSON line on stdout (rank 0).

  value     : whole-job Gaussians/s with every input already resident in HBM: the native C++ mapper (csrc/mapper.cu)
              running activations + forward + fused loss + backward (+ the NVLink exchange at N > 1), optimiser off
  e2e       : the same step through the reference-facing LibTorch symbols (RasterizeGaussiansCUDA ...) under torch
              autograd, called from a C++ host (csrc/torch_host.cpp), with the per-iteration HOST inputs of the
              reference loop (gaussian.cpp:674-699): pinned ground-truth image + camera H2D every step,
              loss scalar D2H every step
  e2e_native: the same step through ONE synchronous C-ABI call of the native mapper per step (pinned image H2D, loss D2H)
  mapping_iter / mapping_iter_native : BASELINE's second half, the body of optimize()'s loop, through the LibTorch
              symbols and through the native mapper
  sort_cfg5 : BASELINE configs[4] (50 M key/value pairs), appended to the N = 1 line of both arms
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


_STDOUT_FD = None


def quiet_stdout():
    """Anything a library prints on fd 1 (NCCL's version banner, a stray warning) goes to stderr; the ONE JSON line is written
    to the real stdout by emit()."""
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _STDOUT_FD is None:
        sys.stdout.buffer.write(line)
        sys.stdout.flush()
    else:
        os.write(_STDOUT_FD, line)


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


def timed_windows(window, steps, world, dev, min_windows=5, min_total_ms=500.0, max_windows=60):
    """Device-timed windows of EXACTLY `steps` steps each.  `window()` runs the steps and returns the device time [ms] of
    the whole window, measured with a CUDA event pair on the stream the steps are launched on.  Every window is bracketed
    by a barrier + torch.cuda.synchronize() on both sides; a window's time is the MAX over ranks.  The reported figure is
    the MEDIAN window (>= 5 windows and >= 0.5 s of timed work in total), so one host-side hiccup on one rank (a 100 ms
    stall inside a 30 ms window was the round-1 N = 8 collapse) cannot set the number; every window is returned so the
    spread is visible.  -> (median ms/step, [window ms/step], per-rank ms/step of the median window)."""
    import torch
    if world > 1:
        import torch.distributed as dist
    wins, ranks = [], []
    n = min_windows
    i = 0
    while i < n:
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)
        mine = window() / steps
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        wins.append(max_over_ranks(mine, world))
        ranks.append(mine)
        if i == 0:                                           # identical on every rank (max-reduced): same window count
            n = int(min(max_windows, max(min_windows, -(-min_total_ms // max(wins[0] * steps, 1e-3)))))
        i += 1
    order = sorted(range(len(wins)), key=lambda k: wins[k])
    mid = order[len(order) // 2]
    return wins[mid], [round(w, 4) for w in wins], [round(x, 4) for x in gather_ranks(ranks[mid], world)]


def torch_window(fn, steps):
    """K calls of fn timed with a torch event pair on the current stream -> ms of the window."""
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def run():
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    return run


def sort_microbench(torch, ref=None, n=50_000_000, end_bit=45, reps=5):
    """BASELINE configs[4]: n (tile|depth) key/value pairs, bits [0, 45): our CUB-free onesweep (and, in the reference arm,
    cub::DeviceRadixSort::SortPairs as rasterizer_impl.cu:419-424 calls it) against the SURVEY 8(d) byte model."""
    syn = load_synthetic()
    keys_h, vals_h = syn.make_sort_pairs(n)
    dev = torch.device("cuda", torch.cuda.current_device())
    keys = torch.as_tensor(keys_h.view(np.int64)).to(dev)
    vals = torch.as_tensor(vals_h.view(np.int32)).to(dev)
    A = (8 + 24 * ((end_bit + 7) // 8)) * n
    peak, _ = peaks()
    out = {"n": n, "end_bit": end_bit, "A_sort_bytes": A}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_in, v_in, k_out, v_out = keys.clone(), vals.clone(), torch.empty_like(keys), torch.empty_like(vals)

    def timeit(fn):
        ts = []
        for _ in range(reps + 1):
            k_in.copy_(keys); v_in.copy_(vals)
            torch.cuda.synchronize()
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts[1:]))

    if ref is None:
        from gaussian_lic_b200 import capi
        lib = capi.lib
        temp = torch.empty(lib.glic_sort_temp_bytes(n), dtype=torch.uint8, device=dev)
        ms = timeit(lambda: capi.check(lib.glic_sort_pairs_u64_u32(n, end_bit, capi.ptr(k_in), capi.ptr(v_in), capi.ptr(k_out), capi.ptr(v_out),
                                                                   capi.ptr(temp), temp.numel(), None), "sort"))
        out.update({"impl": "own onesweep (csrc/radix_sort.cu)", "ms": round(ms, 4), "GBs": round(A / ms / 1e6, 1), "frac_of_hbm_peak": round(A / ms / 1e6 / peak, 4)})
    else:
        nbytes = ref.cub_sort_pairs(torch.empty(0, dtype=torch.uint8, device=dev), k_in, k_out, v_in, v_out, end_bit)
        temp = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        ms = timeit(lambda: ref.cub_sort_pairs(temp, k_in, k_out, v_in, v_out, end_bit))
        out.update({"impl": "cub::DeviceRadixSort::SortPairs (CUDA 12.9 CCCL)", "ms": round(ms, 4), "GBs": round(A / ms / 1e6, 1), "frac_of_hbm_peak": round(A / ms / 1e6 / peak, 4)})
    out["checksum"] = int((k_out[::9973].sum() ^ v_out[::9973].to(torch.int64).sum()).item())     # equal in both arms iff the outputs agree on the sample
    return out


def run_ours(args):
    import torch
    from gaussian_lic_b200 import capi, ops, mapper as gmapper, synthetic as syn
    rank, world, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
    cfg = args.config
    P, W, H, fx, fy, cx, cy, deg, zmax = syn.CONFIGS[cfg]
    k = max(1, args.views_per_gpu)
    S = world * k
    g, _ = syn.make_scene(cfg)
    M = g["sh"].shape[1]
    # the iteration's view batch: slot s renders view s % 8 of an 8-view ring of nearby keyframes around view 0,
    # as a sliding-window mapper would batch them; rank r renders slots [r*k, (r+1)*k)
    poses = [syn.orbit_pose(v, radius=args.rig_radius) for v in range(8)]
    gt_host = torch.as_tensor(syn.make_gt_image(W, H)).pin_memory()
    gt = gt_host.to(dev)
    steps = args.steps

    # ---- the product host: native mapper (csrc/mapper.cu), one per GPU ---------------------------------------------------
    mp = gmapper.Mapper(W, H, fx, fy, cx, cy, sh_degree=deg, capacity=P, rank=rank, world=world, views_per_rank=k)
    mp.initialize(g)
    for R_wc, t_wc in poses:
        mp.add_keyframe(R_wc, t_wc, gt)                       # keyframes 0..7: image RESIDENT in HBM (the `value` leg)
    for R_wc, t_wc in poses:
        mp.add_keyframe(R_wc, t_wc, gt_host)                  # keyframes 8..15: image in pinned host memory (H2D every iteration)
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, mp.export_handle())
        mp.connect(handles)
        dist.barrier()
    batch_res = [s % 8 for s in range(S)]
    batch_host = [8 + s % 8 for s in range(S)]

    def mapper_window(batch):
        def run():
            return mp.optimize(batch * steps).ms_optimize     # K iterations enqueued back to back, device-timed on the mapper's stream
        return run

    # ---- value: rasterization step (activations, forward, fused loss, backward, [exchange]) on resident inputs ----------
    mp.set_optimizer(False)
    mp.optimize(batch_res * max(args.warmup, 3))              # warm-up; settles the binning capacity
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    l0 = capi.launch_count()
    ms_step, windows, per_rank = timed_windows(mapper_window(batch_res), steps, world, dev)
    launches = (capi.launch_count() - l0) // len(windows)
    clocks = sampler.stop() if rank == 0 else None
    regrows = mp.stats().overflow_regrows
    value = P * S / (ms_step * 1e-3)

    # ---- per-stage timing for the roofline (events around each stage inside the library, on the launching stream) ---------
    r = ops.CRasterizer(W, H, dev)
    cam0 = gmapper.camera_block(W, H, fx, fy, cx, cy, *poses[(rank * k) % 8])
    gd = ops.scene_to_device(g, dev)
    _, _, radii0 = r.forward(gd, r.make_view(cam0))
    V = int((radii0 > 0).sum().item())
    Rn, Bn = r.R, r.B
    del r
    capi.profile_enable(True)
    mp.optimize(batch_res * steps)
    prof = capi.profile_read()
    capi.profile_enable(False)
    stage_ms = {kk: (v[0] / v[1] if v[1] else 0.0) for kk, v in prof.items()}
    for kk in ("sort", "zero"):                               # several recordings per view: report per view
        if prof[kk][1]:
            stage_ms[kk] = prof[kk][0] / (steps * k)
    A1, per_stage = algorithmic_bytes(P, V, Rn, H * W, M)
    dom = max((kk for kk in stage_ms if kk in per_stage), key=lambda kk: stage_ms[kk])
    peak, peak_src = peaks()
    ach = per_stage[dom] / (stage_ms[dom] * 1e-3) / 1e9 if stage_ms[dom] > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                "frac": round(ach / peak, 4), "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(per_stage[dom]), "kernel_ms": round(stage_ms[dom], 4),
                "step_algorithmic_bytes": int(A1 * k), "step_frac": round(A1 * k / (ms_step * 1e-3) / 1e9 / peak, 4),
                "stage_ms": {kk: round(v, 4) for kk, v in stage_ms.items() if v > 0}}
    tf = os.path.join(ROOT, "profiles", "traffic_%s.json" % dom)
    if os.path.isfile(tf):
        roofline["traffic"] = json.load(open(tf)).get("dram_bytes_per_launch")

    if args.kernel_only:
        if rank == 0:
            emit(({"metric": METRIC, "value": value, "ms_per_step": ms_step, "roofline": roofline, "kernel_only": True,
                              "config": {"P": P, "V": V, "R": Rn, "B": Bn}}))
        mp.close()
        return

    # ---- e2e through the NATIVE host: one synchronous C-ABI call per step (glic_mapper_optimize on one view batch): pinned image
    # H2D, the `value` step, loss D2H + host synchronisation every step; host wall clock over K calls, optimiser off ----------------
    mp.set_optimizer(False)
    for _ in range(3):
        mp.optimize(batch_host)

    def native_e2e_window():
        t0 = time.perf_counter()
        for _ in range(steps):
            mp.optimize(batch_host)                          # returns after the stream is drained and the loss has been read
        return (time.perf_counter() - t0) * 1e3

    e2n_ms, e2n_windows, _ = timed_windows(native_e2e_window, steps, world, dev, min_windows=3, min_total_ms=200.0)
    e2e_native = {"value": P * S / (e2n_ms * 1e-3), "unit": "Gaussians/s", "ms_per_step": round(e2n_ms, 4), "window_ms": e2n_windows,
                  "h2d_bytes_per_step": int(gt_host.numel() * 4 * k), "d2h_bytes_per_step": 4 + 32,
                  "api": "glic_mapper_optimize (C ABI, csrc/mapper.cu), one synchronous call per step: pinned keyframe image H2D, "
                         "activations + forward + loss + backward%s, loss + binning counters D2H, host wall clock"
                         % (" + NVLink exchange" if world > 1 else "")}
    # ---- native mapping iteration: pinned image H2D every iteration + the step above + the SH-rebuilding masked Adam --------
    mp.set_optimizer(True)
    mp.optimize(batch_host * 3)
    nat_ms, nat_windows, _ = timed_windows(mapper_window(batch_host), steps, world, dev, min_windows=3, min_total_ms=200.0)
    st = mp.stats()
    capi.profile_enable(True)
    mp.optimize(batch_host * steps)
    nprof = capi.profile_read()
    capi.profile_enable(False)
    nat_stage = {kk: round(v[0] / (steps * (k if kk not in ("adam", "allreduce") else 1)), 4) for kk, v in nprof.items() if v[1]}
    native = {"ms_per_iter": round(nat_ms, 4), "views_per_iter": S, "loss": float(st.last_loss), "window_ms": nat_windows,
              "stage_ms_per_iter": nat_stage,
              "what": "C++ mapper (csrc/mapper.cu, no torch, no Python in the loop): pinned keyframe image H2D on a copy stream "
                      "(double-buffered), fused activations, forward, fused L1/D-SSIM loss, backward to compact gradients%s, one masked "
                      "Adam launch that rebuilds dL/d(dc, sh) from dL/dcolour" % (
                          ", colour-gradient push + 11-float two-shot reduce over NVLink peer memory" if world > 1 else "")}
    mp.close()
    del mp

    # ---- e2e: the reference-facing LibTorch symbols driven by a C++ host (csrc/torch_host.cpp), host inputs every step -------
    t = lambda a: torch.as_tensor(a, dtype=torch.float32)
    view_id = (rank * k) % 8
    cam = cam0
    params = [t(g["means"]).to(dev), t(g["dc"]).view(P, 1, 3).to(dev), t(g["sh"]).to(dev), t(g["opacity_logits"]).view(-1, 1).to(dev),
              t(g["log_scales"]).to(dev), t(g["rots"]).to(dev)]                    # trainingSetup order (gaussian.cpp:399-424)
    lrs = [1.6e-4, 2.5e-3, 2.5e-3 / 20.0, 0.05, 0.005, 0.001]                     # config/fastlivo.yaml:18-22
    lims = [float(x) for x in cam["lims"]]
    host = ops.shim().MappingHost(params, lrs, H, W, deg, cam["tanfovx"], cam["tanfovy"], lims, LAMBDA_DSSIM)
    cam_host = torch.cat([t(cam["view"]), t(cam["proj"]), t(cam["campos"])]).pin_memory()
    allreduce = None
    if world > 1:
        from gaussian_lic_b200 import dist as gdist
        try:
            allreduce = gdist.P2PGradAllReduce(P, M, dev)
        except RuntimeError as e:                              # raised on every rank together: IPC not permitted here
            print("[bench] %s -- NCCL all-reduce for the reference-shaped legs" % (e,), file=sys.stderr)
            allreduce = gdist.GradAllReduce(P, M, dev)
    _PACK = ("dL_dmeans3D", "dL_ddc", "dL_dsh", "dL_dopacity", "dL_dscales", "dL_drots")

    def exchange(radii):
        """mean of the per-view gradients + union of visibility through the packed exchange buffer"""
        gr = host.grads()
        for gi, n_ in zip(gr, _PACK):
            allreduce.grads[n_].view_as(gi).copy_(gi)
        _, vis8 = allreduce(radii)
        for gi, n_ in zip(gr, _PACK):
            gi.copy_(allreduce.grads[n_].view_as(gi))
        return vis8.view(torch.bool)

    def e2e_step():
        if world == 1:
            return host.e2e_step(gt_host, cam_host)            # H2D image + camera, render, loss, backward, loss D2H: ONE C++ call
        loss, radii = host.forward_backward(gt_host, cam_host)
        exchange(radii)
        val = float(loss.item())
        host.optimizer_step(torch.zeros(P, dtype=torch.bool, device=dev))     # nothing visible: drops the gradients, moves nothing
        return val

    for _ in range(3):
        e2e_step()
    e2e_ms, e2e_windows, _ = timed_windows(torch_window(e2e_step, steps), steps, world, dev, min_windows=3, min_total_ms=200.0)
    e2e = {"value": P * world / (e2e_ms * 1e-3), "unit": "Gaussians/s", "ms_per_step": round(e2e_ms, 4), "window_ms": e2e_windows,
           "h2d_bytes_per_step": int(gt_host.numel() * 4 + cam_host.numel() * 4), "d2h_bytes_per_step": 8,
           "api": "RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA / fusedssim / fusedssim_backward (LibTorch shim) under torch "
                  "autograd, called from a C++ host (csrc/torch_host.cpp), one view per GPU%s" % (
                      "; N > 1: + mean all-reduce of the six gradient tensors (own NVLink-P2P kernels)" if world > 1 else "")}

    # ---- mapping iteration through the same symbols (BASELINE metric, second half; gaussian.cpp:674-716) ----------------------
    def mapping_iter():
        if world == 1:
            return host.mapping_iter(gt_host, cam_host)
        loss, radii = host.forward_backward(gt_host, cam_host)
        host.optimizer_step(exchange(radii))
        return loss

    for _ in range(3):
        mapping_iter()
    map_ms, _, _ = timed_windows(torch_window(mapping_iter, steps), steps, world, dev, min_windows=3, min_total_ms=200.0)
    mapping = {"ms_per_iter": round(map_ms, 4), "views_per_iter": world,
               "what": "H2D image + render + L1/fused-SSIM loss + backward%s + adamUpdate x 6 groups: LibTorch-shim symbols, C++ host"
                       % (" + mean all-reduce of the gradients" if world > 1 else "")}
    if rank != 0:
        return
    # ---- CPU legs (rank 0): the oracle port, median of 3; the forward-only cfg1 figure BASELINE configs[0] names ---------------
    cpus = [cpu_baseline(cfg, min(P, args.cpu_sample))[0] for _ in range(3)]
    cpu = sorted(cpus, key=lambda c: c["value"])[1]
    cpu["sample"] += "; median of 3 runs (%s)" % ", ".join("%.0f" % c["value"] for c in cpus)
    cpu["cfg1_forward_only"] = cpu_forward_cfg1()
    extra_sort = None
    if world == 1 and not args.no_sort_bench:
        try:
            extra_sort = sort_microbench(torch)
        except Exception as e:                                # an extra metric must never cost the headline line
            print("[bench] sort microbench skipped: %r" % (e,), file=sys.stderr)
    out = {"metric": METRIC, "value": value, "unit": "Gaussians/s", "n_gpus": world, "steps": steps,
           "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
           "timing": {"how": "median of %d device-timed windows of exactly %d steps (barrier + sync both sides, CUDA events on the "
                             "launching stream, max over ranks per window)" % (len(windows), steps),
                      "window_ms_per_step": windows, "per_rank_ms_per_step": per_rank},
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "%s: %d Gaussians, %dx%d, SH degree %d, forward + fused L1/D-SSIM loss + backward, %d view%s per GPU"
                                  % (cfg, P, W, H, deg, k, "" if k == 1 else "s"),
                      "P": P, "V": V, "R": Rn, "B": Bn, "views_per_gpu": k,
                      "parallelism": "dp%d (view-sharded%s)" % (world, ", exchange = NVLink push of per-view colour gradients + two-shot "
                                                                "mean-reduce of 11 geometric floats per Gaussian (own kernels / copy engines)" if world > 1 else ""),
                      "host": "C++ mapper (csrc/mapper.cu), plain launches on one stream, no host synchronisation inside the window",
                      "l2": "per-step working set ~%.1f GB > 126 MB L2 (no explicit flush)" % (A1 * k / 1e9),
                      "binning_overflow_regrows": int(regrows)},
           "clocks": clocks, "e2e": e2e, "e2e_native": e2e_native, "mapping_iter": mapping, "mapping_iter_native": native,
           "gpu_launches": int(launches),
           "roofline": roofline, "cpu_baseline": cpu}
    if extra_sort is not None:
        out["sort_cfg5"] = extra_sort
    emit((out))


def run_loop(args):
    """BASELINE configs[3] (cfg4): the full mapping loop -- per keyframe extend() (alpha-only render, LiDAR z-buffer
    de-duplication, in-place append) then optimize() (view sampler, 100 iterations of render + loss + backward + [exchange] +
    Adam) -- on the native mapper, 1 vs N GPUs.  mapping.cpp has no prune step (gaussian.cpp has none either); "densify" is
    extend().  With N GPUs an iteration consumes N sampled views, so a keyframe's 100 view-renders take ceil(100/N) steps."""
    import torch
    from gaussian_lic_b200 import mapper as gmapper, synthetic as syn
    rank, world, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
    cfg = args.config
    P, W, H, fx, fy, cx, cy, deg, zmax = syn.CONFIGS[cfg]
    K, n_lidar = args.keyframes, args.lidar_points
    g, _ = syn.make_scene(cfg)
    # a map under construction: the right-hand ~30 % of view 0's field of view has no Gaussians yet, so extend() has room to insert
    keep = g["means"][:, 0] < 0.4 * np.abs(g["means"][:, 2]) * W / (2 * fx)
    g = {kk: (v[keep] if isinstance(v, np.ndarray) and v.shape[:1] == keep.shape else v) for kk, v in g.items()}
    P = int(keep.sum())
    gts = [torch.as_tensor(syn.make_gt_image(W, H, seed=7 + i)).pin_memory() for i in range(min(K, 4))]
    mp = gmapper.Mapper(W, H, fx, fy, cx, cy, sh_degree=deg, capacity=P + K * n_lidar, max_iters=100, seed=1, rank=rank, world=world,
                        views_per_rank=1)
    mp.initialize(g)
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, mp.export_handle())
        mp.connect(handles)
        dist.barrier()
    rng = np.random.default_rng(5)
    ext_ms, opt_ms, inserted, iters = [], [], [], 0
    torch.cuda.synchronize(dev)
    t_wall = time.perf_counter()
    for kf in range(K):
        R_wc, t_wc = syn.orbit_pose(kf % 8, radius=args.rig_radius)
        mp.add_keyframe(R_wc, t_wc, gts[kf % len(gts)])
        z = rng.uniform(0.5, zmax, n_lidar)                                    # one LiDAR sweep inside this camera's frustum
        pc = np.stack([z * rng.uniform(-1.0, 1.0, n_lidar) * W / (2 * fx), z * rng.uniform(-1.0, 1.0, n_lidar) * H / (2 * fy), z], 1)
        pts = (pc @ np.asarray(R_wc).T + np.asarray(t_wc)).astype(np.float32)
        cols = rng.uniform(0, 1, (n_lidar, 3)).astype(np.float32)
        inserted.append(int(mp.extend(pts, cols, z.astype(np.float32))))
        st = mp.stats()
        ext_ms.append(float(st.ms_extend))
        st = mp.optimize()
        opt_ms.append(float(st.ms_optimize))
        iters = int(st.iterations)
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t_wall
    tot_opt = max_over_ranks(float(np.sum(opt_ms)), world)
    tot_ext = max_over_ranks(float(np.sum(ext_ms)), world)
    psnr, ssim = mp.evaluate(K - 1)
    st = mp.stats()
    mp.close()
    if rank != 0:
        return
    views = sum(min(i + 1, 100) for i in range(K))
    emit(({
        "metric": "mapping-iter ms (full loop)", "value": round(tot_opt / max(iters, 1), 4), "unit": "ms/iteration", "n_gpus": world,
        "higher_is_better": False, "scaling": "strong", "dtype": "f32", "data": "synthetic", "vs_baseline": None,
        "config": {"workload": "%s recipe: %d initial Gaussians (right-hand 30 %% of the view left empty), %dx%d, SH degree %d, %d keyframes x (extend %d LiDAR points + optimize <= 100 views)"
                               % (cfg, P, W, H, deg, K, n_lidar), "views_per_iteration": world},
        "iterations": iters, "view_renders": views, "ms_per_view_render": round(tot_opt / max(views, 1), 4), "optimize_ms_total": round(tot_opt, 2),
        "extend_ms_total": round(tot_ext, 2), "extend_ms_mean": round(tot_ext / K, 3), "inserted_per_keyframe": inserted,
        "final_gaussians": int(st.num_gaussians), "capacity": int(st.capacity), "overflow_regrows": int(st.overflow_regrows),
        "last_keyframe_psnr": round(psnr, 3), "last_keyframe_ssim": round(ssim, 4), "last_loss": float(st.last_loss),
        "wall_s_including_host_scene_generation": round(wall, 2)}))


def cpu_forward_cfg1():
    """BASELINE configs[0]: 10 k Gaussians, 640x480, SH degree 0, forward only, on the host cores (the reference has no CPU
    path: this is the oracle port), 1 thread and all physical cores."""
    from oracle.oracle import Oracle
    syn = load_synthetic()
    o = Oracle(np.float32)
    g, cam = syn.make_scene("cfg1")
    res = {}
    for label, n in (("1_thread", 1), ("all_cores", physical_cores())):
        o.set_threads(n)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            f = o.forward(g, cam)
            ts.append(time.perf_counter() - t0)
            o.free(f)
        res[label] = {"threads": n, "ms": round(1e3 * float(np.median(ts)), 2), "Gaussians_per_s": round(10_000 / float(np.median(ts)), 1)}
    return res


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
        emit((base))
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
    if not args.no_sort_bench and hasattr(ref, "cub_sort_pairs"):
        try:
            del params
            torch.cuda.empty_cache()
            base["sort_cfg5"] = sort_microbench(torch, ref=ref)
        except Exception as e:
            print("[bench] sort microbench skipped: %r" % (e,), file=sys.stderr)
    emit((base))


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
    ap.add_argument("--views-per-gpu", type=int, default=1, help="views each GPU renders per iteration (gradient accumulation)")
    ap.add_argument("--loop", action="store_true", help="run the full mapping loop (extend + optimize per keyframe) on --config; default for cfg4")
    ap.add_argument("--keyframes", type=int, default=45, help="--loop: keyframes (each: extend + one optimisation pass over the keyframes so far, <= 100 views): 45 keyframes = 1035 view renders")
    ap.add_argument("--lidar-points", type=int, default=20000, help="--loop: LiDAR points offered to extend() per keyframe")
    ap.add_argument("--no-sort-bench", action="store_true", help="skip the cfg5 sort microbenchmark appended to the N = 1 line")
    ap.add_argument("--no-graph", action="store_true", help="(kept for old command lines; the mapper launches plainly)")
    ap.add_argument("--kernel-only", action="store_true", help="skip the e2e and CPU legs (for ncu captures; not a bench value)")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.loop or args.config == "cfg4":
        run_loop(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
