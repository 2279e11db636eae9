"""Where does the end-to-end (LibTorch-shim, autograd) step spend its time?  Host enqueue time vs device time, and the
kernel list of one step (torch.profiler).  Diagnostic only; prints a small JSON + table."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_lic_b200 import ops, synthetic as syn  # noqa: E402

cfg = "cfg2"
P, W, H, fx, fy, cx, cy, deg, zmax = syn.CONFIGS[cfg]
dev = torch.device("cuda:0")
g, cam = syn.make_scene(cfg)
gd = ops.scene_to_device(g, dev)
t = lambda a: torch.as_tensor(a, dtype=torch.float32)
params = dict(means=gd["means"].clone().requires_grad_(True), log_s=t(g["log_scales"]).to(dev).requires_grad_(True),
              rot=gd["rots"].clone().requires_grad_(True), op=t(g["opacity_logits"]).view(-1, 1).to(dev).requires_grad_(True),
              dc=gd["dc"].view(P, 1, 3).clone().requires_grad_(True), sh=gd["sh"].clone().requires_grad_(True))
gt_host = torch.as_tensor(syn.make_gt_image(W, H)).pin_memory()
cam_host = torch.cat([t(cam["view"]), t(cam["proj"]), t(cam["campos"])]).pin_memory()
lims = [float(x) for x in cam["lims"]]
bg = torch.zeros(3, device=dev)
copy_stream = torch.cuda.Stream(dev)


def step(read_loss=True):
    with torch.cuda.stream(copy_stream):
        gt_d = gt_host.to(dev, non_blocking=True)
    cam_d = cam_host.to(dev, non_blocking=True)
    rs = ops.GaussianRasterizationSettings(H, W, cam["tanfovx"], cam["tanfovy"], lims[0], lims[1], lims[2], lims[3], bg, 1.0,
                                           cam_d[:16].view(4, 4), cam_d[16:32].view(4, 4), deg, cam_d[32:35])
    means2D = torch.zeros_like(params["means"], requires_grad=True)
    col, rad, _ = ops.GaussianRasterizer(rs)(params["means"], means2D, torch.sigmoid(params["op"]), params["dc"], params["sh"],
                                             torch.exp(params["log_s"]), torch.nn.functional.normalize(params["rot"]))
    torch.cuda.current_stream(dev).wait_stream(copy_stream)
    loss = 0.8 * ops.l1_loss(col, gt_d) + 0.2 * (1.0 - ops.fused_ssim(col.unsqueeze(0), gt_d.unsqueeze(0)))
    loss.backward()
    for p_ in params.values():
        p_.grad = None
    return float(loss.item()) if read_loss else None


for _ in range(5):
    step()
torch.cuda.synchronize()
N = 30
t0 = time.perf_counter()
for _ in range(N):
    step(False)
t_host = (time.perf_counter() - t0) / N * 1e3
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / N * 1e3
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(N):
    step(True)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"host_enqueue_ms": round(t_host, 3), "async_total_ms": round(t_all, 3),
                  "with_loss_readback_ms": round(e0.elapsed_time(e1) / N, 3)}))
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        step(True)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
prof.export_chrome_trace(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "e2e_trace.json"))
