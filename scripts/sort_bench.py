"""cfg5 sort microbenchmark: 50 M (tile << 32 | depth bits, index) pairs, bits [0, 45).
Ours (glic_sort_pairs_u64_u32, CUB-free onesweep) vs the reference's cub::DeviceRadixSort::SortPairs call
(oracle/_ref, when built), same buffers, CUDA-event timed, A_sort = 152 B/pair.  Prints one JSON line."""
import importlib.util
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gaussian_lic_b200 import capi, synthetic as syn  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
END_BIT, ITERS = 45, 10
keys_np, vals_np = syn.make_sort_pairs(n)
dev = torch.device("cuda:0")
keys = torch.from_numpy(keys_np.view("int64")).to(dev)
vals = torch.from_numpy(vals_np.view("int32")).to(dev)
lib = capi.lib


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ITERS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / ITERS


k_in, v_in = keys.clone(), vals.clone()
k_out, v_out = torch.empty_like(keys), torch.empty_like(vals)
temp = torch.empty(lib.glic_sort_temp_bytes(n), dtype=torch.uint8, device=dev)


def ours():
    k_in.copy_(keys); v_in.copy_(vals)                     # the sort ping-pongs through its inputs
    capi.check(lib.glic_sort_pairs_u64_u32(n, END_BIT, capi.ptr(k_in), capi.ptr(v_in), capi.ptr(k_out), capi.ptr(v_out),
                                           capi.ptr(temp), temp.numel(), None), "sort_pairs")


def copies():
    k_in.copy_(keys); v_in.copy_(vals)


ms_copy = timed(copies)
ms_ours = timed(ours) - ms_copy
out = {"n": n, "end_bit": END_BIT, "A_sort_bytes": 152 * n, "ours_ms": round(ms_ours, 4),
       "ours_GBs": round(152 * n / ms_ours / 1e6, 1)}
ref_so = os.path.join(ROOT, "oracle", "_ref", "glic_ref_ext.so")
if os.path.isfile(ref_so):
    spec = importlib.util.spec_from_file_location("glic_ref_ext", ref_so)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    if hasattr(ref, "cub_sort_pairs"):
        rk, rv = torch.empty_like(keys), torch.empty_like(vals)
        need = ref.cub_sort_pairs(torch.empty(0, dtype=torch.uint8, device=dev), keys, rk, vals, rv, END_BIT)
        rtemp = torch.empty(need, dtype=torch.uint8, device=dev)
        ms_cub = timed(lambda: ref.cub_sort_pairs(rtemp, keys, rk, vals, rv, END_BIT))
        ours()
        torch.cuda.synchronize()
        out.update({"cub_ms": round(ms_cub, 4), "cub_GBs": round(152 * n / ms_cub / 1e6, 1),
                    "identical_to_cub": bool(torch.equal(rk, k_out) and torch.equal(rv, v_out))})
print(json.dumps(out))
