#!/usr/bin/env python
"""STREAM-style copy peaks measured the way MEASURED_PEAKS.json describes (torch b.copy_(a), read+write bytes, best of
10, CUDA events): once with a working set far larger than L2 (HBM) and once L2-resident (an L2 roofline
fraction may only be quoted against a peak measured like this)."""
import json, sys
import torch

def copy_gbs(nbytes, reps=10, inner=20):
    a = torch.empty(nbytes // 2, dtype=torch.bfloat16, device="cuda").normal_()
    b = torch.empty_like(a)
    for _ in range(3):
        b.copy_(a)
    best = 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * nbytes * inner / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best

out = {"hbm_copy_gbs_2GiB": copy_gbs(2 << 30, inner=2), "l2_copy_gbs_32MiB": copy_gbs(32 << 20),
       "l2_copy_gbs_16MiB": copy_gbs(16 << 20), "l2_copy_gbs_48MiB": copy_gbs(48 << 20),
       "how": "torch b.copy_(a), bf16, read+write bytes, best of 10 x 20 back-to-back copies (CUDA events); 16/32/48 MiB "
              "per buffer are L2-resident on a 126 MiB L2, 2 GiB is not",
       "gpu": torch.cuda.get_device_name(0)}
print(json.dumps(out))
