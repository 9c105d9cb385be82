#!/usr/bin/env python
"""Profiling target: one StreamedCoder.run over 16 bench-like 1080p frames with the luma planes derived on the device
(no CUDA graph, so that every kernel shows up under ncu):
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file l.csv python tools/e2e_launches.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivclab_b200 as ivc  # noqa: E402

F, H, W = 16, 1080, 1920
g = torch.Generator(device="cuda").manual_seed(0)
base = torch.randint(0, 256, (1, 3, H + 32, W + 32), generator=g, device="cuda").double()
canvas = (torch.nn.functional.avg_pool2d(base, 5, stride=1, padding=2) - 127.5) * 3.0 + 127.5
sh = torch.randint(-3, 4, (F + 1, 2), generator=g, device="cuda").cpu().tolist()
rgb = torch.stack([(canvas[0, :, 16 + dy:16 + dy + H, 16 + dx:16 + dx + W]
                    + 4.0 * torch.randn((3, H, W), generator=g, device="cuda", dtype=torch.float64)).clamp(0, 255).floor()
                   for dy, dx in sh]).permute(0, 2, 3, 1).contiguous().to(torch.uint8).cpu().pin_memory()
sc = ivc.StreamedCoder(1.0, 4, chunk_frames=4, use_graph=False)
out = sc.run(rgb[1:], first_ref=rgb[0])
torch.cuda.synchronize()
nb = F * (H // 8) * (W // 8) * 3
print(f"symbols per block: intra {sum(out['len_intra']) / nb:.2f}, inter {sum(out['len_inter']) / nb:.2f}")
