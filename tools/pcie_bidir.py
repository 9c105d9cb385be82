#!/usr/bin/env python
"""Developer aid: what the host link gives when uploads and downloads run at the same time (pinned buffers, two
streams, the sizes of one bench step: 200 MB up, 156 MB down), against each direction alone."""
import torch

up_h = torch.empty(200 * 2**20, dtype=torch.uint8).pin_memory()
dn_h = torch.empty(156 * 2**20, dtype=torch.uint8).pin_memory()
up_d = torch.empty_like(up_h, device="cuda")
dn_d = torch.empty_like(dn_h, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(do_up, do_dn, reps=10):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_event(a); s2.wait_event(a)
    for _ in range(reps):
        if do_up:
            with torch.cuda.stream(s1):
                for c in range(8):                      # chunked like the pipeline: 8 chunks per step
                    n = up_h.numel() // 8
                    up_d[c * n:(c + 1) * n].copy_(up_h[c * n:(c + 1) * n], non_blocking=True)
        if do_dn:
            with torch.cuda.stream(s2):
                for c in range(8):
                    n = dn_h.numel() // 8
                    dn_h[c * n:(c + 1) * n].copy_(dn_d[c * n:(c + 1) * n], non_blocking=True)
    e1, e2 = torch.cuda.Event(), torch.cuda.Event()
    e1.record(s1); e2.record(s2)
    torch.cuda.current_stream().wait_event(e1); torch.cuda.current_stream().wait_event(e2)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for name, u, d in (("up only", True, False), ("down only", False, True), ("both", True, True)):
    run(u, d, 3)
    ms = run(u, d)
    gb = ((up_h.numel() if u else 0) + (dn_h.numel() if d else 0)) / 1e9
    print(f"{name:10s}: {ms:.3f} ms per step-sized transfer set, {gb / ms * 1e3:.1f} GB/s in total")
