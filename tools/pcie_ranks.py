#!/usr/bin/env python
"""Developer aid: what the host link gives to 1, 2, 4 ... N GPUs AT THE SAME TIME (one process per GPU, as bench.py
runs), per direction and with both directions busy, with ordinary pinned memory and with write-combined pinned
memory for the upload side.  Answers where the host-fed `e2e` of `bench.py --gpus N` saturates.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/pcie_ranks.py [--out gpurun_out/pcie_ranks.json]

Transfers have the sizes of one bench step (200 MB up, 156 MB down, 8 chunks each, one cudaMemcpyAsync per chunk)."""
import ctypes
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist

UP, DN, CH, REPS = 200 * 2**20, 156 * 2**20, 8, 10


def wc_pinned(nbytes):
    """cudaHostAlloc(cudaHostAllocWriteCombined | cudaHostAllocPortable): not snooped on the way to the device"""
    rt = ctypes.CDLL("libcudart.so.12")
    p = ctypes.c_void_p()
    err = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04 | 0x01))
    if err != 0:
        raise RuntimeError(f"cudaHostAlloc(write-combined) failed: {err}")
    buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.uint8)


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout.strip()
    except Exception as e:
        return f"<{e}>"


def main():
    out_path = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    up_d = torch.empty(UP, dtype=torch.uint8, device=dev)
    dn_d = torch.empty(DN, dtype=torch.uint8, device=dev)
    dn_h = torch.empty(DN, dtype=torch.uint8).pin_memory()
    ups = {"pinned": torch.empty(UP, dtype=torch.uint8).pin_memory()}
    try:
        ups["write_combined"] = wc_pinned(UP)
        assert ups["write_combined"].is_pinned()
    except Exception as e:
        ups["write_combined"] = None
        wc_err = str(e)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(up_h, do_up, do_dn, reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s1.wait_event(a)
        s2.wait_event(a)
        for _ in range(reps):
            if do_up:
                with torch.cuda.stream(s1):
                    n = UP // CH
                    for c in range(CH):
                        up_d[c * n:(c + 1) * n].copy_(up_h[c * n:(c + 1) * n], non_blocking=True)
            if do_dn:
                with torch.cuda.stream(s2):
                    n = DN // CH
                    for c in range(CH):
                        dn_h[c * n:(c + 1) * n].copy_(dn_d[c * n:(c + 1) * n], non_blocking=True)
        e1, e2 = torch.cuda.Event(), torch.cuda.Event()
        e1.record(s1)
        e2.record(s2)
        torch.cuda.current_stream().wait_event(e1)
        torch.cuda.current_stream().wait_event(e2)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    res = {"world": world, "up_bytes": UP, "down_bytes": DN, "chunks": CH, "cases": []}
    sizes = [n for n in (1, 2, 4, 8) if n <= world]
    for kind, up_h in ups.items():
        if up_h is None:
            continue
        for n in sizes:
            for name, u, d in (("up", True, False), ("down", False, True), ("both", True, True)):
                if kind == "write_combined" and not u:
                    continue
                active = rank < n
                barrier()
                if active:
                    run(up_h, u, d, 2)
                barrier()
                ms = run(up_h, u, d, REPS) if active else 0.0
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                if world > 1:
                    allms = [torch.zeros_like(t) for _ in range(world)]
                    dist.all_gather(allms, t)
                    allms = [float(x.item()) for x in allms][:n]
                else:
                    allms = [ms]
                gb = ((UP if u else 0) + (DN if d else 0)) / 1e9
                res["cases"].append({"upload_memory": kind, "active_gpus": n, "direction": name, "ms_per_rank": [round(x, 3) for x in allms],
                                     "gbs_per_gpu_min": round(gb / max(allms) * 1e3, 1), "gbs_per_gpu_max": round(gb / min(allms) * 1e3, 1),
                                     "gbs_aggregate": round(n * gb / max(allms) * 1e3, 1)})
    if rank == 0:
        res["write_combined_error"] = None if ups.get("write_combined") is not None else wc_err
        res["topology"] = sh("nvidia-smi topo -m")
        res["numa"] = sh("lscpu | grep -i -E 'numa|socket|model name|^cpu\\(s\\)'")
        res["affinity_this_process"] = len(os.sched_getaffinity(0))
        res["meminfo"] = sh("grep -E 'MemTotal|HugePages_Total|Hugepagesize' /proc/meminfo")
        res["pcie_links"] = sh("nvidia-smi --query-gpu=index,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv,noheader")
        text = json.dumps(res, indent=1)
        if out_path:
            open(out_path, "w").write(text)
        for c in res["cases"]:
            print(f"{c['upload_memory']:15s} n={c['active_gpus']} {c['direction']:5s}: aggregate {c['gbs_aggregate']:7.1f} GB/s, per GPU {c['gbs_per_gpu_min']:.1f}-{c['gbs_per_gpu_max']:.1f}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
