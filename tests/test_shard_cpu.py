"""Host-side sharding logic on CPU: range arithmetic and a world_size-2
gloo run of run_sharded / gather_in_order (the N>1 path of the benches, minus the kernels)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ivclab_b200.shard import gather_in_order, run_sharded, shard_range, shard_round_robin


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(1024, 3, 8) == (384, 512)            # cfg3: 1024 frames over 8 GPUs
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_round_robin_covers_sequences():
    got = sorted(sum((shard_round_robin(8, r, 3) for r in range(3)), []))   # cfg4: 8 sequences, 3 GPUs
    assert got == list(range(8))
    assert shard_round_robin(8, 5, 8) == [5]


def test_single_process_gather():
    out = run_sharded(5, lambda u: u * u)
    assert out == [0, 1, 4, 9, 16]
    assert gather_in_order(["b", "a"], [1, 0], 2) == ["a", "b"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, contiguous, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # "work" = what a rank returns per unit: an index tensor, a motion-vector array and a PSNR scalar
        def work(u):
            return {"unit": u, "rank": rank, "zz": torch.full((2, 3), u, dtype=torch.int32),
                    "mv": np.full((2, 2, 1), u, dtype=np.int64), "psnr": 30.0 + u}
        res = run_sharded(7, work, contiguous=contiguous)
        if rank == 0:
            q.put([(r["unit"], r["rank"], int(r["zz"][0, 0]), int(r["mv"][0, 0, 0]), r["psnr"]) for r in res])
        else:
            assert res is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("contiguous", [True, False])
def test_world_size_2_gloo_gather_in_unit_order(contiguous):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, contiguous, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [g[0] for g in got] == list(range(7))                       # unit order restored on rank 0
    assert all(g[2] == g[0] and g[3] == g[0] and g[4] == 30.0 + g[0] for g in got)
    owners = [g[1] for g in got]
    if contiguous:
        assert owners == [0, 0, 0, 0, 1, 1, 1]                         # contiguous frame ranges
    else:
        assert owners == [0, 1, 0, 1, 0, 1, 0]                         # round robin


def _rows_worker(rank, world, port, q):
    from ivclab_b200.shard import gather_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 7                                                      # uneven: shards of 4 and 3 units
        lo, hi = shard_range(n, rank, world)
        # the RD sweep's per-rank results: [Q, F_local] squared errors and [Q, F_local, bins] symbol histograms
        sse = torch.arange(lo, hi, dtype=torch.float64)[None, :] + torch.tensor([[0.0], [100.0]], dtype=torch.float64)
        hist = (torch.arange(lo, hi, dtype=torch.int32)[None, :, None] * 10 + torch.arange(5, dtype=torch.int32)[None, None, :]
                ).expand(2, hi - lo, 5).contiguous()
        out_sse = gather_rows(sse, n, axis=1)
        out_hist = gather_rows(hist, n, axis=1, out=torch.empty((2, n, 5), dtype=torch.int32) if rank == 0 else None)
        if rank == 0:
            q.put((out_sse.tolist(), out_hist.tolist()))
        else:
            assert out_sse is None and out_hist is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_gather_rows():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rows_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    sse, hist = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sse == [[float(u) for u in range(7)], [100.0 + u for u in range(7)]]
    assert hist == [[[10 * u + k for k in range(5)] for u in range(7)]] * 2


def test_gather_rows_single_process():
    from ivclab_b200.shard import gather_rows
    x = torch.arange(12).reshape(3, 4)
    assert gather_rows(x, 4, axis=1) is x
    with pytest.raises(ValueError):
        gather_rows(x, 5, axis=1)


def _shared_worker(rank, world, port, q):
    from ivclab_b200.shard import SharedPinned
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 7
        lo, hi = shard_range(n, rank, world)
        sp = SharedPinned((2, n, 3), torch.int32, "test", register=False)     # CPU test: shared, not page-locked
        # every rank writes its own unit range (on a GPU box: the destination of its device-to-host copies)
        sp.tensor[:, lo:hi] = torch.arange(lo, hi, dtype=torch.int32)[None, :, None] * 10 + rank
        dist.barrier()                                                       # the whole "gather"
        if rank == 0:
            q.put(sp.tensor.clone().tolist())
        path = sp._path
        sp.close()
        if rank == 0:
            assert not os.path.exists(path)
    finally:
        dist.destroy_process_group()


def test_world_size_2_shared_pinned_host_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shared_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [[[10 * u + (0 if u < 4 else 1)] * 3 for u in range(7)]] * 2
    assert got == want
