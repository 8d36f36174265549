"""N>1 host logic on CPU: frame partition + the single all-gather of detection records (world_size 2, gloo)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b2second import dist as b2dist


def test_frame_partition_is_contiguous_and_complete():
    for n, w in [(32, 8), (7, 2), (3, 4), (0, 2), (8, 1)]:
        spans = [b2dist.frames_for_rank(n, r, w) for r in range(w)]
        covered = [g for a, b in spans for g in range(a, b)]
        assert covered == list(range(n))


def test_pack_unpack_roundtrip():
    det = torch.randn(3, 5, 9)
    cnt = torch.tensor([5, 0, 2], dtype=torch.int32)
    d2, c2 = b2dist.unpack_records(b2dist.pack_records(det, cnt), 5, 9)
    assert torch.equal(d2, det) and torch.equal(c2, cnt)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, post_max, stride, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = b2dist.DetectionGatherer(B, post_max, stride, torch.device("cpu"))
    # rank r holds global frames r*B .. r*B+B-1; frame g has g+1 detections whose first field is 100*g + i
    det = torch.zeros(B, post_max, stride)
    cnt = torch.zeros(B, dtype=torch.int32)
    for lb in range(B):
        gidx = rank * B + lb
        cnt[lb] = gidx + 1
        for i in range(gidx + 1):
            det[lb, i, 0] = 100 * gidx + i
    d_all, c_all = g.gather(det, cnt)
    # the engine's record layout (detections followed by the count as a float) goes out without a packing step
    d_rec, c_rec = g.gather_records(b2dist.pack_records(det, cnt))
    ok = c_all.tolist() == [i + 1 for i in range(world * B)]
    ok &= bool(torch.equal(d_rec, d_all)) and bool(torch.equal(c_rec, c_all))
    for gidx in range(world * B):
        for i in range(gidx + 1):
            ok &= float(d_all[gidx, i, 0]) == 100 * gidx + i
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_all_gather_of_detections_world2_gloo():
    world, B = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, 6, 9, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=90) for _ in range(world)]
    for p in procs:
        p.join(30)
    assert sorted(res) == [(0, True), (1, True)]
