"""CPU, world_size 2, gloo: the host-side sharding / broadcast / gather logic of the N>1 path."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from chinesechessai_b200 import dist as xd
    from chinesechessai_b200.neural_network import ChessNet
    torch.manual_seed(100 + rank)  # different weights per rank before the broadcast
    net = ChessNet(num_channels=8)
    sent = xd.broadcast_weights(net, src=0)
    chk = sum(float(p.double().sum()) for p in net.parameters())
    lo, hi = xd.shard_range(11, rank, world)
    boards = torch.full((hi - lo, 90), rank, dtype=torch.int8)
    ids = torch.arange(lo, hi, dtype=torch.int64)
    g = xd.gather_samples({"board": boards, "game": ids}, dst=0)
    res = dict(rank=rank, sent=sent, chk=chk, lo=lo, hi=hi)
    if rank == 0:
        res["games"] = g["game"].tolist()
        res["board_ranks"] = g["board"][:, 0].tolist()
    else:
        assert g is None
    out.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_broadcast_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0]["chk"] == res[1]["chk"]            # identical weights after the broadcast
    assert res[0]["sent"] == res[1]["sent"] > 0
    assert (res[0]["lo"], res[0]["hi"], res[1]["lo"], res[1]["hi"]) == (0, 6, 6, 11)
    assert res[0]["games"] == list(range(11))         # rank order, ragged shards
    assert res[0]["board_ranks"] == [0] * 6 + [1] * 5


def test_shard_range_partitions():
    from chinesechessai_b200.dist import shard_range
    for n in (0, 1, 7, 65536, 131072 + 3):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
