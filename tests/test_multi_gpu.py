"""GPU, world_size 2, NCCL: games sharded over two GPUs reproduce the single-GPU run of the same
game ids (SURVEY §7 test plan item 6); weights are identical after the broadcast.
Run with: gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu"""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from chinesechessai_b200 import dist as xd
    from chinesechessai_b200.mcts import HashEvaluator
    from chinesechessai_b200.neural_network import ChessNet
    from chinesechessai_b200.samples import training_tensors
    from chinesechessai_b200.self_play import BatchedSelfPlay
    torch.manual_seed(100 + rank)
    net = ChessNet(num_channels=16).cuda()
    samples, sp = xd.distributed_self_play(HashEvaluator(), 41, 15, 1.0, seed=5, network=net)
    chk = torch.tensor([sum(float(p.double().sum()) for p in net.parameters())], device="cuda",
                       dtype=torch.float64)
    both = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(both, chk)
    ok = True
    if rank == 0:
        ref = BatchedSelfPlay(HashEvaluator(), 41, 15, 1.0, seed=5)
        ref.play()
        want = training_tensors(ref)
        ok = all(torch.equal(samples[k], want[k]) for k in ("board", "player", "reward", "game", "ply"))
        ok = ok and float(both[0]) == float(both[1]) and len(want["reward"]) > 41 * 10
    else:
        ok = samples is None
    # a bf16 evaluator kept across two iterations must play the SECOND one with the weights of the
    # second broadcast (its folded inference copy is rebuilt): the trainer rank changes the
    # weights in between, every rank's visit counts change with them, and the ranks agree
    from chinesechessai_b200.mcts import NetEvaluator
    torch.manual_seed(7)                      # same initial weights everywhere
    net2 = ChessNet().cuda().eval()
    ev = NetEvaluator(net2, torch.bfloat16)
    first, _ = xd.distributed_self_play(ev, 16, 15, 1.0, seed=9, network=net2)
    if rank == 0:
        with torch.no_grad():
            for p in net2.parameters():
                p.add_(torch.randn_like(p) * 0.05)
    second, sp2 = xd.distributed_self_play(ev, 16, 15, 1.0, seed=9, network=net2)
    mine = sp2.rec_move[:8].to(torch.int64).sum().reshape(1)      # depends on the weights used
    alone = BatchedSelfPlay(NetEvaluator(net2, torch.bfloat16), 8, 15, 1.0, seed=9, first_game_id=rank * 8)
    alone.play()
    fresh = alone.rec_move[:8].to(torch.int64).sum().reshape(1)
    ok = ok and bool(torch.equal(mine, fresh))                    # = a fresh evaluator on the new weights
    if rank == 0:
        ok = ok and not torch.equal(first["board"], second["board"])   # and the games did change
    # one full iteration (broadcast -> sharded self-play -> gather -> update on rank 0)
    from chinesechessai_b200.iteration import self_play_iteration
    opt = torch.optim.Adam(net2.parameters(), lr=1e-3)
    it = self_play_iteration(net2, opt, 24, 15, seed=3, net_dtype=torch.bfloat16)
    ok = ok and it["plies"] > 0 and (rank != 0 or (it["samples"] >= 24 * 10 and it["loss"] == it["loss"]))
    out.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_shards_equal_single_gpu_run():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}
