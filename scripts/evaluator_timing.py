#!/usr/bin/env python
"""Leaf-evaluator throughput by precision / layout: rows per second of NetEvaluator on a batch of
middle-game positions (what one MCTS wave sends to the network).  GPU only; prints one JSON line."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from chinesechessai_b200.engine import BoardBatch  # noqa: E402
from chinesechessai_b200.mcts import NetEvaluator  # noqa: E402
from chinesechessai_b200.neural_network import ChessNet  # noqa: E402

FLOP_PER_ROW = None


def flops_per_row(net):
    f = 0
    for m in net.modules():
        if isinstance(m, torch.nn.Conv2d):
            f += 2 * 90 * m.out_channels * m.in_channels * m.kernel_size[0] * m.kernel_size[1]
        elif isinstance(m, torch.nn.Linear):
            f += 2 * m.in_features * m.out_features
    return f


def main():
    torch.manual_seed(0)
    net = ChessNet().cuda().eval()
    fl = flops_per_row(net)
    out = {"flop_per_row": fl}
    variants = {
        "bf16_folded": lambda: NetEvaluator(net, torch.bfloat16),
        "tf32_folded": lambda: NetEvaluator(net, torch.float32, tf32=True),
        "tf32_module": lambda: NetEvaluator(net, torch.float32, tf32=True, folded=False),
        "fp32_folded": lambda: NetEvaluator(net, torch.float32, tf32=False, folded=True),
        "fp32_module": lambda: NetEvaluator(net, torch.float32, tf32=False),
    }
    for rows in (4096, 16384):
        bb = BoardBatch(rows)
        bb.playout(11, 12)
        mv, nm = bb.legal_moves()
        pl = bb.meta[:, 0].view(torch.int8).contiguous()
        for name, mk in variants.items():
            ev = mk()
            for _ in range(3):
                ev(bb.board, pl, mv, nm)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10 if "fp32" in name else 30
            a.record()
            for _ in range(reps):
                ev(bb.board, pl, mv, nm)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
            out[f"{name}@{rows}"] = {"ms": round(ms, 3), "rows_per_s": round(rows / ms * 1e3),
                                     "TFLOPs": round(rows * fl / ms / 1e9, 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
