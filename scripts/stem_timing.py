"""Fused encode+conv1 lookup kernel vs encode + cuDNN conv+bias+ReLU for the network's first layer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chinesechessai_b200.engine import BoardBatch, encode_planes_nhwc16, stem_lookup
from chinesechessai_b200.mcts import _FoldedNet
from chinesechessai_b200.neural_network import ChessNet
torch.manual_seed(0)
f = _FoldedNet(ChessNet().cuda().eval(), torch.bfloat16)
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / reps
with torch.no_grad():
    for n in (4096, 16384):
        bb = BoardBatch(n); bb.playout(1, 8)
        pl = bb.meta[:, 0].view(torch.int8)
        t_lookup = timeit(lambda: stem_lookup(bb.board, pl, f.stem_table, f.stem_bias))
        t_enc = timeit(lambda: encode_planes_nhwc16(bb.board, pl))
        x = encode_planes_nhwc16(bb.board, pl)
        t_conv = timeit(lambda: f._cr(f.stem, x))
        out_mb = n * 90 * 128 * 2 / 1e6
        print(f"n={n}: lookup {t_lookup:.1f} us ({out_mb / t_lookup * 1e3 / 1e3:.0f} GB/s of output), "
              f"encode {t_enc:.1f} us + cuDNN stem {t_conv:.1f} us")
