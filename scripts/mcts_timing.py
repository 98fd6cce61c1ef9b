"""Component timing of one batched MCTS ply (cfg 3: 4,096 games x 15 sims) on one GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chinesechessai_b200.engine import BoardBatch, encode_planes, policy_priors
from chinesechessai_b200.mcts import BatchedMCTS, NetEvaluator, HashEvaluator
from chinesechessai_b200.neural_network import ChessNet

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 15
torch.manual_seed(0)
net = ChessNet().cuda().eval()
bb = BoardBatch(n)
bb.playout(1, 4)
m = BatchedMCTS(n, sims)

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

for name, ev in [("hash", HashEvaluator()), ("fp32", NetEvaluator(net, torch.float32)),
                 ("bf16", NetEvaluator(net, torch.bfloat16))]:
    t = timeit(lambda: m.search(bb.board, bb.meta, ev))
    print(f"search[{name}] n={n} sims={sims}: {t:.2f} ms/ply -> {n*sims/t*1e3:.3e} sims/s")
m.init(bb.board, bb.meta)
print("init    %.3f ms" % timeit(lambda: m.init(bb.board, bb.meta)))
print("select  %.3f ms (root wave)" % timeit(lambda: (m.init(bb.board, bb.meta), m.select(8))))
m.init(bb.board, bb.meta); m.select(8)
x = encode_planes(m.leaf_board, m.leaf_player)
print("encode  %.3f ms" % timeit(lambda: encode_planes(m.leaf_board, m.leaf_player)))
with torch.no_grad():
    print("forward fp32 %.3f ms" % timeit(lambda: net(x)))
    torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
    print("forward tf32 %.3f ms" % timeit(lambda: net(x)))
    xb = x.bfloat16()
    def fb():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return net(xb)
    print("forward bf16 autocast %.3f ms" % timeit(fb))
    netb = ChessNet().cuda().eval().bfloat16().to(memory_format=torch.channels_last)
    xcl = xb.to(memory_format=torch.channels_last)
    print("forward bf16 weights channels_last %.3f ms" % timeit(lambda: netb(xcl)))
    netb2 = ChessNet().cuda().eval().bfloat16()
    print("forward bf16 weights nchw %.3f ms" % timeit(lambda: netb2(xb)))
    logits, v = net(x)
print("priors  %.3f ms" % timeit(lambda: policy_priors(logits, m.leaf_moves, m.leaf_n)))
pri = policy_priors(logits, m.leaf_moves, m.leaf_n)
val = v.reshape(-1).float().contiguous()
print("backup  %.3f ms" % timeit(lambda: m.backup(pri, val)))
mv = bb.pick(1, 0)
bb.legal_moves()
print("legal_moves %.3f ms ; step %.3f ms" % (timeit(lambda: bb.legal_moves()), timeit(lambda: bb.step(torch.full_like(mv, -1)))))
