import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chinesechessai_b200.neural_network import ChessNet
from chinesechessai_b200.mcts import _FoldedNet
torch.manual_seed(0)
net = ChessNet().cuda().eval()
for m in net.modules():
    if isinstance(m, torch.nn.BatchNorm2d):
        m.running_mean.normal_(); m.running_var.uniform_(0.5, 2); m.weight.data.normal_(1, 0.1); m.bias.data.normal_()
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
with torch.no_grad():
    for n in (4096, 16384):
        x = torch.randn(n, 15, 10, 9, device="cuda")
        ref = net(x)
        for dt in (torch.bfloat16, torch.float16):
            f = _FoldedNet(net, dt)
            xb = x.to(dt)
            if f.in_channels == 16:
                xb = torch.cat([xb, torch.zeros_like(xb[:, :1])], 1)
            xb = xb.contiguous(memory_format=torch.channels_last)
            out = f(xb)
            err_p = (out[0][:, :8100].float() - ref[0]).abs().max().item()
            err_v = (out[1].float() - ref[1]).abs().max().item()
            agree = (out[0][:, :8100].float().argmax(1) == ref[0].argmax(1)).float().mean().item()
            if dt == torch.bfloat16 and f.stem_table is not None:
                from chinesechessai_b200.engine import BoardBatch
                bb = BoardBatch(n); bb.playout(1, 8)
                plr = bb.meta[:, 0].view(torch.int8)
                tb = timeit(lambda: f.forward_boards(bb.board, plr))
                print(f"   forward_boards (fused encode+stem lookup): {tb:.3f} ms -> {n*263.2e6/tb/1e9:.0f} TFLOP/s")
            t = timeit(lambda: f(xb))
            print(f"n={n} {dt} fused={f.fused}: {t:.3f} ms  -> {n*263.2e6/t/1e9:.0f} TFLOP/s  max|dlogit|={err_p:.3f} max|dv|={err_v:.4f} argmax agree={agree:.3f}")
            if dt == torch.bfloat16:
                f.own_epilogue = False
                t = timeit(lambda: f(xb))
                print(f"   cudnn conv+add+relu for the residual: {t:.3f} ms")
            f.fused = False
            t = timeit(lambda: f(xb))
            print(f"   unfused: {t:.3f} ms")
