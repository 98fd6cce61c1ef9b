"""HBM roofline of xq_bias_residual_relu_bf16 (the one bandwidth-bound kernel of the repo)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chinesechessai_b200.engine import bias_residual_relu
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
y = torch.randn(n, 128, 10, 9, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
x = torch.randn_like(y).contiguous(memory_format=torch.channels_last)
b = torch.randn(128, device="cuda").bfloat16()
out = torch.empty_like(y)
for _ in range(3): bias_residual_relu(y, x, b, out)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ms = 0.0
for _ in range(10):
    flush.zero_()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(); bias_residual_relu(y, x, b, out); e.record(); torch.cuda.synchronize()
    ms += a.elapsed_time(e)
ms /= 10
bytes_alg = 3 * y.numel() * 2
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6540.8
print(json.dumps({"kernel": "xq::bias_residual_relu_kernel", "n": n, "ms": ms, "algorithmic_bytes": bytes_alg,
                  "achieved_GBps": bytes_alg / ms / 1e6, "peak_GBps": peak, "frac": bytes_alg / ms / 1e6 / peak}))
