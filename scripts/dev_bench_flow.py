"""Developer timing of the flow (reverse) engine alone."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tts_b200.layers import ResidualCouplingBlocks
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = ResidualCouplingBlocks(192, 192, 5, 1, 4).eval().to(dev)
b, t = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 192
z = torch.randn(b, 192, t, device=dev)
mask = torch.ones(b, 1, t, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for do_flush in (0, 1):
    ts = []
    for i in range(8):
        if do_flush: flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = m(z, mask, reverse=True); e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1), 3))
    print(json.dumps({"T": t, "flush": do_flush, "ms": ts, "GFLOP": 14.16e6 * b * t / 1e9}))
