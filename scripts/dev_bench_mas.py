"""Developer timing of the MAS kernel at BASELINE config 4 (B=512, Tx=200, Ty=1000)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tts_b200.helpers import maximum_path_lengths, maximum_path

dev = torch.device("cuda:0")
b, tx, ty = 512, 200, 1000
torch.manual_seed(0)
v = torch.randn(b, tx, ty, device=dev)
t_x = torch.full((b,), tx, dtype=torch.int32, device=dev)
t_y = torch.full((b,), ty, dtype=torch.int32, device=dev)
mask = torch.ones(b, tx, ty, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}
for name, fn in (("lengths_i32", lambda: maximum_path_lengths(v, t_x, t_y)),
                 ("lengths_f32", lambda: maximum_path_lengths(v, t_x, t_y, out_dtype=torch.float32)),
                 ("drop_in_mask", lambda: maximum_path(v, mask))):
    for _ in range(3): fn()
    ts = []
    for _ in range(int(os.environ.get("ITERS", "10"))):
        flush.fill_(0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sum(ts) / len(ts)
    res[name] = {"ms": ms, "GBs_algorithmic_8B_per_cell": b * tx * ty * 8 / ms / 1e6, "Gcell_s": b * tx * ty / ms / 1e6}
print(json.dumps(res))
