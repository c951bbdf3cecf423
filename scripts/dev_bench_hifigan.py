"""Developer timing of the HiFiGAN engine alone (not the driver's bench.py)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tts_b200.hifigan import HifiganGenerator
from tts_b200 import _lib

def main():
    shapes = [(32, 150), (32, 1024)] if len(sys.argv) < 3 else [(int(sys.argv[1]), int(sys.argv[2]))]
    iters = int(os.environ.get("ITERS", "5"))
    dev = torch.device("cuda:0")
    net = HifiganGenerator(192, 1, "1", [[1, 3, 5]] * 3, [3, 7, 11], [16, 16, 4, 4], 512, [8, 8, 2, 2],
                           inference_padding=0, cond_channels=0, conv_pre_weight_norm=False,
                           conv_post_weight_norm=False, conv_post_bias=False).eval().to(dev)
    for b, t in shapes:
        x = torch.randn(b, 192, t, device=dev)
        for _ in range(2): y = net(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        e0.record()
        for _ in range(iters): y = net(x)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        samples = b * t * 256
        print(json.dumps({"B": b, "T": t, "ms": ms, "Msamples_s": samples / ms / 1e3,
                          "TFLOPs": samples * 2.402e6 / ms / 1e9, "launches": (_lib.launch_count() - n0) // iters}))
main()
