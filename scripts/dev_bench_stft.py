"""Developer timing of the STFT / mel front end kernels (a9): wav_to_spec + spec_to_mel on B x 10 s of audio."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tts_b200.audio import spec_to_mel, wav_to_spec

dev = torch.device("cuda:0")
b, t = 32, 220500
wav = torch.rand(b, 1, t, device=dev) * 2 - 1
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}
spec = wav_to_spec(wav, 1024, 256, 1024)
for name, fn in (("wav_to_spec", lambda: wav_to_spec(wav, 1024, 256, 1024)),
                 ("spec_to_mel", lambda: spec_to_mel(spec, 1024, 80, 22050, 0, None))):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(5):
        flush.fill_(0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res[name] = sum(ts) / len(ts)
frames = spec.shape[-1]
in_bytes, spec_bytes, mel_bytes = wav.numel() * 4, spec.numel() * 4, b * 80 * frames * 4
res["frames"] = b * frames
res["wav_to_spec_GBs"] = (in_bytes + spec_bytes) / res["wav_to_spec"] / 1e6     # read the hop once, write 513 bins
res["spec_to_mel_GBs"] = (spec_bytes + mel_bytes) / res["spec_to_mel"] / 1e6
print(json.dumps(res))
