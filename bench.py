#!/usr/bin/env python
"""bench.py -- audio-samples/sec of batched VITS end-to-end inference (BASELINE.json configs[1]).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
  python bench.py --impl reference ...                      # the reference's CPU algorithm (oracle port) on host cores

A "step" is one pass of the hot path over one batch: Vits.inference on a [32, 64] token batch per GPU
(random-init VitsConfig, LJSpeech-shaped synthetic tokens), i.e. text encoder -> stochastic duration predictor
-> path expansion -> flow (reverse) -> HiFiGAN.  `value` times it with inputs resident in HBM; `e2e` times the
same call from pinned HOST buffers (tokens + SDP noise in, waveform out) per step.  Weak scaling: every rank
synthesises its own batch and rank 0 gathers the waveforms over NCCL inside the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU = 32
T_TEXT = 64
SR = 22050
HIFIGAN_FLOP_PER_SAMPLE = 2.402e6      # SURVEY.md section 8d (Cin=192)
FLOW_FLOP_PER_FRAME = 14.16e6
FP32_FMA_PEAK_TFLOPS = 73.5            # measured on this pool with tools/microbench_fma.cu (FFMA2), see DESIGN.md
# dram__bytes_read+write summed over the 78 conv launches of one HiFiGAN pass at this workload's shape (B=32, 192
# padded frames), from one ncu capture: profiles/r01_decoder_dram_traffic_final.csv (19.79 GB read + 8.89 GB written)
DECODER_DRAM_BYTES_PER_PASS = {6144: 28.68e9}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).

    NVML through pynvml when importable (one nvmlInit before the warm-up, then cheap per-sample queries from a
    thread); otherwise one looping `nvidia-smi -lms` process.  Either way nothing is spawned inside the timed region:
    attaching a new NVML client stalls the GPU for tens of milliseconds."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self._stop = index, [], None, None, False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            threading.Thread(target=self._poll_nvml, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        n = self.nvml
        bits = [(getattr(n, "nvmlClocksEventReasonHwSlowdown", getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8))),
                (getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40))),
                (getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20))),
                (getattr(n, "nvmlClocksEventReasonSwPowerCap", getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)))]
        try:
            mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        except Exception:
            mx = None
        while not self._stop:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = int(get(self.handle))
                self.rows.append([str(sm), str(mx)] + ["Active" if mask & b else "Not Active" for b in bits])
            except Exception:
                pass
            time.sleep(0.05)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Index of the next sample: call at the start of the timed region."""
        return len(self.rows)

    def stop(self, first=0):
        self._stop = True
        if self.proc is not None:
            self.proc.terminate()
        rows = self.rows[first:] or self.rows[-1:]
        self.rows = rows
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = [n for i, n in enumerate(self.NAMES) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self.nvml else "nvidia-smi"}


def build_model(seed=1234):
    import torch
    from tts_b200.vits import Vits, VitsConfig
    torch.manual_seed(seed)
    return Vits(VitsConfig()).eval()


def make_batch(rank, device=None):
    import torch
    gen = torch.Generator().manual_seed(4321 + rank)
    tokens = torch.randint(0, 100, (B_PER_GPU, T_TEXT), generator=gen)
    lengths = torch.full((B_PER_GPU,), T_TEXT, dtype=torch.int64)
    sdp_noise = torch.randn(B_PER_GPU, 2, T_TEXT, generator=gen)
    return tokens, lengths, sdp_noise


def host_threads():
    """Threads the CPU arm can really use: the affinity mask capped by the container's CPU quota.  (On this pool the
    GPU boxes report 128 logical CPUs but run under a 16-CPU cgroup quota; 128 torch threads then run ~100x slower
    than 16 -- measured with tools/probe_host_threads.py -- which would flatter the GPU/CPU ratio.)"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, -(-int(quota) // int(period))))
    except Exception:
        try:
            quota = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if quota > 0:
                n = min(n, max(1, -(-quota // period)))
        except Exception:
            pass
    return max(1, n)


def cpu_reference_samples_per_s(nbatch, steps=1, warmup=0, threads=None):
    """The reference's CPU algorithm (oracle port, bit-identical to the reference modules) on host cores."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vits_oracle as O
    from dataclasses import asdict
    # every host thread the container may use (torchrun exports OMP_NUM_THREADS=1; override it at run time)
    torch.set_num_threads(threads or host_threads())
    model = build_model()
    sd = model.state_dict()
    args = asdict(model.args)
    tokens, lengths, sdp_noise = make_batch(0)
    tokens, lengths, sdp_noise = tokens[:nbatch], lengths[:nbatch], sdp_noise[:nbatch]
    gen = torch.Generator().manual_seed(7)
    times, samples = [], 0
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            out = O.vits_inference(sd, tokens, lengths, sdp_noise, lambda s: torch.randn(s, generator=gen), args=args)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
                samples = int(out["y_lengths"].sum()) * 256
    return samples / (sum(times) / len(times)), sum(times) / len(times), samples, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nb = B_PER_GPU if host_threads() >= 8 else 4          # the whole batch per step when the host can afford it
    v, sec, samples, cores = cpu_reference_samples_per_s(nb, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": "audio_samples_per_sec", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "vits_e2e_inference_b32_t64 (BASELINE configs[1])", "tokens": T_TEXT,
                       "batch_per_gpu": B_PER_GPU, "sample": f"first {nb} of {B_PER_GPU} utterances of the batch per step"},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": f"{nb} of {B_PER_GPU} utterances, {samples} samples per step",
                             "logical_cpus": os.cpu_count()},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_cuda(args):
    import torch
    import torch.distributed as dist
    from tts_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    model = build_model().to(dev)
    tokens_h, lengths_h, sdp_noise_h = make_batch(rank)
    tokens_pin, lengths_pin, noise_pin = tokens_h.pin_memory(), lengths_h.pin_memory(), sdp_noise_h.pin_memory()
    tokens_d, lengths_d, noise_d = tokens_h.to(dev), lengths_h.to(dev), sdp_noise_h.to(dev)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def prior_noise(shape):
        return torch.randn(shape, generator=gen, device=dev, dtype=torch.float32)

    def gather(wav, y_lengths):
        if world == 1:
            return
        # the one collective on the data path: rank 0 gathers the (padded) waveforms + lengths over NVLink
        from tts_b200.parallel import gather_waveforms
        gather_waveforms(wav, y_lengths * 256, dst=0)

    def step_resident(stage_events=None):
        model._stage_events = stage_events
        out = model.inference(tokens_d, {"x_lengths": lengths_d}, sdp_noise=noise_d, prior_noise=prior_noise,
                              return_alignments=True)
        model._stage_events = None
        gather(out["model_outputs"], out["y_lengths"])
        return out

    host_wav = {}

    def step_e2e():
        tok = tokens_pin.to(dev, non_blocking=True)
        ln = lengths_pin.to(dev, non_blocking=True)
        nz = noise_pin.to(dev, non_blocking=True)
        out = model.inference(tok, {"x_lengths": ln}, sdp_noise=nz, prior_noise=prior_noise, return_alignments=True)
        gather(out["model_outputs"], out["y_lengths"])
        wav = out["model_outputs"]
        key = tuple(wav.shape)
        if key not in host_wav:
            host_wav[key] = torch.empty(wav.shape, dtype=wav.dtype).pin_memory()
        host_wav[key].copy_(wav, non_blocking=True)
        yl = out["y_lengths"].cpu()  # device->host read of the step's result (also synchronises)
        return out, wav.numel() * 4 + yl.numel() * 8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the sampler process attaches to the GPU when it starts (a tens-of-ms stall): start it before the warm-up and
    # only keep the samples taken inside the timed region
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()

    # ---------------- timed region 1: device-resident inputs
    first_sample = sampler.mark()
    launches0 = _lib.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stage_ev = []
    samples_rank = 0
    padded_samples = 0
    frames_padded = 0
    barrier()
    for s, e in ev:
        flush.fill_(1)       # evict L2 between timed iterations (outside the events)
        torch.cuda.synchronize()
        s.record()
        out = step_resident(stage_ev)
        e.record()
        samples_rank += int(out["y_lengths"].sum().item()) * 256
        padded_samples += out["model_outputs"].numel()
        frames_padded += out["y_mask"].shape[0] * out["y_mask"].shape[-1]
    barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop(first_sample)
    t_resident = sum(s.elapsed_time(e) for s, e in ev) / 1e3
    dec_ms = sum(a.elapsed_time(b) for n, a, b in stage_ev if n == "waveform_decoder")
    stage_ms = {}
    for n, a, b in stage_ev:
        stage_ms[n] = stage_ms.get(n, 0.0) + a.elapsed_time(b)
    if os.environ.get("BENCH_DEBUG") and rank == 0:
        print("per-step ms:", [round(s.elapsed_time(e), 2) for s, e in ev], file=sys.stderr)
        print("per-stage:", [(n, round(a.elapsed_time(b), 2)) for n, a, b in stage_ev], file=sys.stderr)

    # ---------------- timed region 2: end to end from pinned host buffers
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_samples, d2h = 0, 0
    for _ in range(args.steps):
        out, nbytes = step_e2e()
        e2e_samples += int(out["y_lengths"].sum().item()) * 256
        d2h = nbytes
    barrier()
    t_e2e = time.perf_counter() - t0

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    t_resident_max, t_e2e_max = allmax(t_resident), allmax(t_e2e)
    total_samples, total_e2e_samples = allsum(samples_rank), allsum(e2e_samples)

    if rank == 0:
        peaks = load_peaks()
        value = total_samples / t_resident_max
        e2e_value = total_e2e_samples / t_e2e_max
        dec_tflops = padded_samples * HIFIGAN_FLOP_PER_SAMPLE / (dec_ms / 1e3) / 1e12 if dec_ms > 0 else None
        h2d = tokens_pin.numel() * 8 + lengths_pin.numel() * 8 + noise_pin.numel() * 4
        if os.environ.get("BENCH_SKIP_CPU"):   # developer A/B runs only: the contract line always carries cpu_baseline
            cpu_nb, cpu_v, cpu_sec, cpu_samples, cores = 0, float("nan"), 0.0, 0, 0
        else:
            cpu_nb = B_PER_GPU if host_threads() >= 8 else 2
            cpu_v, cpu_sec, cpu_samples, cores = cpu_reference_samples_per_s(cpu_nb, steps=2, warmup=1)
        line = {
            "metric": "audio_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_resident_max / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "vits_e2e_inference_b32_t64 (BASELINE configs[1])", "batch_per_gpu": B_PER_GPU,
                       "tokens": T_TEXT, "frames_padded_per_step": frames_padded // args.steps,
                       "parallelism": f"dp{world}", "l2": "256 MiB buffer written between timed steps (outside the events)",
                       "rtf": (t_resident_max / args.steps) / (total_samples / args.steps / SR),
                       "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()}},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "conv1d_tc3_kernel + conv1d_tc3g_kernel<GRP,DIL> + conv1d_row1_kernel (the 78 HiFiGAN launches of a step; conv1d_tc3_kernel alone is 56 % of the step)",
                         "achieved": dec_tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": (dec_tflops / peaks["bf16_tflops_sustained"]) if dec_tflops else None,
                         "traffic": DECODER_DRAM_BYTES_PER_PASS.get(frames_padded // args.steps),
                         "traffic_unit": "bytes per decoder pass (78 launches), ncu dram__bytes_read+write; algorithmic layer-granular = 21.2 KB/sample",
                         "peak_source": peaks["source"],
                         "note": "algorithmic fp32 FLOPs; the MRF/pre convs run on tcgen05 kind::tf32 as 3xTF32 (3 MMAs per "
                                 "algorithmic MAC at half the bf16 rate => ceiling = peak/6); conv_post is a streaming FP32 kernel "
                                 f"(FP32 FMA peak measured {FP32_FMA_PEAK_TFLOPS} TFLOP/s)",
                         "frac_of_3xtf32_ceiling": (dec_tflops / (peaks["bf16_tflops_sustained"] / 6.0)) if dec_tflops else None,
                         "frac_fp32_fma": (dec_tflops / FP32_FMA_PEAK_TFLOPS) if dec_tflops else None},
            "cpu_baseline": {"value": cpu_v, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": f"{cpu_nb} of {B_PER_GPU} utterances, 1 warm-up + 2 timed steps ({cpu_samples} samples, {cpu_sec:.2f} s per step)",
                             "logical_cpus": os.cpu_count()},
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def emit(line):
    """The one JSON line of the contract, on the real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    # stdout carries exactly one JSON line: libraries that print to fd 1 (NCCL's "NCCL version ..." banner under
    # torchrun, NCCL_DEBUG output) are sent to stderr instead, the result is written to the saved descriptor
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
