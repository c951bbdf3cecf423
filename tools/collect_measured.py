"""Turns the raw captures of tools/final_r02.sh (gpurun_out/final2/) into the committed evidence under profiles/:
  profiles/r02_bench_step_breakdown.txt   per-kernel share of one bench step (ncu launch list)
  profiles/r02_decoder_dram_traffic.csv   ncu dram bytes per launch of one dense decoder pass (B=32 x 192 frames)
  profiles/measured.json                  numbers bench.py reads (decoder DRAM bytes per pass keyed by padded frames, peaks)
Run here (no GPU needed):  python tools/collect_measured.py [gpurun_out/final2]"""
import collections
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "final2")
prof = os.path.join(ROOT, "profiles")


def rows_of(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 6]
    hdr = rows[0]
    return hdr, rows[1:]


# ---- one bench step
hdr, data = rows_of(os.path.join(src, "bench_launches.csv"))
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
starts = [i for i, r in enumerate(data) if "embed_kernel" in r[ik]]
step = data[starts[-1]:]
agg, tot = collections.OrderedDict(), 0.0
for r in step:
    n, t = r[ik].split("(")[0][:70], float(r[iv]) / 1e6
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += t
    tot += t
with open(os.path.join(prof, "r02_bench_step_breakdown.txt"), "w") as f:
    f.write("# One bench.py step (the last of `bench.py --steps 1 --warmup 3`, ragged batch) from gpurun_out/final2/bench_launches.csv\n"
            "# ncu --metrics gpu__time_duration.sum --clock-control none: per-launch times are cold-cache and serialised;\n"
            "# the SHARE per kernel is what is comparable with the live CUDA-event stage times in the bench line.\n")
    f.write(f"# launches {len(step)}, sum of kernel durations {tot:.2f} ms\n       ms  share    n  kernel\n")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{t:9.3f}  {100 * t / tot:4.1f}%  {c:3d}  {n}\n")
shutil.copy(os.path.join(src, "bench_launches.csv"), os.path.join(prof, "r02_bench_launches.csv"))

# ---- decoder DRAM traffic (dense pass)
hdr, data = rows_of(os.path.join(src, "decoder_dram.csv"))
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
iid = hdr.index("ID")
per = collections.OrderedDict()
for r in data:
    d = per.setdefault(r[iid], {"k": r[ik].split("(")[0][-50:]})
    d[r[im]] = float(r[iv])
launches = list(per.values())
# dev_bench_hifigan.py runs 2 warm-up passes + ITERS=1 timed pass: keep the last third
n = len(launches) // 3
last = launches[-n:]
rd = sum(l.get("dram__bytes_read.sum", 0) for l in last)
wr = sum(l.get("dram__bytes_write.sum", 0) for l in last)
with open(os.path.join(prof, "r02_decoder_dram_traffic.csv"), "w") as f:
    f.write("kernel,dram_read_bytes,dram_write_bytes,time_ns\n")
    for l in last:
        f.write(f"{l['k']},{l.get('dram__bytes_read.sum', 0):.0f},{l.get('dram__bytes_write.sum', 0):.0f},{l.get('gpu__time_duration.sum', 0):.0f}\n")
    f.write(f"# one dense decoder pass, B=32 x 192 frames: {len(last)} launches, {rd / 1e9:.2f} GB read + {wr / 1e9:.2f} GB written\n")

mp = os.path.join(prof, "measured.json")
m = json.load(open(mp))
m.setdefault("decoder_dram_bytes_per_pass", {})["6144"] = rd + wr
m["decoder_dram_source"] = ("profiles/r02_decoder_dram_traffic.csv (ncu dram__bytes_read+write over the launches of one DENSE decoder pass, "
                            "B=32 x 192 frames, r02 kernels)")
json.dump(m, open(mp, "w"), indent=1)
print(f"step: {len(step)} launches {tot:.2f} ms; decoder pass {len(last)} launches {(rd + wr) / 1e9:.2f} GB")
for name in ("bench_1gpu.json", "bench_1gpu_dense.json", "sanitizer.txt", "ncu_tc3_c128k11.details.csv", "ncu_mas2.details.csv"):
    p = os.path.join(src, name)
    if os.path.exists(p):
        shutil.copy(p, os.path.join(prof, "r02_" + name))
