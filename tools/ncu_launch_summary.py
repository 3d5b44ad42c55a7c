"""Summarises an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` launch list
(one graph replay inside the NVTX range of tools/profile_step.py) per kernel: python tools/ncu_launch_summary.py in.csv"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
per = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
for d in csv.DictReader(lines):
    name = re.sub(r"[<(].*", "", d["Kernel Name"]).replace("void ", "").replace("sdeo::", "")
    name = name.replace("conv_gemm_kernel_occ2", "conv_gemm_kernel")   # (the 80-register entry point of the same kernel body)
    v = float(d["Metric Value"].replace(",", ""))
    unit, m = d["Metric Unit"], d["Metric Name"]
    if m == "gpu__time_duration.sum":
        cnt[name] += 1
        v = v / 1000.0 if unit in ("ns", "nsecond") else v
    else:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    per[name][m] += v
tot = sum(p["gpu__time_duration.sum"] for p in per.values())
print('ncu --nvtx --nvtx-include "sdeo_step/" --graph-profiling node --metrics gpu__time_duration.sum,dram__bytes_read.sum,'
      "dram__bytes_write.sum --clock-control none --csv")
print("command: python tools/profile_step.py --graph --steps 2   (256x384, cond+uncond batch 2; the NVTX range holds ONE replay of "
      "the captured step graph)")
print("(cold-cache, serialised per-kernel replays: compare SHARES with bench.py's in-graph trace, not absolute times)")
print(f"kernel nodes in the step: {sum(cnt.values())}; sum of kernel times {tot:.1f} us")
for name, p in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    t = p["gpu__time_duration.sum"]
    print(f"{name:24s} {cnt[name]:4d} launches {t:9.1f} us {100 * t / tot:6.2f}%  avg {t / cnt[name]:7.2f} us   DRAM read "
          f"{p['dram__bytes_read.sum'] / 1e6:8.1f} MB  write {p['dram__bytes_write.sum'] / 1e6:7.1f} MB")
c = per["conv_gemm_kernel"]
n = cnt["conv_gemm_kernel"]
print(f"conv_gemm_kernel DRAM traffic per launch: {(c['dram__bytes_read.sum'] + c['dram__bytes_write.sum']) / n:.0f} bytes "
      f"(algorithmic: 2.442 GB of bf16 weights per step / {n} launches = {2.442e9 / n:.0f})")
