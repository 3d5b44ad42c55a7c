"""Per-kernel time of one VAE decode (CUPTI via torch.profiler): python tools/vae_breakdown.py [--batch 16] [--latent 64 64]"""
import argparse, os, sys, re, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablediffusioneo_b200 import synth
from stablediffusioneo_b200.cldm.cldm import ControlLDM
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--latent", type=int, nargs=2, default=[64, 64])
ap.add_argument("--list", default="", help="also list every launch whose kernel name contains this string, in order")
args = ap.parse_args()
dev = torch.device("cuda:0")
with torch.device(dev):
    model = ControlLDM().eval()
synth.randomize_(model)
z = torch.randn((args.batch, 4, *args.latent), device=dev)
for _ in range(4):  # the third call captures the decode graph: time (and profile) replays
    u8 = model.decode_first_stage_u8(z)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); u8 = model.decode_first_stage_u8(z); e.record(); torch.cuda.synchronize()
print(f"decode: {s.elapsed_time(e):.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    u8 = model.decode_first_stage_u8(z)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
listed = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"^void ", "", ev.name)
        name = re.sub(r"sdeo::", "", name)
        if args.list and args.list in name:
            listed.append((ev.time_range.start, name[:40], ev.device_time if hasattr(ev, "device_time") else ev.cuda_time))
        key = name[:90]
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(a[1] for a in agg.values())
print(f"sum of kernel times {tot / 1e3:.2f} ms")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{a[1] / 1e3:8.2f} ms {100 * a[1] / tot:5.1f}%  x{a[0]:3d}  {k}")
if listed:
    print(f"launches matching {args.list!r}, in start order (us):")
    for _, name, t in sorted(listed):
        print(f"  {t:9.1f}  {name}")
