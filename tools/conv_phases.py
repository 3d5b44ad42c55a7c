"""Per-CTA phase timing of conv_gemm_kernel via the SDEO_CONV_DEBUG clock64 slots. python tools/conv_phases.py [SUBSTR]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stablediffusioneo_b200 import ops  # noqa: E402

BF = torch.bfloat16
dev = torch.device("cuda:0")
CASES = {
    "out": (2, 32, 48, 320, 4, 3, "plain32"),
    "conv1": (2, 32, 48, 320, 320, 1, "stream"),
    "conv3": (2, 32, 48, 320, 320, 3, "stream"),
    "mid": (2, 8, 12, 1280, 1280, 3, "stream"),
    "c1_plain": (2, 32, 48, 320, 320, 1, "plain"),
    "c1_res16": (2, 32, 48, 320, 320, 1, "res16"),
    "c1_f32": (2, 32, 48, 320, 320, 1, "plain32"),
    "c1_f32res": (2, 32, 48, 320, 320, 1, "f32res"),
    "conv3_plain": (2, 32, 48, 320, 320, 3, "plain"),
    "conv3_64": (2, 64, 64, 320, 320, 3, "stream"),
    "conv3_96": (8, 96, 96, 320, 320, 3, "stream"),
    "vae128": (4, 512, 512, 128, 128, 3, "plain"),
    # folded GroupNorm (+SiLU) in the operand path: "gnf" = ResBlock conv2 (residual, stream out), "gnfh" = conv1 (bf16 out)
    "gnf3": (2, 32, 48, 320, 320, 3, "gnf"),
    "gnf3h": (2, 32, 48, 320, 320, 3, "gnfh"),
    "gnf1": (2, 32, 48, 320, 320, 1, "gnf1"),
    "gnf_l2": (2, 16, 24, 640, 640, 3, "gnf"),
    "gnf_mid": (2, 8, 12, 1280, 1280, 3, "gnf"),
    "gnf_deep": (2, 4, 6, 1280, 1280, 3, "gnf"),
    "gnf_vae": (4, 512, 512, 128, 128, 3, "gnfv"),
}
names = [a for a in sys.argv[1:] if "=" not in a] or list(CASES)
# optional plan overrides: combos=00,01,10,11 (PAIR HALO digits) bn=160 splits=1
opts = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
combos = opts.get("combos", "").split(",") if opts.get("combos") else [None]
if "bn" in opts:
    os.environ["SDEO_FORCE_BN"] = opts["bn"]
if "splits" in opts:
    os.environ["SDEO_FORCE_SPLITS"] = opts["splits"]
for name, combo in [(n_, c_) for n_ in names for c_ in combos]:
    if combo:
        os.environ["SDEO_PAIR"], os.environ["SDEO_HALO"] = combo[0], combo[1]
        print(f"#### {name} PAIR={combo[0]} HALO={combo[1]}")
    n, h, w, cin, cout, k, mode = CASES[name]
    wt = torch.randn((cout, cin, k, k), device=dev) / math.sqrt(cin * k * k)
    x = torch.randn((n, h, w, cin), device=dev).to(BF)
    pw = ops.pack_conv_weight(wt)
    kw = dict(bias=torch.randn((cout,), device=dev))
    if mode == "stream":
        kw.update(residual=torch.randn((n, h, w, cout), device=dev), out_fp32=True, twin=True)
    elif mode == "f32res":
        kw.update(residual=torch.randn((n, h, w, cout), device=dev), out_fp32=True)
    elif mode == "res16":
        kw.update(residual=torch.randn((n, h, w, cout), device=dev).to(BF))
    elif mode == "plain32":
        kw.update(out_fp32=True)
    if mode.startswith("gnf"):
        # producer: 3x3 conv that leaves statistics (bf16 output = the raw operand)
        wp = torch.randn((cin, cin, 3, 3), device=dev) / math.sqrt(cin * 9)
        x = ops.conv2d(x, ops.pack_conv_weight(wp), gn_stats=True)
        st = ops.gn_stats_fold(x._gn_stats, n, cin)
        gnf = ops.GnFold(st, None, torch.ones(cin, device=dev), torch.zeros(cin, device=dev), 32, 1e-5, mode != "gnf1")
        kw.update(gnf=gnf)
        if mode in ("gnf", "gnf1"):
            kw.update(residual=torch.randn((n, h, w, cout), device=dev), out_fp32=True, twin=True, gn_stats=(mode == "gnf"))
        elif mode == "gnfh":
            kw.update(gn_stats=True)
        elif mode == "gnfv":
            kw.update(gn_stats=True)
        if opts.get("nofold"):   # the same conv without the fold (on the raw operand: timing reference only)
            del kw["gnf"]
    dbg = torch.zeros((40000, 16), dtype=torch.int64, device=dev)
    try:
        for _ in range(3):
            ops.conv2d(x, pw, **kw)
    except Exception as ex:
        print(f"   not available: {ex}")
        continue
    torch.cuda.synchronize()
    if opts.get("cold"):   # weights (and everything else) out of L2: inside the step every layer's weights come from HBM
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        flush.fill_(1)
        torch.cuda.synchronize()
    os.environ["SDEO_CONV_DEBUG"] = hex(dbg.data_ptr())
    ops.conv2d(x, pw, **kw)
    torch.cuda.synchronize()
    del os.environ["SDEO_CONV_DEBUG"]
    d = dbg.cpu()
    used = d[d[:, 0] != 0]
    rel = (used - used[:, :1]).float()
    lab = ["start", "prologue", "first data", "mma issued", "acc ready", "phase1", "barrier", "phase2", "end"]
    print(f"== {name}: {used.shape[0]} CTAs; cycles since CTA start (median / max)")
    for i in range(1, 9):
        col = rel[:, i]
        print(f"   {lab[i]:11s} {col.median().item():9.0f} {col.max().item():9.0f}")
    if used[:, 13].max() > 0 and os.environ.get("SDEO_PHASE2_DETAIL"):
        print("   phase 2 detail (thread 64, first item batch): after residual wait %.0f | gather done %.0f | tile loads issued %.0f | item 0 "
              "stored %.0f | batch done %.0f" % tuple(rel[:, k].median().item() for k in (13, 9, 10, 11, 12)))
    if used[:, 13].max() > 0:
        print(f"   phase 2: residual tile waited for until {rel[:, 13].median().item():.0f}; first item batch done at {rel[:, 14].median().item():.0f}")
    if used[:, 15].max() > 0:
        print(f"   folded GroupNorm: table ready at {rel[:, 2].median().item():.0f}; cycles spent normalising tiles (thread 64) {used[:, 15].float().median().item():.0f}")
    if used[:, 9].max() > 0 or used[:, 10].max() > 0:
        lead = used[used[:, 9] > 0]
        print(f"   MMA warp: cycles waiting for operand data (full barriers) {lead[:, 9].float().median().item():.0f}; "
              f"producer: cycles waiting for free stages (empty barriers) {used[:, 10].float().median().item():.0f}; "
              f"producer issued its last load at {(used[:, 11] - used[:, 0]).float().median().item():.0f}; stages {int(used[0, 12])}")
    if False:
        for a, b, c, nm in ((9, 10, 11, "chunk 8"), (12, 13, 14, "chunk 16")):
            print(f"   {nm}: at {rel[:, a].median().item():.0f}; wait-for-data {(used[:, b] - used[:, a]).float().median().item():.0f}"
                  f" cycles; issue 4 MMAs + commit {(used[:, c] - used[:, b]).float().median().item():.0f} cycles")
