"""Per-source-line stall-sample table from an ncu report captured with --import-source on:
python tools/ncu_lines.py report.ncu-rep [top]. Uses `ncu --page source --print-source cuda,sass --csv`."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and len(r) >= 8:
        i_s = hdr.index("# Samples")
        i_ni = hdr.index("Warp Stall Sampling (Not-issued Samples)")
        stalls = {}
        for j, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h and j < len(r):
                try:
                    v = int(r[j])
                except ValueError:
                    v = 0
                if v:
                    stalls[h[6:]] = v
        try:
            lines.append((int(r[i_s]), fname, int(r[0]), r[1].strip(), stalls))
        except ValueError:
            pass
tot = sum(l[0] for l in lines)
print(f"total samples {tot}")
for s, f, ln, src, st in sorted(lines, key=lambda l: -l[0])[:top]:
    ss = " ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{s:6d} {100.0 * s / max(tot, 1):5.1f}% {f}:{ln:<5d} {src[:90]:90s} {ss}")
