"""Per-source-line totals of an ncu report (needs -lineinfo + --import-source on): warp instructions
executed and stall samples of the top N source lines.  Usage: python tools/ncu_hot_lines.py report.ncu-rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
lines = []
fpath, hdr = "?", None
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 8 or r[0] == "":
        continue                                  # SASS rows under a source line: already summed in the line's row
    try:
        lines.append((float(r[7] or 0), float(r[6] or 0), fpath, r[0], r[1].strip()))
    except ValueError:
        pass
ti, ts = sum(x[0] for x in lines), sum(x[1] for x in lines)
print(f"total warp instructions {ti:.0f}, stall samples {ts:.0f}")
for inst, samp, f, ln, src in sorted(lines, key=lambda x: -x[0])[:top]:
    print(f"{inst:11.0f} inst {100 * inst / ti:5.1f}% {samp:7.0f} samp {100 * samp / max(ts, 1):5.1f}%  {f}:{ln}  {src[:110]}")
