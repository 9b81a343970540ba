"""Aggregate an ncu report's source page by CUDA source line: share of executed warp instructions and of stall samples.
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [launch_skip] [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
skip = sys.argv[2] if len(sys.argv) > 2 else "0"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, data, hdr = None, [], None
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        iex, ism = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr and len(r) > iex and r[0] != "":
        try:
            data.append((cur, int(r[0]), r[1], int(r[iex]), int(r[ism])))
        except ValueError:
            pass
tot = sum(d[3] for d in data) or 1
ts = sum(d[4] for d in data) or 1
print(f"total warp instructions {tot}, stall samples {ts}")
for d in sorted(data, key=lambda d: -d[3])[:top]:
    print(f"{d[3] / tot:6.3f} stall={d[4] / ts:6.3f} {d[0]}:{d[1]:<4d} {d[2][:100]}")
