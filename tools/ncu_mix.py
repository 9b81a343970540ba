"""Dynamic opcode mix of one kernel launch of an ncu report, per CUDA source line (warp instructions per tile-warp).
    python tools/ncu_mix.py gpurun_out/x.ncu-rep [tile_warps=2**21] [top=40]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
W = float(eval(sys.argv[2])) if len(sys.argv) > 2 else 2.0 ** 21
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = f = None
byline = collections.defaultdict(collections.Counter)
ops = collections.Counter()
stl = collections.Counter()
for r in rows:
    if r and r[0] == "File Path":
        f = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[0] == "Line No":
        continue
    if r[0] != "":
        cur = (f, int(r[0]), r[1][:70])
        continue
    s = r[3].strip()
    if s.startswith("@"):
        s = s.split(None, 1)[1]
    op = s.split()[0]
    try:
        n = int(r[7])
    except ValueError:
        continue
    k = op.split(".")[0] + (".MOV" if ".MOV" in op else "")
    byline[cur][k] += n
    ops[k] += n
    stl[cur] += int(r[6])
tot = sum(ops.values())
ts = sum(stl.values()) or 1
print(f"total per tile-warp {tot / W:.0f}")
print(", ".join(f"{o} {n / W:.0f}" for o, n in ops.most_common(26)))
for k, c in sorted(byline.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
    t = sum(c.values())
    print(f"{t / W:7.1f} st={stl[k] / ts:5.3f} {k[0][-12:]}:{k[1]} {k[2][:44]} | " + ", ".join(f"{o} {n / W:.0f}" for o, n in c.most_common(6)))
