"""Static SASS instruction counts per source line, per kernel, from an object built with -lineinfo (no GPU needed).

    cuobjdump -xelf all spinoza_b200/lib/obj/kernels_tile.cu.o && nvdisasm -g -c kernels_tile.sm_100a.cubin > tile.sass
    python tools/sass_lines.py tile.sass [kernel-name-substring]

Prints the total per kernel, the lines that carry local-memory (spill) instructions, and -- with a kernel substring -- the
histogram file:line -> instructions.  Used for profiles/round1_summary.md section 10.
"""
import collections
import re
import sys

path = sys.argv[1]
cur = kern = None
spill = collections.defaultdict(collections.Counter)
total = collections.Counter()
lines = collections.defaultdict(collections.Counter)
ops = collections.defaultdict(collections.Counter)
for line in open(path):
    m = re.match(r"\.text\.(\S+):", line)
    if m:
        kern = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        total[kern] += 1
        lines[kern][cur] += 1
        ops[kern][m.group(1).split(".")[0]] += 1
        if "STL" in line or "LDL" in line:
            spill[kern][cur] += 1
for k, n in total.items():
    top = ", ".join(f"{o} {c}" for o, c in ops[k].most_common(8))
    print(f"{n:6d}  {k}\n        {top}")
    if spill[k]:
        print("        spill sites:", sorted(spill[k].items()))
if len(sys.argv) > 2:
    k = [x for x in lines if sys.argv[2] in x][0]
    for key, n in sorted(lines[k].items()):
        print(key, n)
