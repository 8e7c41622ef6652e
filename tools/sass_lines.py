"""Join an ncu source-page SASS csv (per-instruction 'Instructions Executed') with nvdisasm's line info of the same kernel:
warp-instructions executed per source line of the .cu file.
usage: python tools/sass_lines.py <cubin> <kernel substring> <ncu sass csv> [top N]"""
import csv, re, subprocess, sys, collections

cubin, kname, path = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# split per function
lines_of = []   # line number per instruction in order, for the wanted function
cur = None; infn = False; line = 0
for l in txt.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+),', l)
    if m:
        infn = kname in m.group(1) and not lines_of
        continue
    if not infn:
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        line = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        lines_of.append(line)
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; ie = h.index("Instructions Executed"); isrc = h.index("Source")
ins = rows[hi + 1:]
assert len(ins) == len(lines_of), (len(ins), len(lines_of))
per = collections.Counter(); tot = 0; perop = collections.Counter()
for r, ln in zip(ins, lines_of):
    n = int(r[ie]); per[ln] += n; tot += n; perop[r[isrc].split()[0] if not r[isrc].strip().startswith('@') else r[isrc].split()[1]] += n
import os
srcname = os.environ.get("SRC", "orb.cu")
src = open("/root/repo/airdos_b200/csrc/" + srcname).read().splitlines()
print("total warp-instructions", tot)
for ln, n in per.most_common(top):
    fn, k = ln if ln else ("?", 0)
    print(f"{n:>12} {100*n/tot:5.1f}%  {fn}:{k:5d}: {src[k-1].strip()[:120] if fn == srcname and 0 < k <= len(src) else ''}")
print("by opcode:")
for op, n in perop.most_common(25):
    print(f"{n:>12} {100*n/tot:5.1f}%  {op}")
