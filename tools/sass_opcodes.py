"""Per-kernel SASS opcode histogram of airdos_b200/lib/libairdos_b200.so -> profiles/<tag>_sass_opcodes.csv.
Runs in the build container (cuobjdump -sass needs no GPU).  usage: python tools/sass_opcodes.py r2"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["UTMALDG", "UTMASTG", "UCGABAR_ARV", "DMMA", "DFMA", "DMUL", "DADD", "MUFU", "VIMNMX3", "VIMNMX", "IDP", "POPC", "SHFL", "REDG", "ATOMG", "ATOMS", "LDG", "STG",
        "LDS", "STS", "LDL", "STL", "BAR"]


def main(tag):
    so = os.path.join(ROOT, "airdos_b200", "lib", "libairdos_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", name).replace("adb::", "").replace("void ", "").replace("(anonymous namespace)::", "")
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            op, mods = m.group(1), m.group(2)
            per[cur]["total"] += 1
            key = op
            if op == "VIMNMX3" or (op == "VIMNMX" and False):
                key = "VIMNMX3"
            if key in COLS:
                per[cur][key] += 1
    out = os.path.join(ROOT, "profiles", f"{tag}_sass_opcodes.csv")
    with open(out, "w") as f:
        f.write(f"# {tag}: cuobjdump -sass airdos_b200/lib/libairdos_b200.so (sm_100a), static instruction counts per kernel\n")
        f.write("kernel,total," + ",".join(COLS) + "\n")
        for k, c in per.items():
            f.write(f"\"{k}\",{c['total']}," + ",".join(str(c[x]) for x in COLS) + "\n")
        tot = collections.Counter()
        for c in per.values():
            tot.update(c)
        f.write("\"ALL KERNELS\"," + str(tot["total"]) + "," + ",".join(str(tot[x]) for x in COLS) + "\n")
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r2")
