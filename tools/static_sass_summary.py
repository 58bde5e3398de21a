"""Static summary of the built library (no GPU needed): per kernel the resource usage cuobjdump reports (registers, shared memory,
local-memory stack) and SASS instruction counts that matter on this path -- DMMA (FP64 tensor pipe), LDGSTS (cp.async), LDL / STL
(spills), BAR, SHFL, DFMA / DADD / DMUL, MUFU.

    python tools/static_sass_summary.py [yaqs_b200/libyaqs_b200.so] > profiles/r02_static_sass_final_build.txt
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "yaqs_b200/libyaqs_b200.so"
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True, check=True).stdout
usage = {}
name = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        name = m.group(1)
        continue
    if name and "REG:" in line:
        usage[name] = {k: int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL|CONSTANT\[0\]):(\d+)", line)}
        name = None
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
counts = collections.defaultdict(collections.Counter)
cur = None
WATCH = ("DMMA", "LDGSTS", "LDL", "STL", "BAR", "SHFL", "DFMA", "DADD", "DMUL", "MUFU", "LDG", "STG", "LDS", "STS", "UTMALDG", "ATOMG", "RED")
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur:
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    counts[cur][w] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
def _short(d: str) -> str:
    d = re.sub(r"^void ", "", d).replace("(anonymous namespace)::", "")
    depth, out = 0, []
    for ch in d:                                     # keep template arguments, drop the parameter list
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            break
        out.append(ch)
    return "".join(out).replace("(anonymous namespace)::", "")


short = {n: _short(d) for n, d in zip(counts, demangle)}
arch = re.findall(r"arch = (\S+)", sass)
print(f"# {lib}: {len(counts)} kernels, arch {sorted(set(arch))}")
cols = ["REG", "SHARED", "STACK", "total", "DMMA", "LDGSTS", "LDL", "STL", "DFMA", "BAR", "SHFL", "LDG", "STG", "LDS", "STS", "MUFU", "UTMALDG"]
print("kernel".ljust(72) + "".join(c.rjust(8) for c in cols))
for n in sorted(counts, key=lambda k: short[k]):
    u = usage.get(n, {})
    row = [u.get("REG", ""), u.get("SHARED", ""), u.get("STACK", "")] + [counts[n][c] for c in cols[3:]]
    print(short[n][:71].ljust(72) + "".join(str(v).rjust(8) for v in row))
