"""Per-SOURCE-LINE shares of executed instructions and stall samples of one kernel of an .ncu-rep (no GPU needed).

ncu's source page lists SASS addresses; the CUDA line of each address comes from the object file the library was built
from (`cuobjdump -xelf` + `nvdisasm -g`), matched by the offset inside the kernel's function — so the object must hold
the same code for that kernel as the profiled run did.

    python tools/ncu_lines.py gpurun_out/r2_prof.ncu-rep gather_forward_planar streammos_b200/lib/bilinear_gather.o [--nth 2] [--top 30]
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ap = argparse.ArgumentParser()
ap.add_argument("report")
ap.add_argument("kernel", help="substring of the (demangled) kernel name")
ap.add_argument("object", help=".o the kernel was compiled into")
ap.add_argument("--nth", type=int, default=0, help="which launch of the matching kernels (0 = first)")
ap.add_argument("--top", type=int, default=30)
a = ap.parse_args()

# ---- the kernel's source page (the CSV holds one section per profiled launch) ------------------------------------------
src = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
sections, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        sections.append(cur)
    elif cur is not None and "Address" in r and "Source" in r:
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
match = [x for x in sections if a.kernel in x["name"]]
if not match:
    sys.exit("no kernel matches %r" % a.kernel)
sec = match[min(a.nth, len(match) - 1)]
name, hdr, data = sec["name"], sec["hdr"], sec["rows"]
kid = "%d of %d matching" % (min(a.nth, len(match) - 1), len(match))
plain = name.replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
mangled_hint = re.sub(r"[^A-Za-z0-9_]", " ", plain.split("(")[0].split("<")[0]).split()[-1]  # e.g. gather_forward_planar_kernel
col = {h: i for i, h in enumerate(hdr)}
base = int(data[0][col["Address"]], 16)
sass_of = {int(r[col["Address"]], 16) - base: r[col["Source"]].split()[0] for r in data}

# ---- offset -> source line from the object file -------------------------------------------------------------------
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(a.object)], cwd=tmp, capture_output=True)
best = None
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout.splitlines()
    # functions of this cubin whose name holds the kernel's identifier: pick the one whose opcodes match the report
    starts = [i for i, l in enumerate(dis) if l.endswith(":") and mangled_hint in l and not l.startswith("\t") and ".text" not in l]
    for s in starts:
        cur, m = None, {}
        for l in dis[s + 1:]:
            g = re.search(r'//## File "(.*)", line (\d+)', l)
            if g:
                cur = (os.path.basename(g.group(1)), int(g.group(2)))
                continue
            g = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if g:
                m[int(g.group(1), 16)] = (cur, g.group(2).split()[0] if not g.group(2).startswith("@") else g.group(2).split()[1])
            elif l.strip().startswith(".L_x_") and l.strip().endswith(":") and m and max(m) >= max(sass_of):
                break
        agree = sum(1 for o, op in sass_of.items() if o in m and (m[o][1] == op or op.startswith("@")))
        if best is None or agree > best[0]:
            best = (agree, m)
agree, off2line = best
print("%s\n  launch id %s, %d SASS instructions, %d matched by opcode against %s" % (name[:110], kid, len(sass_of), agree, a.object))
inst, samp = collections.Counter(), collections.Counter()
for r in data:
    off = int(r[col["Address"]], 16) - base
    ln = off2line.get(off, (("?", 0), ""))[0] or ("?", 0)
    inst[ln] += int(r[col["Instructions Executed"]] or 0)
    samp[ln] += int(r[col["# Samples"]] or 0)
ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
print("  %d warp instructions, %d stall samples" % (ti, ts))
cache = {}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for (f, l), n in sorted(inst.items(), key=lambda x: -(x[1] / ti + samp[x[0]] / ts))[: a.top]:
    if f not in cache:
        p = os.path.join(root, "streammos_b200", "csrc", f)
        cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = cache[f][l - 1].strip()[:84] if 0 < l <= len(cache[f]) else ""
    print("  %-22s %4d  inst %5.1f %%  stalls %5.1f %%  %s" % (f, l, 100 * n / ti, 100 * samp[(f, l)] / ts, text))
