#!/usr/bin/env python3
"""Attribute an ncu report's per-instruction samples to CUDA source lines (join of `ncu --page source --csv` with
`nvdisasm -g`).  Usage: ncu_lines.py report.ncu-rep kernel_mangled_prefix [top_n]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, kprefix = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "openwurli_b200/lib/libowgpu.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith("owgpu")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith("//--------------------- .text." + kprefix)][0]
cur, amap = None, {}
for l in dis[start + 1:]:
    if l.startswith("//--------------------- .text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        amap[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.split("\n")))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > ix["# Samples"] and r[ix["Address"]].startswith("0x")]
base = min(int(r[ix["Address"]], 16) for r in data)
agg, inst = collections.Counter(), collections.Counter()
tot = tinst = 0.0
for r in data:
    a = int(r[ix["Address"]], 16) - base
    s, ie = float(r[ix["# Samples"]] or 0), float(r[ix["Instructions Executed"]] or 0)
    k = amap.get(a, (("?", 0), ""))[0] or ("?", 0)
    agg[k] += s; inst[k] += ie; tot += s; tinst += ie
src = {}
for f in os.listdir(os.path.join(ROOT, "openwurli_b200/csrc")):
    if f.endswith((".cuh", ".cu")):
        src[f] = open(os.path.join(ROOT, "openwurli_b200/csrc", f)).read().split("\n")
print(f"total samples {tot:.0f}, warp instructions {tinst:.3e}")
cum = 0
for (f, ln), s in agg.most_common(topn):
    text = src[f][ln - 1].strip()[:100] if f in src and 0 < ln <= len(src[f]) else ""
    cum += s
    print(f"{100*s/tot:5.2f}% (cum {100*cum/tot:5.1f}%) instr {100*inst[(f,ln)]/tinst:5.2f}%  {f}:{ln}  {text}")

# ---- per-function (stage) aggregation: the enclosing __device__/__global__ function of every source line
def enclosing(f, ln):
    if f not in src:
        return f
    for i in range(min(ln, len(src[f])) - 1, -1, -1):
        m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__device__|__global__|__launch_bounds__).*?([A-Za-z_0-9]+)\s*\(", src[f][i])
        if m and not src[f][i].strip().startswith("//"):
            return f"{f}:{m.group(1)}"
    return f
fagg, finst = collections.Counter(), collections.Counter()
for (f, ln), s in agg.items():
    k = enclosing(f, ln)
    fagg[k] += s; finst[k] += inst[(f, ln)]
print("\nper function (samples = time share, instr = executed warp-instruction share)")
for k, s in fagg.most_common(30):
    print(f"{100*s/tot:5.2f}%  instr {100*finst[k]/tinst:5.2f}%  {k}")
