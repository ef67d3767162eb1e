#!/usr/bin/env python3
"""Per-instruction view of an ncu report (--set full --import-source on): address, share of warp-stall samples, executions, top
stall reasons, source line, SASS.  Usage: ncu_sass.py report.ncu-rep [min_share_percent] [lo_hex hi_hex]
The kernel's source lines come from the report itself (ncu --page source --csv --print-source sass,cuda is not joined; the
CUDA line of each SASS instruction is taken from nvdisasm -g on the in-tree libowgpu.so, which must be the profiled build)."""
import csv, os, re, subprocess, sys, tempfile
rep = sys.argv[1]
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
lo = int(sys.argv[3], 16) if len(sys.argv) > 4 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.split("\n")))
kname = rows[0][1]
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > ix["stall_wait"] and r[0].startswith("0x")]
base = min(int(r[0], 16) for r in data)
def f(r, k):
    try: return float(r[ix[k]] or 0)
    except ValueError: return 0.0
tot = sum(f(r, "# Samples") for r in data)
# source lines from the in-tree build
mangled = "_ZN4owgd17chain_tile_kernelILb%dELb%dE" % (1 if "(bool)1, (bool)" in kname else 0, 1 if kname.count("(bool)1") == 2 or ", (bool)1>" in kname else 0)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "openwurli_b200/lib/libowgpu.so")], cwd=tmp, capture_output=True)
amap = {}
try:
    cubin = [x for x in os.listdir(tmp) if x.startswith("owgpu")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(dis) if l.startswith("//--------------------- .text." + mangled)][0]
    cur = None
    for l in dis[start + 1:]:
        if l.startswith("//--------------------- .text."): break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: cur = "%s:%s" % (m.group(1).split("/")[-1], m.group(2)); continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m: amap[int(m.group(1), 16)] = (cur, m.group(2))
except Exception as e:
    print("no line info:", e)
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
print("kernel:", kname[:80], " total samples %.0f" % tot)
cum = 0.0
for r in data:
    a = int(r[0], 16) - base
    s = f(r, "# Samples")
    if not (lo <= a <= hi): continue
    cum += s
    if 100 * s / tot < minp: continue
    st = sorted(((f(r, k), k[6:]) for k in stalls), reverse=True)[:3]
    ln, sass = amap.get(a, ("", r[1]))
    if sass.split()[:2] != r[1].split()[:2] and a in amap: ln = "?" + (ln or "")
    print("%05x %5.2f%% ex=%-9s %-34s %-26s %s" % (a, 100 * s / tot, r[ix["Instructions Executed"]], " ".join("%s=%.2f" % (k, 100 * v / tot) for v, k in st if v > 0), ln or "", r[1][:70]))
print("range share: %.2f%%" % (100 * cum / tot))
