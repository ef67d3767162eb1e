#!/usr/bin/env python3
"""Static SASS statistics of one kernel of libowgpu.so, attributed to CUDA source lines (nvdisasm -g):
instruction counts per source line / per enclosing function, opcode classes, and every local-memory (spill) access.
Usage: sass_lines.py kernel_mangled_prefix [lo_line hi_line file]  -- the optional range restricts the per-class summary."""
import collections, os, re, subprocess, sys, tempfile
kprefix = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "openwurli_b200/lib/libowgpu.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith("owgpu")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith("//--------------------- .text." + kprefix)][0]
cur = None
ins = []
for l in dis[start + 1:]:
    if l.startswith("//--------------------- .text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), cur, m.group(2)))
def cls(op):
    o = op.split()[0] if not op.startswith("@") else op.split()[1]
    o = o.split(".")[0]
    if o in ("DADD", "DMUL", "DFMA", "DSETP"): return "fp64"
    if o in ("FSEL", "SEL"): return "select"
    if o in ("LDS", "STS", "LDSM"): return "smem"
    if o in ("LDL", "STL"): return "local"
    if o in ("LDG", "STG", "LDC", "ULDC", "LD", "ST"): return "gmem/const"
    if o in ("MOV", "IMAD", "UMOV", "PRMT"): return "mov/imad"
    if o in ("SHFL", "VOTE", "WARPSYNC", "BAR", "SYNCS", "VOTEU"): return "warp"
    if o in ("BRA", "BSSY", "BSYNC", "CALL", "RET", "EXIT", "BREAK", "WARPSYNC"): return "branch"
    if o in ("MUFU",): return "mufu"
    if o in ("ISETP", "PLOP3", "LOP3", "IADD3", "IADD", "SHF", "LEA", "VIADD", "UIADD3", "ULOP3", "USHF", "ULEA", "UISETP", "R2UR", "S2R", "FSETP", "FFMA", "FADD", "FMUL", "I2F", "F2I", "VIMNMX", "IMNMX", "UPLOP3", "USEL", "UIMAD", "R2P", "P2R", "POPC", "FLO", "BREV", "I2FP", "F2FP", "UFLO", "UPOPC", "LOP", "IABS", "ISCADD", "CS2R", "NOP", "UMNMX", "UVIADD", "UVIMNMX"): return "int/pred/fp32"
    return "other:" + o
lo = hi = None
if len(sys.argv) > 4:
    lo, hi, fname = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
per_line = collections.Counter(); per_cls = collections.Counter()
for a, c, op in ins:
    per_line[c] += 1
    if lo is None or (c and c[0] == fname and lo <= c[1] <= hi):
        per_cls[cls(op)] += 1
print("total instructions", len(ins))
print("classes%s:" % ("" if lo is None else f" in {fname}:{lo}-{hi}"), dict(per_cls.most_common()), "sum", sum(per_cls.values()))
print("local-memory accesses:")
for a, c, op in ins:
    if re.search(r"\b(LDL|STL)\b", op): print("  %05x %s %s" % (a, c, op))
if "--lines" in sys.argv:
    for c, n in sorted(per_line.items(), key=lambda kv: (str(kv[0][0]) if kv[0] else "", kv[0][1] if kv[0] else 0)):
        print(c, n)
