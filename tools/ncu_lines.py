#!/usr/bin/env python3
"""Join an ncu report's per-SASS-instruction samples with source lines (nvdisasm -gi of the in-tree cubin) and
print the hottest source lines with their stall mix.

    python tools/ncu_lines.py gpurun_out/k1.ncu-rep k1_decode_kernel libcimbar_b200/csrc/k1_decode.cu [top_n]

Needs the libcb200.so that was profiled (same SASS) to be the one in the tree."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kernel, srcfile = sys.argv[1], sys.argv[2], sys.argv[3]
# kernel may be "ncu_regex:mangled_substring" to pick one template instance in the cubin
mangled = kernel.split(":", 1)[1] if ":" in kernel else kernel
kernel = kernel.split(":", 1)[0]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "libcimbar_b200", "lib", "libcb200.so")

# ---- SASS rows from the report
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

# ---- line info from the cubin
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
lines_of = []   # per instruction: (line, opcode text)
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    m = re.search(r"^\.text\.\S*" + mangled + r"\S*:\n", dis, re.M)
    if not m:
        continue
    cur = None
    for ln in dis[m.end():].split("\n"):
        if ln.startswith("//---") or ln.startswith("\t.section"):
            break
        mm = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if mm:
            # innermost line in our file, else the inlining site
            cur = int(mm.group(2)) if mm.group(1).endswith(os.path.basename(srcfile)) else (int(mm.group(4)) if mm.group(4) else cur)
            continue
        mi = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", ln)
        if mi:
            lines_of.append((cur, mi.group(1).strip()))
    break
if len(lines_of) != len(data):
    print(f"warning: {len(lines_of)} disassembled instructions vs {len(data)} profiled rows (different build?)", file=sys.stderr)

src = open(os.path.join(ROOT, srcfile)).read().split("\n")
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot_s = tot_i = 0
for i, r in enumerate(data):
    line = lines_of[i][0] if i < len(lines_of) else None
    s_, n_ = int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]])
    a = agg[line]
    a[0] += s_; a[1] += n_
    for c in stall_cols:
        a[2][c] += int(r[ix[c]] or 0)
    tot_s += s_; tot_i += n_
print(f"kernel {kernel}: {len(data)} SASS instrs, samples {tot_s}, warp-instr {tot_i}")
allst = collections.Counter()
for a in agg.values():
    allst.update(a[2])
ts = sum(allst.values())
print("stall mix: " + ", ".join(f"{k[6:]} {100 * v / ts:.1f}%" for k, v in allst.most_common(10)))
print(f"{'line':>5} {'samp%':>6} {'inst%':>6}  stall mix / source")
for line, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    mix = ", ".join(f"{k[6:]}:{100 * v / max(1, a[0]):.0f}" for k, v in a[2].most_common(3))
    text = src[line - 1].strip()[:90] if line and line <= len(src) else "?"
    print(f"{str(line):>5} {100 * a[0] / tot_s:6.2f} {100 * a[1] / tot_i:6.2f}  [{mix}]  {text}")
