"""Aggregate the per-instruction warp-stall samples of an ncu report by CUDA source line.

The ncu CLI prints the source page per SASS instruction only; this joins it with `nvdisasm -g` line info of the
cubin extracted from the built library (same instruction order).  Usage:
  cuobjdump -xelf all uxsim_b200/libuxsim_b200.so && nvdisasm -g -c engine.sm_100a.cubin > dis.txt
  ncu -i rep.ncu-rep --page source --csv --kernel-name regex:k_node > src.csv
  python profiles/tools/ncu_lines.py dis.txt src.csv _Z6k_nodeILi32ELi16EEvPK8DevWorldi uxsim_b200/csrc/ux_kernels_step.cuh [top]
"""
import collections
import csv
import re
import sys

dis_path, csv_path, symbol, src_path = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
dis = open(dis_path).read().split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + symbol + ":"))
seq, cur = [], None
for l in dis[start + 1:]:
    if l.startswith("//--------------------- .text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append(cur)
rows = list(csv.reader(open(csv_path)))
blocks, hdr = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        blocks.append([])
    elif r and r[0] == "Address":
        hdr = r
    elif blocks and r and r[0].startswith("0x"):
        blocks[-1].append(r)
iS, iI, iL = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("stall_long_sb")
agg, aggI, aggL, tot = collections.Counter(), collections.Counter(), collections.Counter(), 0
for b in blocks:
    assert len(b) == len(seq), (len(b), len(seq))
    for ln, r in zip(seq, b):
        s = int(r[iS] or 0)
        agg[ln] += s; aggI[ln] += int(r[iI] or 0); aggL[ln] += int(r[iL] or 0); tot += s
src = open(src_path).read().split("\n")
name = src_path.split("/")[-1]
print(f"{len(blocks)} launch(es), {len(seq)} SASS instructions, {tot} samples")
for ln, s in agg.most_common(top):
    txt = src[ln[1] - 1].strip()[:100] if ln and ln[0] == name else str(ln)
    print(f"{s:7d} {100 * s / tot:5.1f}%  long_sb {aggL[ln]:6d}  inst {aggI[ln]:9d}  L{ln[1] if ln else 0}: {txt}")
if len(sys.argv) > 6:  # optional: comma-separated line boundaries -> samples per range
    bounds = [int(x) for x in sys.argv[6].split(",")]
    rng = collections.Counter()
    for ln, s in agg.items():
        if ln and ln[0] == name:
            k = sum(1 for b in bounds if ln[1] >= b)
            rng[k] += s
        else:
            rng[-1] += s
    print("samples per line range:", {("other" if k < 0 else f"{([0] + bounds)[k]}+"): f"{100 * v / tot:.1f}%" for k, v in sorted(rng.items())})
