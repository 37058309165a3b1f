"""per-kernel SASS instruction histogram of libmgpu.so (cuobjdump -sass): which memory / sync / math instructions the sm_100a
kernels are made of. usage: python scripts/sass_hist.py > profiles/r02_sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "manticoresearch_b200", "libmgpu.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
cur_arch, kernel_archs = None, set()
for line in out.split("\n"):
    a = re.search(r"arch = (\S+)", line)
    if a:
        cur_arch = a.group(1)
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void mgpu::", "").replace("mgpu::", "")
        hist[kern] = collections.Counter()
        kernel_archs.add(cur_arch)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
print("# SASS instruction histogram of `libmgpu.so` (kernels built for %s), `cuobjdump -sass`\n" % ", ".join(sorted(x for x in kernel_archs if x)))
print("Hand-written integer / byte kernels: global loads are 128-bit (`LDG.E.128`) or 32-bit coalesced words of the presence bitmaps,")
print("warp primitives are `SHFL` / `VOTE` / `MATCH` / `REDUX`, no tensor-core (`UTCMMA` / `HMMA`) and no bulk-async (`UTMALDG` / `SYNCS`) instruction by design")
print("(DESIGN.md section 3: the units are 140-byte doclist blocks and 128-byte bitmap lines consumed once).\n")
groups = [("LDG.E.128", r"^LDG\.E\.128"), ("LDG other", r"^LDG(?!\.E\.128)"), ("STG", r"^STG"), ("LDS/STS", r"^(LDS|STS)"), ("LDL/STL (local)", r"^(LDL|STL)"),
          ("ATOM/RED (global)", r"^(ATOMG|RED|ATOM)\b|^ATOMG|^RED"), ("ATOMS (shared)", r"^ATOMS"), ("SHFL", r"^SHFL"), ("VOTE/MATCH/REDUX", r"^(VOTE|VOTEU|MATCH|REDUX)"),
          ("BAR", r"^BAR"), ("POPC/FLO/BREV", r"^(POPC|FLO|BREV)"), ("LOP3/SHF/PRMT", r"^(LOP3|SHF|PRMT|LOP)"), ("IMAD/IADD3", r"^(IMAD|IADD3|IADD)"),
          ("FP32 (FADD/FMUL/MUFU)", r"^(FADD|FMUL|FFMA|MUFU|F2I|I2F|FSETP)"), ("CCTL/prefetch", r"^CCTL"), ("UTMALDG/SYNCS/UTCMMA", r"^(UTMALDG|UTMASTG|SYNCS|UTCMMA|HMMA|QMMA)")]
print("| kernel | instructions | " + " | ".join(g for g, _ in groups) + " |")
print("|---|---|" + "---|" * len(groups))
for k, h in hist.items():
    tot = sum(h.values())
    if tot < 50:
        continue
    row = [sum(v for n, v in h.items() if re.search(p, n)) for _, p in groups]
    print("| `%s` | %d | %s |" % (k, tot, " | ".join(str(x) for x in row)))
