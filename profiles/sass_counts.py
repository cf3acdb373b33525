"""SASS mnemonic counts per kernel of libmsa_b200.so (cuobjdump -sass | c++filt), kernels with tensor-core / TMA / TMEM / cp.async
instructions only.    python profiles/sass_counts.py > profiles/r02_sass_counts.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "metaspeakeradaptation-tts_b200", "libmsa_b200.so")
sass = subprocess.run(f"cuobjdump -sass {LIB} | c++filt", shell=True, capture_output=True, text=True).stdout
cols = ["HMMA", "UTCHMMA", "LDTM", "UTMALDG", "LDGSTS"]
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (.*)", line)
    if m:
        name = m.group(1).replace("(anonymous namespace)::", "")
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("msa::", "")
        cur = counts.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1).split(".")[0]
        cur["instr"] += 1
        if op in cols:
            cur[op] += 1
print("SASS mnemonic counts per kernel of libmsa_b200.so (cuobjdump -sass; only kernels with tensor-core / TMA / TMEM / cp.async instructions).")
print("HMMA = mma.sync (bf16x3 products of the grouped recurrences, 3xTF32 of the inference step), UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld,")
print("UTMALDG = TMA tensor loads, LDGSTS = cp.async (k_attn_*_mma<G, true>: the streamed weight fragments of the per-task-weight variants).")
print(f"{'kernel':54s} {'instr':>6s} " + " ".join(f"{c:>8s}" for c in cols))
for name, c in counts.items():
    if any(c[k] for k in cols):
        print(f"{name[:54]:54s} {c['instr']:6d} " + " ".join(f"{c[k]:8d}" for k in cols))
