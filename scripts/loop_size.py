#!/usr/bin/env python
"""SASS size of the dequant loop of one tcgen05 GEMV instantiation (tuning helper, no GPU needed):
the innermost backward-branch loop that contains the full-chunk STTM sequence, with an opcode histogram."""
import collections
import re
import subprocess
import sys

LIB = "any4_b200/lib/libtinygemm_b200.so"
pat = sys.argv[1] if len(sys.argv) > 1 else "gemv_w4_tc_kernelIL8tg_dtype0ELi4ELi4ELb1ELb0ELb1E"
out = subprocess.run(["cuobjdump", "-sass", "-fun", next(
    l.split()[-1] for l in subprocess.run(["cuobjdump", "-elf", LIB], capture_output=True, text=True).stdout.splitlines()
    if pat in l and "FUNC" in l)] + [LIB], capture_output=True, text=True).stdout if False else None
# cuobjdump -fun wants the mangled name: find it in the symbol table
syms = subprocess.run(["cuobjdump", "-elf", LIB], capture_output=True, text=True).stdout
name = None
for m in re.finditer(r"(_ZN2tg2tc\w+)", syms):
    if pat in m.group(1) and "peer" not in m.group(1):
        name = m.group(1)
        break
sass = subprocess.run(["cuobjdump", "-sass", "-fun", name, LIB], capture_output=True, text=True).stdout
ins = []
for line in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_idx = {a: i for i, (a, _) in enumerate(ins)}
best = None
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_idx:
            j = addr_idx[tgt]
            body = ins[j:i + 1]
            n_st = sum("STTM" in x for _, x in body)
            if n_st >= 8 and (best is None or len(body) < len(best)):
                best = body
print(name)
print("function instructions:", len(ins))
if best:
    ops = collections.Counter()
    for _, t in best:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        ops[t.split()[0].split(".")[0]] += 1
    print("dequant loop instructions (static, incl. the k-tail path if inside):", len(best))
    print(dict(ops.most_common()))
