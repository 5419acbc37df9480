"""Instruction mix of the loops (backward branches) of one kernel's SASS:
    cuobjdump -sass -fun <mangled> lib.so | python scripts/sass_loops.py [min_len]"""
import re
import sys
from collections import Counter

min_len = int(sys.argv[1]) if len(sys.argv) > 1 else 40
ins = []
for line in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_idx = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr_idx and i - addr_idx[tgt] >= min_len:
            body = ins[addr_idx[tgt]: i + 1]
            c = Counter()
            for _, s in body:
                s = re.sub(r"^@!?U?P\d+\s+", "", s)
                op = s.split()[0].split(".")[0]
                c[op] += 1
            n = len(body)
            packed = c["FFMA2"] + c["FMUL2"] + c["FADD2"]
            print(f"loop {tgt:#x}..{a:#x}: {n} instr, packed {packed}, MUFU {c['MUFU']}, "
                  f"slots {n + packed}")
            print("   ", dict(c.most_common(24)))
