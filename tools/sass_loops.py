#!/usr/bin/env python3
"""List the backward-branch loops of one kernel in a cubin/object with an instruction-mix count per loop.
usage: sass_loops.py <object> <function-substring>"""
import re, subprocess, sys, collections
obj, needle = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, funcs = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if needle not in name: continue
    print("==", name, len(ins), "instructions")
    addr = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
        if m and "BRA.DIV" not in t:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr:
                body = ins[addr[tgt]:i + 1]
                mix = collections.Counter()
                for _, b in body:
                    op = re.sub(r"^@!?U?P\d+\s+", "", b).split()[0].split(".")[0]
                    mix[op] += 1
                print("  loop 0x%04x..0x%04x: %d instr  %s" % (tgt, a, len(body), dict(mix.most_common(14))))
