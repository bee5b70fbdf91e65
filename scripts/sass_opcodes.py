#!/usr/bin/env python
"""Per-kernel SASS opcode evidence for the tensor-core / TMA claims: dumps libbf_b200.so with cuobjdump and counts, per
kernel, the mnemonics that prove the path -- UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG /
UTMASTG (cp.async.bulk.tensor), UBLKCP (cp.async.bulk), SYNCS (mbarrier), plus HMMA / FFMA for the SIMT fall-backs.

    python scripts/sass_opcodes.py > profiles/r02_sass_opcodes.txt          (no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "brainforge_b200", "libbf_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "FFMA", "LDG", "STG", "LDS", "STS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    names = dict(zip(re.findall(r"Function : (\S+)", sass), demangle))
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = names.get(m.group(1), m.group(1))
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            per[cur][m.group(1).split(".")[0]] += 1
    arch = re.findall(r"arch = (sm_\w+)", sass)
    print("# cuobjdump -sass brainforge_b200/libbf_b200.so  (arch %s), mnemonic counts per kernel" % ", ".join(sorted(set(arch))))
    print("# %-96s %s" % ("kernel", " ".join("%7s" % w for w in WATCH)))
    def short(k):
        k = re.sub(r"\((?:int|bool|unsigned int|long)\)", "", k)
        k = re.sub(r"\(.*", "", k).replace("void ", "").replace("bf::", "")
        return k if len(k) <= 98 else k[:95] + "..."

    fams = collections.OrderedDict()
    for k, c in per.items():
        fams.setdefault(short(k).split("<")[0], []).append(c)
    print("# ---- by kernel family: instances, then min..max count over the instances")
    for fam, cs in fams.items():
        cells = []
        for w in WATCH:
            lo, hi = min(c.get(w, 0) for c in cs), max(c.get(w, 0) for c in cs)
            cells.append("%7s" % (str(lo) if lo == hi else "%d..%d" % (lo, hi)))
        print("%-98s %s" % ("%s  x%d" % (fam, len(cs)), " ".join(cells)))
    print("# ---- every instance")
    for k, c in per.items():
        print("%-98s %s" % (short(k), " ".join("%7d" % c.get(w, 0) for w in WATCH)))
    tc = [k for k, c in per.items() if c.get("UTCHMMA", 0)]
    print("# %d kernels, %d with tcgen05.mma (UTCHMMA), %d with tensor TMA loads (UTMALDG), %d with bulk copies (UBLKCP)" %
          (len(per), len(tc), sum(1 for c in per.values() if c.get("UTMALDG", 0)), sum(1 for c in per.values() if c.get("UBLKCP", 0))))


if __name__ == "__main__":
    sys.exit(main())
