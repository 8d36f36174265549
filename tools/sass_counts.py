"""SASS evidence for the tcgen05 / TMA kernels: per kernel of libb2second.so the count of the mnemonics that prove the
Blackwell-native path (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / tcgen05.st, UTMALDG = TMA tensor load, LDGSTS = cp.async,
ARRIVES.LDGSTSBAR / SYNCS = mbarrier traffic) and of the legacy ones that must be absent (HMMA, HGMMA).
Usage: python tools/sass_counts.py [lib.so] > profiles/r2_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                                                          "second.pytorch_b200", "csrc", "libb2second.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "ARRIVES.LDGSTSBAR", "SYNCS", "HMMA", "HGMMA",
         "FFMA", "LDG", "STG", "ATOMG", "RED"]
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        mangled = m.group(1)
        cur = mangled
        k = re.search(r"_cu_[0-9a-f]{8}(\d+)(k_)", mangled)    # <file hash (8 hex)><length><identifier>
        if k:                                            # Itanium: <length><identifier>[I<template args>E]
            n, start = int(k.group(1)), k.start(2)
            name, rest = mangled[start:start + n], mangled[start + n:]
            t = re.match(r"I((?:Li\d+E)+)E", rest)
            cur = name + ("<" + ",".join(re.findall(r"Li(\d+)E", t.group(1))) + ">" if t else "")
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                counts[cur][w] += 1
print("# cuobjdump -sass %s : mnemonic counts per kernel" % os.path.basename(lib))
print("# tcgen05.mma -> UTCHMMA, tcgen05.ld / st -> LDTM / STTM, cp.async.bulk.tensor -> UTMALDG, cp.async -> LDGSTS; HMMA/HGMMA (legacy "
      "mma.sync / wgmma) must be 0")
print("%-44s " % "kernel" + " ".join("%9s" % w[:9] for w in WATCH))
tot = collections.Counter()
for k, c in counts.items():
    print("%-44s " % k[:44] + " ".join("%9d" % c[w] for w in WATCH))
    tot.update(c)
print("%-44s " % "TOTAL" + " ".join("%9d" % tot[w] for w in WATCH))
