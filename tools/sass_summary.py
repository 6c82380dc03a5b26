#!/usr/bin/env python
"""Per-kernel count of the SASS mnemonics that prove which units a kernel uses (tcgen05 = UTC*MMA / LDTM / UTCBAR,
TMA = UTMALDG / UBLKCP, fp64 tensor = DMMA, cp.async = LDGSTS, clusters = UCGABAR, mbarrier = SYNCS):
  python tools/sass_summary.py build/*.o > profiles/rNN_sass_mnemonics_per_kernel.txt   (cuobjdump -sass underneath)"""
import collections
import re
import subprocess
import sys

PAT = re.compile(r"\b(UTCIMMA|UTCHMMA|UTCQMMA|UTCBAR|UTCCP|UTMALDG(?:\.\dD)?|UTMASTG|UBLKCP(?:\.[A-Z.]+)?|LDTM(?:\.x\d+)?|STTM|DMMA\.\w+|HMMA\.\w+|IMMA\.\w+|"
                 r"LDGSTS(?:\.[A-Z.0-9]+)?|UCGABAR_\w+|SYNCS\.[A-Z0-9.]+|MUFU\.RSQ64H|DFMA|MEMBAR\.[A-Z.]+|CCTL\.\w+|FENCE\.[A-Z.]+|ATOMS?\.\w+|REDUX\.\w+)\b")
for obj in sys.argv[1:]:
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn, counts, total = None, collections.OrderedDict(), {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()[:150]
            counts[fn] = collections.Counter()
            total[fn] = 0
            continue
        if fn and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            total[fn] += 1
            for k in PAT.findall(line):
                counts[fn][k] += 1
    print(f"== {obj}")
    for fn, c in counts.items():
        if not c:
            continue
        print(f"  {fn}  [{total[fn]} instructions]")
        print("     " + "  ".join(f"{k} x{v}" for k, v in sorted(c.items())))
