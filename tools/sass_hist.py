#!/usr/bin/env python
"""Opcode histogram of one kernel in a cuobjdump -sass listing: tools/sass_hist.py <binary-or-.so> <substring-of-mangled-name> [top]"""
import collections
import re
import subprocess
import sys

binary, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
sass = subprocess.run(["cuobjdump", "-sass", binary], capture_output=True, text=True).stdout
cur, hist = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        hist[cur][m.group(2)] += 1
for fn, h in hist.items():
    if pat in fn:
        print(fn, "total", sum(h.values()))
        for op, c in h.most_common(top):
            print(f"  {c:6d} {op}")
