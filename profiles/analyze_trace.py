#!/usr/bin/env python
"""Summarise a profiles/trace_fused.py timeline: per-tag mean gaps of each slot and the per-layer period."""
import collections
import sys

rows = [l.split() for l in open(sys.argv[1]) if l[:1] in '012']
rows = [(int(a), int(b), int(c), int(d)) for a, b, c, d in rows]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else None
for slot in range(3):
    r = [x for x in rows if x[0] == slot]
    if not r:
        continue
    print("slot", slot, "events", len(r), "span", r[-1][3] - r[0][3])
    d = collections.defaultdict(list)
    for x in r[len(r) // 4:]:
        d[x[1]].append(x[2])
    print("  mean gap before tag:", {k: int(sum(v) / len(v)) for k, v in sorted(d.items())})
if lo is not None:
    for x in sorted([x for x in rows if lo < x[3] < lo + int(sys.argv[3])], key=lambda x: x[3]):
        print(x)
