#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: stall reasons, opcode mix, hottest SASS lines.

    ncu -i prof.ncu-rep --page source --csv -k regex:k_backward > src.csv
    python tools/ncu_source_summary.py src.csv [top_n]
"""
import collections
import csv
import re
import sys


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for sec in sections[:1]:
    hdr, data = sec["hdr"], sec["data"]
    idx = {h: i for i, h in enumerate(hdr)}
    ts = sum(num(r[idx["# Samples"]]) for r in data)
    ti = sum(num(r[idx["Instructions Executed"]]) for r in data)
    print(sec["name"][:90])
    print("sass lines", len(data), "samples", ts, "warp insts", ti)
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(num(r[idx[h]]) for r in data) for h in reasons}
    print("stalls:", ", ".join(f"{k[6:]}={v / max(ts, 1) * 100:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
    ops, smp = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[idx["Source"]])
        op = m.group(2).split(".")[0] if m else "?"
        ops[op] += num(r[idx["Instructions Executed"]])
        smp[op] += num(r[idx["# Samples"]])
    print("opcode     exec%  samples%")
    for op, c in ops.most_common(24):
        print(f"{op:10s} {c / max(ti, 1) * 100:5.1f}  {smp[op] / max(ts, 1) * 100:5.1f}")
    print("hottest SASS (samples, executed, instruction, dominant stalls)")
    for r in sorted(data, key=lambda r: -num(r[idx["# Samples"]]))[:top]:
        st = sorted(((num(r[idx[h]]), h[6:]) for h in reasons), reverse=True)[:2]
        print(r[idx["Address"]][-5:], str(num(r[idx["# Samples"]])).rjust(6), str(num(r[idx["Instructions Executed"]])).rjust(11),
              r[idx["Source"]][:64].ljust(64), " ".join(f"{n}:{c}" for c, n in st))
