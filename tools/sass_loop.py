#!/usr/bin/env python
"""Print the SASS of one kernel (demangled-name regex) and instruction counts between backward branches.

    python tools/sass_loop.py 'k_forward<float, 2, true, true, 1>' [--dump]
"""
import re, subprocess, sys
so = "libcpab_b200/csrc/libcpab_b200.so"
pat = sys.argv[1]
for a in sys.argv[2:]:
    if a.endswith((".cubin", ".so", ".o")):
        so = a
names = subprocess.run(["cuobjdump", "-elf", so], capture_output=True, text=True).stdout
syms = sorted(set(re.findall(r"\b(_ZN\S+?)\b", names)))
dem = subprocess.run(["c++filt"] + syms, capture_output=True, text=True, stdin=subprocess.DEVNULL).stdout.splitlines()
hit = [s for s, d in zip(syms, dem) if pat in d and not s.startswith("_ZN4cpab") is False][:1] or [s for s, d in zip(syms, dem) if pat in d][:1]
if not hit:
    sys.exit("no kernel matches " + pat)
sass = subprocess.run(["cuobjdump", "-sass", "-fun", hit[0], so], capture_output=True, text=True).stdout
ins = []
for line in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\*", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print(hit[0], "instructions:", len(ins))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s+)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr_index:
        j = addr_index[int(m.group(1), 16)]
        body = [x for _, x in ins[j:i + 1]]
        ops = {}
        for x in body:
            op = re.sub(r"^@!?U?P\d+\s+", "", x).split()[0].split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        print(f"loop {ins[j][0]:#x}..{a:#x}: {len(body)} instrs ", dict(sorted(ops.items(), key=lambda kv: -kv[1])))
if "--dump" in sys.argv:
    for a, t in ins:
        print(f"{a:05x}  {t}")
