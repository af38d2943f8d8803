#!/usr/bin/env python
"""Summarise nvcc -Xptxas -v logs: kernel, registers, stack, spills, smem."""
import re, subprocess, sys, glob
for log in sorted(glob.glob('libcpab_b200/csrc/build/*.ptxas.log')):
    txt = open(log).read()
    names = re.findall(r"Compiling entry function '(\S+)'", txt)
    if not names:
        continue
    dem = subprocess.run(['c++filt'] + names, capture_output=True, text=True, stdin=subprocess.DEVNULL).stdout.splitlines()
    blocks = re.split(r"Compiling entry function '\S+' for 'sm_100a'", txt)[1:]
    for n, b in zip(dem, blocks):
        if len(sys.argv) > 1 and not re.search(sys.argv[1], n): continue
        regs = re.search(r"Used (\d+) registers", b).group(1)
        stack = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
        smem = re.search(r"(\d+) bytes smem", b)
        short = re.sub(r"\(.*", "", n).replace("void cpab::", "").replace("(anonymous namespace)::", "")
        print(f"{short:70s} regs={regs:>3s} stack={stack.group(1):>4s} spill={stack.group(2)}/{stack.group(3)} smem={smem.group(1) if smem else 0}")
