"""Summarise `nvcc -Xptxas -v` output: one line per kernel (demangled): regs, spills, stack.
usage: nvcc ... -Xptxas -v 2>&1 | python tools/ptxas_summary.py [filter]"""
import re, subprocess, sys
txt = sys.stdin.read()
flt = sys.argv[1] if len(sys.argv) > 1 else ""
names = re.findall(r"Compiling entry function '(\S+)'", txt)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
blocks = txt.split("Compiling entry function '")[1:]
for blk, d in zip(blocks, dem):
    if flt and flt not in d:
        continue
    regs = re.search(r"Used (\d+) registers", blk)
    sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", blk)
    short = re.sub(r"\(.*", "", d).replace("void cpab::", "")
    print(f"{short:70s} regs={regs.group(1) if regs else '?':>3} stack={sp.group(1) if sp else '?':>4} spillst={sp.group(2) if sp else '?':>4} spillld={sp.group(3) if sp else '?':>4}")
