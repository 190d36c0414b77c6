"""Reads `nvcc -Xptxas -v ... | c++filt` on stdin and prints one line per kernel: registers, stack / spills, shared memory."""
import re
import sys

name, spill = None, ""
for line in sys.stdin:
    m = re.search(r"Compiling entry function .(.*). for .sm_100a.", line)
    if m:
        name, spill = m.group(1), ""
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        spill = f"stack {m.group(1)} B, spill st/ld {m.group(2)}/{m.group(3)} B"
        continue
    m = re.search(r"Used (\d+) registers", line)
    if m and name:
        sm = re.search(r"(\d+) bytes smem", line)
        smem = (sm.group(1) + " B smem") if sm else ""
        print(f"{int(m.group(1)):4d} regs  {spill:34s} {smem:14s} {name[:150]}")
        name = None
