"""Summarise `nvcc -Xptxas -v` output: registers / stack / spills per unproject_kernel instantiation.
usage: python scripts/ptxas_table.py LOG [filter-substring]"""
import re
import sys

t = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
pat = re.compile(r"Compiling entry function '(\S+)'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers")
for m in pat.finditer(t):
    name = m.group(1)
    k = re.search(r"unproject_kernelILi(\d+)ELb(\d)ELb(\d)ELb(\d)ELi(\d)ELi(\d)ELi(\d)", name)
    tag = "V%s EX%s CACHE%s BF%s M%s LPB%s OUT%s" % k.groups() if k else name[:70]
    if flt in tag:
        print("%-44s stack %4s spill st/ld %4s/%4s regs %s" % (tag, m.group(2), m.group(3), m.group(4), m.group(5)))
