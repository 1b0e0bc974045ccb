#!/usr/bin/env python
"""Print the SASS of one kernel from a built library: tools/sass_fn.py lib.so kernel_name"""
import subprocess
import sys

lib, name = sys.argv[1], sys.argv[2]
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
on = False
for line in out.splitlines():
    if 'Function :' in line:
        on = line.strip().endswith(': ' + name) or line.strip().endswith(name)
    if on:
        print(line)
