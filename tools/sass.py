#!/usr/bin/env python
"""Prints the SASS of one kernel from an object file: tools/sass.py <obj> <name substring> [--grep PATTERN]"""
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", out)
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if pat in name:
        lines = [l for l in b.split("\n") if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l)]
        for l in lines:
            m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m:
                print(m.group(1), m.group(2).strip())
        break
