#!/usr/bin/env python
"""Instruction count of the straight-line body of a strip kernel: from the first 128-bit load to the last 128-bit store.
usage: sass_body.py <object> <mangled-name-substring>"""
import re, subprocess, sys, collections
obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
on, ins = False, []
for line in txt.splitlines():
    if "Function : " in line:
        on = pat in line
        continue
    if on:
        m = re.match(r"\s*/\*([0-9a-f]{4})\*/\s+(.*?);", line)
        if m:
            ins.append(m.group(2).strip())
first = next(i for i, s in enumerate(ins) if "LDG.E.128" in s or "LDG.E.64" in s and ".128" in s)
last = max(i for i, s in enumerate(ins) if "STG.E.128" in s)
body = ins[first:last + 1]
ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0] for s in body)
print("total", len(ins), "preamble", first, "body", len(body), "tail", len(ins) - last - 1)
print(" ".join("%s:%d" % kv for kv in ops.most_common(22)))
print("branches in body:", sum(1 for s in body if re.search(r"\bBRA\b", s)))
