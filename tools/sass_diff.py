#!/usr/bin/env python
"""Compares the SASS of every kernel function in two builds of libqoi_gpu.so, ignoring addresses.
Used to prove that adding an experimental (default-off) template variant left the shipped kernels untouched.
usage: sass_diff.py <old.so> <new.so> [old_name_substr=new_name_substr ...]   (name rewrites applied to the NEW names)"""
import hashlib
import re
import subprocess
import sys


def funcs(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    out, name, body = {}, None, []
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            if name:
                out[name] = "\n".join(body)
            name, body = m.group(1), []
        elif name:
            ln = re.sub(r"/\*[0-9a-f]{4,}\*/", "", ln).strip()
            if ln:
                body.append(ln)
    if name:
        out[name] = "\n".join(body)
    return out


old, new = funcs(sys.argv[1]), funcs(sys.argv[2])
rewrites = [a.split("=", 1) for a in sys.argv[3:]]
renamed = {}
for k, v in new.items():
    for a, b in rewrites:
        k = k.replace(b, a)
    renamed.setdefault(k, v)
bad = 0
for k, v in sorted(old.items()):
    if k not in renamed:
        print("missing in new:", k)
        bad += 1
    elif hashlib.md5(v.encode()).digest() != hashlib.md5(renamed[k].encode()).digest():
        print("DIFFERENT:", k, len(v.splitlines()), "->", len(renamed[k].splitlines()), "lines")
        bad += 1
print(f"{len(old)} functions in old, {len(new)} in new, {bad} differences")
sys.exit(1 if bad else 0)
