"""Print the SASS of one function (substring match on the mangled name) from a cubin/.so: python sass_fn.py lib.so pattern"""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout.splitlines()
on = False
for l in out:
    if "Function : " in l:
        on = sys.argv[2] in l
        continue
    if on:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            print(m.group(1), m.group(2).strip())
