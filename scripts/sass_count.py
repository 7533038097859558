"""count SASS instructions per kernel in an object / shared library (cuobjdump -sass)"""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
name, cnt = None, {}
for l in out.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m:
        name = m.group(1); cnt[name] = 0
    elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        cnt[name] += 1
for k, v in sorted(cnt.items(), key=lambda kv: kv[1]):
    print("%7d  %6.1f KB  %s" % (v, v * 16 / 1024, k))
