"""Static SASS view of one kernel in an object / library: opcode histogram, and the same restricted to the
instructions between two addresses (e.g. the frame loop's fast path).

    python tools/sass_fn.py <file> <substring of mangled name> [--dump] [--range lo hi]
"""
import collections
import re
import subprocess
import sys

path, want = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
cur, body = None, collections.defaultdict(list)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m and cur:
        body[cur].append((int(m.group(1), 16), m.group(2).strip()))
lo, hi = 0, 1 << 30
if "--range" in sys.argv:
    i = sys.argv.index("--range")
    lo, hi = int(sys.argv[i + 1], 16), int(sys.argv[i + 2], 16)
for name, ins in body.items():
    if want not in name:
        continue
    ops = collections.Counter()
    n = 0
    for addr, i in ins:
        if not (lo <= addr <= hi):
            continue
        i = re.sub(r"^@!?U?P\d+\s+", "", i)
        op = i.split()[0]
        key = op.split(".")[0]
        if key in ("LDG", "STG", "LDS", "STS", "LDL", "STL"):
            key = ".".join(op.split(".")[:1]) + ("." + [x for x in op.split(".") if x in ("64", "128")][0] if any(x in ("64", "128") for x in op.split(".")) else "")
        ops[key] += 1
        n += 1
    print(name, "static", n)
    print("  " + "  ".join("%s %d" % kv for kv in ops.most_common(60)))
    if "--dump" in sys.argv:
        for addr, i in ins:
            if lo <= addr <= hi:
                print("%05x  %s" % (addr, i))
