"""Static SASS opcode histogram of one kernel of libflan_b200.so (whole function; the frame loop dominates).

    python tools/sass_static.py <substring of the mangled name> [--dump]
"""
import collections
import re
import subprocess
import sys

lib = "flan_b200/lib/libflan_b200.so"
want = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, body = None, collections.defaultdict(list)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m and cur:
        body[cur].append(m.group(2).strip())
for name, ins in body.items():
    if want not in name:
        continue
    ops = collections.Counter()
    for i in ins:
        i = re.sub(r"^@!?U?P\d+\s+", "", i)
        ops[i.split()[0].split(".")[0]] += 1
    print(name, "static", len(ins))
    print("  " + "  ".join("%s %d" % kv for kv in ops.most_common(40)))
    if "--dump" in sys.argv:
        print("\n".join(ins))
