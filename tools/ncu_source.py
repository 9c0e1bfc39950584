"""Instruction mix and top stall lines of one kernel from an ncu report's source page (development aid).
    python tools/ncu_source.py <rep> <kernel regex> [top=25]"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several launches may match: keep the last block
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
rows = rows[starts[-1]:]
print(rows[0][1][:100])
hdr = rows[1]
ci = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ci["# Samples"]]) for r in data)
ex, st = Counter(), Counter()
for r in data:
    toks = r[ci["Source"]].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    ex[op] += int(r[ci["Instructions Executed"]])
    st[op] += int(r[ci["# Samples"]])
te = sum(ex.values())
print("warp instructions executed: %d, samples: %d, SASS lines: %d" % (te, tot, len(data)))
print("%-10s %8s %8s" % ("op", "exec%", "stall%"))
for op, c in ex.most_common(24):
    print("%-10s %7.1f%% %7.1f%%" % (op, 100.0 * c / te, 100.0 * st[op] / tot))
print()
for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:top]:
    print(r[ci["# Samples"]].rjust(6), r[ci["Source"]].strip()[:100])
