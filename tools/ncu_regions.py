"""Where a kernel spends its time along the program: stall samples and executed instructions per block of
consecutive SASS instructions, from `ncu -i <rep> --page source --csv --print-source sass > file.csv`.
    python tools/ncu_regions.py file.csv <kernel substring> [chunk=50]"""
import csv
import re
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 50
starts = [k for k, r in enumerate(rows) if r and r[0] == "Kernel Name" and want in r[1]]
i = starts[0]
hdr = rows[i + 1]
H = {h: k for k, h in enumerate(hdr)}
body = []
for r in rows[i + 2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr):
        body.append(r)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[H["# Samples"]] or 0) for r in body)
tex = sum(int(r[H["Instructions Executed"]] or 0) for r in body)
print(rows[i][1][:90], "samples", tot, "warp-inst", tex)
for s in range(0, len(body), chunk):
    blk = body[s:s + chunk]
    smp = sum(int(r[H["# Samples"]] or 0) for r in blk)
    ex = sum(int(r[H["Instructions Executed"]] or 0) for r in blk)
    st = {k: sum(int(r[H[k]] or 0) for r in blk) for k in stalls}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    ops = [re.sub(r"^@!?U?P\d+\s+", "", r[H["Source"]]).split()[0].split(".")[0] for r in blk]
    oc = Counter(ops).most_common(5)
    print("%5s  samp %5.1f%%  exec %5.1f%%  %s   %s" % (blk[0][H["Address"]][-5:] if "Address" in H else s, 100 * smp / tot, 100 * ex / tex,
          [(k.replace("stall_", ""), v) for k, v in top], oc))
