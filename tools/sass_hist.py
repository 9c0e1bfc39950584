"""Dynamic SASS opcode histogram and stall summary from `ncu --page source --csv --print-source sass` output."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            body.append(rows[j])
            j += 1
        i = j
        if want not in name:
            continue
        H = {h: k for k, h in enumerate(hdr)}
        ie, ss, src = H["Instructions Executed"], H["# Samples"], H["Source"]
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        ops, samp = collections.Counter(), collections.Counter()
        st = collections.Counter()
        tot = 0
        for r in body:
            if len(r) < len(hdr):
                continue
            n = int(r[ie] or 0)
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[src])
            op = m.group(2) if m else "?"
            key = op.split(".")[0]
            if key in ("LDG", "STG", "LDS", "STS", "MUFU", "FRND", "F2F", "F2I", "I2F"):
                key = ".".join(op.split(".")[:3]) if key in ("MUFU",) else key
            ops[key] += n
            samp[key] += int(r[ss] or 0)
            tot += n
            for s_ in stalls:
                st[s_] += int(r[H[s_]] or 0)
        print(name[:70], "warp-inst", tot, "static", len(body))
        print(" stalls:", [(k.replace("stall_", ""), v) for k, v in st.most_common(9)])
        ts = sum(samp.values())
        print(" opcodes (% of executed | % of samples):")
        print("  " + "  ".join("%s %.1f|%.1f" % (k, 100.0 * v / tot, 100.0 * samp[k] / max(ts, 1)) for k, v in ops.most_common(36)))
    else:
        i += 1
