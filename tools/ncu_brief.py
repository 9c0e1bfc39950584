"""Key metrics + per-barrier-phase breakdown of one kernel from an .ncu-rep (development aid).

    python tools/ncu_brief.py report.ncu-rep [kernel substring]
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "sm__warps_active.avg.per_cycle_active", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if want not in name:
        continue
    print(name[:100])
    for k in keys:
        if k in hdr:
            print("  %-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    st = [(h, float(r[i] or 0)) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    st.sort(key=lambda x: -x[1])
    print("  stalls/issue:", " ".join("%s=%.2f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v) for h, v in st[:9]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) >= len(hdr):
                body.append(rows[j])
            j += 1
        i = j
        if want not in name:
            continue
        H = {h: k for k, h in enumerate(hdr)}
        tot = sum(int(r[H["# Samples"]] or 0) for r in body)
        totx = sum(int(r[H["Instructions Executed"]] or 0) for r in body)
        seg, cur = [], []
        for r in body:
            cur.append(r)
            if "BAR.SYNC" in r[H["Source"]]:
                seg.append(cur)
                cur = []
        seg.append(cur)
        print(name[:100], "dyn warp-inst", totx, "static", len(body))
        for k, s in enumerate(seg):
            smp = sum(int(r[H["# Samples"]] or 0) for r in s)
            ex = sum(int(r[H["Instructions Executed"]] or 0) for r in s)
            ops, st = collections.Counter(), collections.Counter()
            for r in s:
                m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[H["Source"]])
                ops[m.group(2) if m else "?"] += int(r[H["Instructions Executed"]] or 0)
                for h in hdr:
                    if h.startswith("stall_") and "Not Issued" not in h:
                        st[h[6:]] += int(r[H[h]] or 0)
            print(" phase%d static %d exec %.1f%% samples %.1f%% | %s | %s" % (
                k, len(s), 100.0 * ex / max(totx, 1), 100.0 * smp / max(tot, 1),
                " ".join("%s:%.0f" % (o, 100.0 * c / max(ex, 1)) for o, c in ops.most_common(10)),
                " ".join("%s:%d" % kv for kv in st.most_common(6))))
    else:
        i += 1
