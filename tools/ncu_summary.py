"""Summarise ncu captures into profiles/: a markdown table of the metrics DESIGN.md cites and profiles/ncu_traffic.json
(dram bytes per launch, read by bench.py for roofline.traffic).

    python tools/ncu_summary.py <tag> <title> <rep> [<rep> ...]
Each .ncu-rep is read with `ncu -i <rep> --page raw --csv`; one column per captured kernel launch (the LAST launch of
each kernel name in a report is kept, i.e. a warm one)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
]
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TO_MS = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def short(name):
    m = re.search(r"(pv_\w+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name


def read(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[head], rows[head + 1]
    col = {n: i for i, n in enumerate(names)}
    res = {}
    for r in rows[head + 2:]:
        if len(r) != len(names):
            continue
        k = short(r[col["Kernel Name"]])
        res[k] = {m: (r[col[m]], units[col[m]]) for m in METRICS if m in col}
    return res


def main():
    tag, title, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    kernels = {}
    for rep in reps:
        kernels.update(read(rep))
    names = list(kernels)
    lines = ["# " + title, "", "Read with `ncu -i <rep> --page raw --csv` by tools/ncu_summary.py from: " +
             ", ".join("`%s`" % os.path.basename(r) for r in reps) + " (gpurun_out/, scratch).", "",
             "| metric | unit | " + " | ".join(names) + " |", "|---|---|" + "---|" * len(names)]
    for m in METRICS:
        vals = [kernels[k].get(m, ("", "")) for k in names]
        unit = next((u for _, u in vals if u), "")
        lines.append("| %s | %s | " % (m, unit) + " | ".join(v for v, _ in vals) + " |")
    open(os.path.join(ROOT, "profiles", tag + "_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        traffic = json.load(open(tpath))
    except Exception:
        traffic = {"kernels": {}}
    traffic["note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full captures summarised in "
                       "profiles/*_ncu_summary.md; bench.py copies these into roofline.traffic")
    for k in names:
        try:
            rd, ru = kernels[k]["dram__bytes_read.sum"]
            wr, wu = kernels[k]["dram__bytes_write.sum"]
            tm, tu = kernels[k]["gpu__time_duration.sum"]
            rd, wr = float(rd.replace(",", "")) * TO_BYTES[ru], float(wr.replace(",", "")) * TO_BYTES[wu]
            key = re.sub(r"<(\d+),.*>", r"<\1>", k)
            if k.startswith("pv_analysis_kernel<") and k.count(",") == 5 and re.search(r",\s*1>$", k):
                key += " +summaries"        # the instantiation that also leaves the phase summaries (last template argument)
            traffic["kernels"][key] = {"dram_bytes_per_launch": rd + wr, "read": rd, "write": wr, "capture": tag,
                                       "time_ms": float(tm.replace(",", "")) * TO_MS[tu]}
        except Exception as e:
            print("skip", k, e)
    json.dump(traffic, open(tpath, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
