"""Summarise an .ncu-rep (ncu --set full) into a per-kernel table + profiles/traffic.json.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_full [c2]       -> .csv / .md / traffic.json[workload]
  python tools/ncu_summary.py --launches gpurun_out/launches.csv profiles/r01_launches  -> per-kernel share of the step
"""
import csv
import json
import os
import subprocess
import sys
from collections import OrderedDict, defaultdict

METRICS = OrderedDict([
    ("gpu__time_duration.sum", "dur_us"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
])


def to_bytes(v, unit):
    v = float(v)
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def to_us(v, unit):
    v = float(v)
    return v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6, "second": 1e6}.get(unit.lower(), 1)


def short(name):
    n = name.split("(")[0].replace("void ", "").replace("vmp::", "")
    base = n.split("<")[0]
    if base == "k_fill":                                    # k_fill<true> is the CTA path (bench.py: k_fill_heavy)
        return "k_fill_heavy" if ("(bool)1" in n or "<true>" in n or "<1>" in n) else "k_fill"
    return base


def full(rep, out, key="c2"):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    per = defaultdict(list)
    for r in rows[2:]:
        rec = {}
        for m, k in METRICS.items():
            if m not in idx or r[idx[m]] == "":
                continue
            v, u = r[idx[m]].replace(",", ""), units[idx[m]]
            try:
                rec[k] = to_us(v, u) if k == "dur_us" else to_bytes(v, u) if k in ("dram_rd", "dram_wr") else float(v)
            except ValueError:
                pass
        per[short(r[idx["Kernel Name"]])].append(rec)
    cols = ["kernel", "launches", "work_launches"] + list(METRICS.values())
    lines, traffic = [], {}
    for k, recs in per.items():
        # launches that did work (the IEKF kernels early-exit once the filter has converged)
        top = max(x.get("warp_inst", 0) for x in recs)
        work = [x for x in recs if x.get("warp_inst", 0) > 0.1 * top] or recs
        avg = {c: sum(x.get(c, 0.0) for x in work) / len(work) for c in METRICS.values()}
        lines.append([k, len(recs), len(work)] + [round(avg[c], 3) for c in METRICS.values()])
        traffic[k] = round(avg["dram_rd"] + avg["dram_wr"], 1)
    lines.sort(key=lambda r: -r[3] * r[2])
    with open(out + ".csv", "w", newline="") as f:
        w = csv.writer(f); w.writerow(cols); w.writerows(lines)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu --set full summary of `{os.path.basename(rep)}` (cold-cache, serialised replays; averages over the launches that did work)\n\n")
        f.write("| " + " | ".join(cols) + " |\n|" + "---|" * len(cols) + "\n")
        for r in lines:
            f.write("| " + " | ".join(str(x) for x in r) + " |\n")
        f.write("\ntensor_inst = 0 everywhere: the path is a gather/reduction, tensor cores are not used (DESIGN.md §4).\n")
    tpath = os.path.join(os.path.dirname(out), "traffic.json")
    old = json.load(open(tpath)) if os.path.exists(tpath) else {}
    old.setdefault(key, {}).update({k.replace("<1>", ""): v for k, v in traffic.items()})   # keyed by bench workload
    json.dump(old, open(tpath, "w"), indent=1, sort_keys=True)
    print("wrote", out + ".csv", out + ".md", tpath)


def launches(csv_path, out):
    rows = [r for r in csv.reader(open(csv_path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    per = defaultdict(list)
    for r in rows[1:]:
        try:
            per[short(r[ki])].append(to_us(r[vi].replace(",", ""), r[ui]))
        except (ValueError, IndexError):
            pass
    tot = sum(sum(v) for v in per.values())
    lines = sorted(((k, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot) for k, v in per.items()), key=lambda r: -r[3])
    with open(out + ".md", "w") as f:
        f.write(f"# ncu launch list `{os.path.basename(csv_path)}`: gpu__time_duration.sum per launch (cold-cache, serialised — compare SHARES)\n\n")
        f.write("| kernel | launches | avg us | total us | share % |\n|---|---|---|---|---|\n")
        for k, n, a, s, p in lines:
            f.write(f"| {k} | {n} | {a:.2f} | {s:.1f} | {p:.1f} |\n")
    print("wrote", out + ".md")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "c2")
