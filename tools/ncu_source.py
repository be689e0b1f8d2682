"""Per-source-line view of one kernel of an .ncu-rep (ncu --set full --import-source on; the library is built with -lineinfo):
stall samples and executed warp instructions per line of the CUDA sources.

  python tools/ncu_source.py gpurun_out/prof.ncu-rep k_merge_rounds [top=30] [launch_index=0] > profiles/<name>.md
"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur, hdr, launch = None, None, -1
    items = {}
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) >= 2 and r[0] == "Function Name":
            continue
        if len(r) >= 2 and r[0] == "Kernel Name":
            launch += 1
            continue
        if len(r) >= 3 and r[0] == "Line No":
            hdr = r
            if launch < 0:
                launch = 0
            continue
        if hdr and len(r) == len(hdr) and r[0] != "" and launch == which:
            d = {}
            for k, v in zip(hdr, r):
                d.setdefault(k, v)
            try:
                s, ie = int(d.get("# Samples") or 0), int(d.get("Instructions Executed") or 0)
            except ValueError:
                continue
            key = (cur, int(r[0]))
            a = items.setdefault(key, [0, 0, r[1].strip()])
            a[0] += s
            a[1] += ie
    ts = sum(v[0] for v in items.values()) or 1
    ti = sum(v[1] for v in items.values()) or 1
    print(f"# {kern}: ncu source view (`--set full --import-source on`), launch {which}: {ts} stall samples, {ti} warp instructions\n")
    print("| % samples | % instructions | file:line | source |\n|---|---|---|---|")
    for (f, l), (s, ie, src) in sorted(items.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"| {100 * s / ts:.1f} | {100 * ie / ti:.1f} | {f}:{l} | `{src[:170]}` |")


if __name__ == "__main__":
    main()
