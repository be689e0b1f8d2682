"""SASS listings of the hot kernels (profiles/sass_<tag>_<kernel>.txt): `cuobjdump -sass` of the built library, one file per kernel,
encoding words stripped, with an instruction count by mnemonic in the header (UBLKCP / SYNCS = cp.async.bulk + mbarrier,
UCGABAR_* = cluster barrier, no HMMA / UTCMMA anywhere: the path has no tensor-shaped work).
  python tools/sass_listing.py [tag=r02]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = os.path.join(ROOT, "voxelmapplus_fastlio2_b200", "libvmp_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
want = {"k_iekf_loopILb0E": "k_iekf_loop", "k_fillILb1E": "k_fill_heavy", "k_fillILb0E": "k_fill", "k_merge_rounds": "k_merge_rounds",
        "k_world_insert_count": "k_world_insert_count", "k_merge_prefilter": "k_merge_prefilter"}
notable = ("UBLKCP", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "UTMALDG", "HMMA", "UTCHMMA", "ATOMS", "ATOMG", "RED", "MEMBAR", "CCTL", "LDG", "STG", "LDS", "STS",
           "DFMA", "DADD", "DMUL", "MUFU", "BAR")
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name = part.split("\n", 1)[0].strip()
    for key, short in want.items():
        if key not in name:
            continue
        ops = collections.Counter()
        lines = []
        for line in part.split("\n"):
            if re.match(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", line):
                continue
            line = re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line).rstrip()
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                ops[m.group(1).split(".")[0]] += 1
            lines.append(line)
        hdr = (f"# SASS of {name} (sm_100a): cuobjdump -sass voxelmapplus_fastlio2_b200/libvmp_b200.so, encoding words stripped (tools/sass_listing.py)\n"
               f"# {sum(ops.values())} instructions; by mnemonic: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(28)) + "\n"
               "# notable: " + ", ".join(f"{k} {ops[k]}" for k in notable if ops[k]) + "\n\n")
        out = os.path.join(ROOT, "profiles", f"sass_{tag}_{short}.txt")
        open(out, "w").write(hdr + "Function : " + "\n".join(lines).rstrip() + "\n")
        print(short, sum(ops.values()), {k: ops[k] for k in notable if ops[k]})
