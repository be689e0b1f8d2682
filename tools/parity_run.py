"""Parity at the full BASELINE.json lengths (the same procedures as tests/test_gpu_parity_sizes.py, tests/parity_cases.py):

  python tools/parity_run.py [--c1 1000] [--c2 200] [--c3 10000] [--c4 50000000] [--out profiles/r02_parity_full.json]

Runs on a B200 box (gpurun); every case asserts its tiers (bit-exact keys / correspondences / eviction streams / maps on
identical inputs, H and b to 1e-9, trajectory to 1 mm / 0.01 deg) and the summaries are written as one JSON document.
A case whose length is 0 is skipped."""
import argparse
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c1", type=int, default=1000)
    ap.add_argument("--c2", type=int, default=200)
    ap.add_argument("--c3", type=int, default=10000)
    ap.add_argument("--c4", type=int, default=50_000_000)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import parity_cases as pc
    from oracle import oracle_py
    oracle_py.build()
    cases = []
    if a.c1 > 0:
        cases += [("c1_free_running", lambda: pc.c1_free_running(oracle_py, scans=a.c1)),
                  ("c1_teacher_forced", lambda: pc.c1_teacher_forced(oracle_py, scans=a.c1))]
    if a.c2 > 0:
        cases.append(("c2_teacher_forced", lambda: pc.c2_teacher_forced(oracle_py, scans=a.c2)))
    if a.c3 > 0:
        cases.append(("c3_city_eviction", lambda: pc.c3_city_eviction(oracle_py, scans=a.c3)))
    if a.c4 > 0:
        cases.append(("c4_map_slice", lambda: pc.c4_map_slice(oracle_py, points=a.c4)))
    out = {"what": "parity of the CUDA path against the CPU oracle at the BASELINE.json configuration sizes (tests/parity_cases.py)", "cases": []}
    ok = True
    for name, fn in cases:
        t = time.time()
        try:
            r = fn()
            r["passed"] = True
        except Exception as e:          # an assertion of a tier: record it, keep going
            ok = False
            r = {"case": name, "passed": False, "error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc()[-1500:]}
        r["wall_s"] = round(time.time() - t, 1)
        print(f"[parity] {name}: {json.dumps(r)[:600]}", file=sys.stderr, flush=True)
        out["cases"].append(r)
    out["all_passed"] = ok
    txt = json.dumps(out, indent=1)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
