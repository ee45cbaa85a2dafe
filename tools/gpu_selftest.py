"""Runs every per-kernel parity case (tests/kernel_cases.py) in its own subprocess on the GPU box and writes a
summary to gpurun_out/selftest.json -- one faulting kernel (trap / illegal address) cannot hide the others.
Usage: python tools/gpu_selftest.py [--only substr] [--timeout 120]"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_one(name):
    import kernel_cases as kc
    res = kc.ALL_CASES[name]()
    print("RESULT " + json.dumps({"name": name, "ok": res.ok, "items": res.items}))
    print(str(res))
    return 0 if res.ok else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case")
    ap.add_argument("--only", default="")
    ap.add_argument("--timeout", type=int, default=180)
    a = ap.parse_args()
    if a.case:
        sys.exit(run_one(a.case))
    import kernel_cases as kc
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    summary = []
    for name in kc.ALL_CASES:
        if a.only and a.only not in name:
            continue
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, "--case", name], capture_output=True, text=True, timeout=a.timeout)
            rc, out = p.returncode, p.stdout + p.stderr
        except subprocess.TimeoutExpired as e:
            rc, out = -9, "TIMEOUT " + str(e.stdout)[-2000:] + str(e.stderr)[-2000:]
        line = [l for l in out.splitlines() if l.startswith("RESULT ")]
        rec = json.loads(line[0][7:]) if line else {"name": name, "ok": False, "items": []}
        rec.update({"rc": rc, "secs": round(time.time() - t0, 1)})
        if not rec["ok"]:
            rec["tail"] = out[-1500:]
        summary.append(rec)
        print(("PASS " if rec["ok"] else "FAIL ") + name + f" rc={rc} {rec['secs']}s " +
              " ".join(f"{w}={e:.2e}" for w, e, _ in rec["items"]), flush=True)
        if not rec["ok"]:
            print("    " + out[-600:].replace("\n", "\n    "), flush=True)
    with open(os.path.join(out_dir, "selftest.json"), "w") as f:
        json.dump(summary, f, indent=1)
    n_ok = sum(r["ok"] for r in summary)
    print(f"SELFTEST {n_ok}/{len(summary)} passed")


if __name__ == "__main__":
    main()
