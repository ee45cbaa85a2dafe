#!/usr/bin/env python
"""Key numbers of `ncu --set full` reports: python tools/ncu_summary.py gpurun_out/ncu_*.ncu-rep
(duration, DRAM bytes read/written per launch, achieved DRAM GB/s vs the measured copy peak, pipe utilisation)."""
import csv, io, json, os, subprocess, sys

WANT = {
    "gpu__time_duration.sum": "ns",
    "dram__bytes_read.sum": "rd",
    "dram__bytes_write.sum": "wr",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram%",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2%",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm%",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "hmma%",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor%",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ%",
    "launch__registers_per_thread": "regs",
    "smsp__inst_executed.sum": "inst",
    "sm__cycles_active.avg": "cycles",
}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "usecond": 1e3, "msecond": 1e6, "nsecond": 1, "second": 1e9}

def main():
    peak = 6555.8
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(path, "no data"); continue
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {}
        for h, u, v in zip(hdr, units, vals):
            if h in WANT:
                try:
                    x = float(v.replace(",", ""))
                except ValueError:
                    continue
                d[WANT[h]] = x * UNIT.get(u, 1) if WANT[h] in ("ns", "rd", "wr") else x
            if h == "Kernel Name":
                d["kernel"] = v[:60]
        ns = d.get("ns", 0)
        tot = d.get("rd", 0) + d.get("wr", 0)
        gbs = tot / ns if ns else 0
        print(f"{os.path.basename(path):44s} {d.get('kernel','')[:44]:44s} {ns/1e3:9.1f} us  dram rd {d.get('rd',0)/1e6:8.1f} MB wr {d.get('wr',0)/1e6:8.1f} MB "
              f"= {gbs:7.1f} GB/s ({100*gbs/peak:5.1f}% of {peak:.0f})  dram% {d.get('dram%',0):5.1f} l2% {d.get('l2%',0):5.1f} sm% {d.get('sm%',0):5.1f} "
              f"tensor% {d.get('tensor%', d.get('hmma%', 0)):5.1f} occ% {d.get('occ%',0):5.1f} regs {d.get('regs',0):.0f} inst {d.get('inst',0)/1e6:.2f}M")

if __name__ == "__main__":
    main()
