"""Turn an .ncu-rep capture into the files kept under profiles/: the details page, the raw page as CSV, and one entry
of profiles/traffic.json (dram bytes per launch of the dominant kernel, with the capture file and the git revision it
was taken from, so that bench.py's `roofline.traffic` says where its number comes from).

    python scripts/ncu_summary.py <capture.ncu-rep> <out_prefix> [--traffic-key cfg2 --kernel csr_topk_main_kernel]
"""
import argparse
import csv
import io
import json
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("out_prefix")
ap.add_argument("--traffic-key")
ap.add_argument("--kernel", default="")
ap.add_argument("--git", default="")
args = ap.parse_args()

details = subprocess.run(["ncu", "-i", args.rep, "--page", "details"], capture_output=True, text=True).stdout
Path(args.out_prefix + "_ncu_details.txt").write_text(details)
raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
Path(args.out_prefix + "_ncu_full_raw.csv").write_text(raw)
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
data = [r for r in rows[2:] if len(r) == len(hdr)]


def col(name):
    return hdr.index(name) if name in hdr else None


out = []
for r in data:
    rec = {"kernel": r[col("Kernel Name")] if col("Kernel Name") is not None else "?"}
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
              "sm__inst_issued.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__warps_eligible.avg.per_cycle_active",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"):
        c = col(m)
        if c is not None:
            unit = rows[1][c]
            try:
                rec[m + (f" [{unit}]" if unit else "")] = float(r[c].replace(",", ""))
            except ValueError:
                rec[m] = r[c]
    out.append(rec)
print(json.dumps(out, indent=1))
if args.traffic_key and out:
    sel = [r for r in out if args.kernel in r["kernel"]] or out

    def bytes_of(rec, key):
        for k2, v in rec.items():
            if k2.startswith(key):
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                unit = k2[k2.index("[") + 1:-1] if "[" in k2 else "byte"
                return v * scale.get(unit, 1.0)
        return 0.0
    per = [bytes_of(r, "dram__bytes_read.sum") + bytes_of(r, "dram__bytes_write.sum") for r in sel]
    tj = ROOT / "profiles" / "traffic.json"
    t = json.loads(tj.read_text()) if tj.exists() else {}
    t[args.traffic_key] = {"bytes": int(sum(per) / len(per)), "launches_averaged": len(per), "kernel": sel[0]["kernel"][:120],
                           "capture": Path(args.out_prefix).name + "_ncu_full_raw.csv", "git": args.git,
                           "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none"}
    tj.write_text(json.dumps(t, indent=1) + "\n")
