"""Extracts the headline counters of every profiles/*_raw.csv (ncu --page raw --csv export of one
launch) into profiles/SUMMARY.json, and writes the DRAM traffic of the bench kernel where
bench.py picks it up (profiles/r01_fft4096_traffic.json)."""
import csv, glob, json, os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}
out = {}
for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_raw.csv"))):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    rec = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            try:
                v = float(vals[i].replace(",", ""))
            except ValueError:
                continue
            rec[k] = v * UNIT.get(units[i], 1.0) if units[i] in UNIT else v
    if "dram__bytes_read.sum" in rec and "dram__bytes_write.sum" in rec:
        rec["dram_bytes_per_launch"] = rec["dram__bytes_read.sum"] + rec["dram__bytes_write.sum"]
    out[os.path.basename(path).replace("_raw.csv", "")] = rec
json.dump(out, open(os.path.join(ROOT, "profiles", "SUMMARY.json"), "w"), indent=1)
import subprocess
for rnd in ("r01", "r02"):
    latest = [k for k in sorted(out) if k.startswith(rnd + "_fft4096_mag")]
    if not latest:
        continue
    r = out[latest[-1]]
    rec = {"source": latest[-1] + "_raw.csv (ncu --set full, one launch, 2^27 samples)",
           "dram_bytes_per_launch": r["dram_bytes_per_launch"],
           "algorithmic_bytes_per_launch": 12 * 4096 * 32768,
           "kernel_ms_under_ncu": r.get("gpu__time_duration.sum", 0) / 1e3}
    path = os.path.join(ROOT, "profiles", rnd + "_fft4096_traffic.json")
    if rnd != "r01":
        try:    # commit whose build produced the capture: recorded next to the csv by the capture script, else HEAD
            cfile = os.path.join(ROOT, "profiles", latest[-1] + "_commit.txt")
            rec["commit"] = open(cfile).read().strip() if os.path.exists(cfile) else subprocess.check_output(
                ["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], text=True).strip()
        except Exception:
            pass
    elif os.path.exists(path):
        continue
    json.dump(rec, open(path, "w"), indent=1)
for k, r in out.items():
    print(k, {kk: (round(v, 2) if isinstance(v, float) else v) for kk, v in r.items() if kk != "kernel"})
