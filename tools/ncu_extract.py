"""Summarise `ncu --set full` reports into profiles/: python tools/ncu_extract.py <round tag> name=report.ncu-rep ...

Writes profiles/<tag>_ncu_full_<name>.csv (the raw page of the report, one row per captured launch) and
merges the headline numbers into profiles/<tag>_ncu_summary.json, which bench.py reads for `roofline.traffic`
(dram__bytes_read.sum + dram__bytes_write.sum per launch)."""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg"]


def main():
    tag = sys.argv[1]
    out_json = os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.json")
    summary = json.load(open(out_json)) if os.path.exists(out_json) else {}
    for spec in sys.argv[2:]:
        name, rep = spec.split("=", 1)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_{name}.csv"), "w").write(raw)
        rows = list(csv.reader(raw.splitlines()))
        hdr, units, vals = rows[0], rows[1], rows[2]
        d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))

        def val(k):
            try:
                return float(d[k].replace(",", "")) * UNIT.get(u.get(k, ""), 1.0)
            except Exception:
                return None
        entry = {k: val(k) for k in KEYS}
        entry["kernel"] = d.get("Kernel Name"); entry["grid"] = d.get("Grid Size"); entry["block"] = d.get("Block Size")
        rd, wr = entry["dram__bytes_read.sum"], entry["dram__bytes_write.sum"]
        entry["dram_bytes_per_launch"] = (rd or 0.0) + (wr or 0.0)
        entry["what"] = "one launch over a batch of 5 frames at 1024x2048 (tools/profile_once.py 5), ncu --set full --clock-control none"
        summary[name] = entry
        print(name, entry["kernel"], "time", entry["gpu__time_duration.sum"], "dram", entry["dram_bytes_per_launch"])
    json.dump(summary, open(out_json, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
