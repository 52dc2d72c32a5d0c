"""Per-CUDA-source-line summary of an ncu report: python tools/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if r[0] != "":      # a CUDA source line with aggregated metrics
        d = dict(zip(hdr, r))
        key = (cur_file, int(r[0]))
        def num(k):
            try: return float(d.get(k, "0").replace(",", ""))
            except Exception: return 0.0
        a = agg.setdefault(key, {"src": r[1].strip(), "samples": 0.0, "inst": 0.0, "tinst": 0.0})
        a["samples"] += num("# Samples"); a["inst"] += num("Instructions Executed"); a["tinst"] += num("Thread Instructions Executed")
tot_s = sum(a["samples"] for a in agg.values()) or 1; tot_i = sum(a["inst"] for a in agg.values()) or 1
print(f"total samples {tot_s:.0f}  warp-inst {tot_i:.3g}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    eff = a["tinst"] / a["inst"] if a["inst"] else 0
    print(f"{f}:{ln:<4d} samp {a['samples']/tot_s*100:5.1f}%  inst {a['inst']/tot_i*100:5.1f}%  lanes {eff:4.1f}  | {a['src'][:110]}")
