"""Summarise an ncu gpu__time_duration launch list (csv) per kernel."""
import csv, collections, re, sys
def summarise(path, skip_first_batch=True):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for row in rows:
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        agg.setdefault(name, []).append(float(row["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    out = []
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"{k[:52]:52s} n={len(v):4d} total={sum(v)/1e3:10.1f} us mean={sum(v)/len(v)/1e3:8.1f} us share={sum(v)/tot*100:5.1f}%")
    out.append(f"TOTAL {tot/1e3:.1f} us over {len(rows)} launches")
    return "\n".join(out)
if __name__ == "__main__":
    for p in sys.argv[1:]:
        print("==", p); print(summarise(p))
