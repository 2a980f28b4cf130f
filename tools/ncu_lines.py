"""Instructions and stall samples per CUDA source line from an `ncu --page source --csv --print-source cuda,sass` export."""
import csv, sys
fn = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(fn)))
agg = {}; cur_file = ""; h = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": h = r; ii = h.index("Instructions Executed"); si = h.index("# Samples"); continue
    if h is None or len(r) != len(h): continue
    if r[0] != "":      # a CUDA line row (aggregated over its SASS)
        try: agg[(cur_file, int(r[0]), r[1].strip()[:110])] = (float(r[ii]), float(r[si]))
        except ValueError: pass
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("total inst %.3g samples %.3g" % (ti, ts))
for (f, ln, src), (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    print("%5.1f%% i %5.1f%% s  %s:%d  %s" % (100 * i / ti, 100 * s / max(ts, 1), f, ln, src))
