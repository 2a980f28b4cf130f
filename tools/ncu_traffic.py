"""profiles/fill_traffic.json entry from an `ncu --set full --page raw --csv` capture of the fill kernels of one whole-shard ticket
(tools/ncu_fill.sh): per kernel duration, DRAM bytes, issue / pipe utilisation; the sum of the DRAM bytes is bench.py's
roofline.traffic.  usage: ncu_traffic.py raw.csv algo pairs source-note [fill_traffic.json]"""
import csv, json, sys
raw, algo, pairs, note = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
out = sys.argv[5] if len(sys.argv) > 5 else "profiles/fill_traffic.json"
rows = list(csv.reader(open(raw))); h = rows[0]
ks = []
for r in rows[2:]:
    d = dict(zip(h, r))
    f = lambda k: float(d[k].replace(",", "")) if d.get(k) not in (None, "") else 0.0
    unit = dict(zip(h, rows[1]))
    def bytes_of(k):
        v = f(k); u = unit.get(k, "byte")
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
    secs = f("gpu__time_duration.sum") * {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1, "second": 1, "nsecond": 1e-9}.get(unit.get("gpu__time_duration.sum", "ns"), 1e-9)
    if secs < 1e-4: continue                                   # empty classes
    ks.append(dict(kernel=d["Kernel Name"].split("(")[0].replace("void ", ""), seconds=secs, dram_read_bytes=bytes_of("dram__bytes_read.sum"),
                   dram_write_bytes=bytes_of("dram__bytes_write.sum"), issue_active_pct=f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   alu_pipe_pct=f("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
                   fma_pipe_pct=f("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")))
ks = ks[:len(ks) // 2] if len(ks) > 3 and len(ks) % 2 == 0 and ks[0]["kernel"] == ks[len(ks) // 2]["kernel"] else ks    # submit + rerun: keep one pass
try: allj = json.load(open(out))
except Exception: allj = {}
allj[algo] = dict(pairs=pairs, seed=1, scorefn="distance", kernels=ks,
                  dram_bytes_per_step=sum(k["dram_read_bytes"] + k["dram_write_bytes"] for k in ks), source=note)
json.dump(allj, open(out, "w"), indent=1)
print(algo, "kernels", len(ks), "dram GB", allj[algo]["dram_bytes_per_step"] / 1e9, "ms", [round(k["seconds"] * 1e3, 2) for k in ks])
