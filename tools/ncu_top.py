"""Top lines of an `ncu --page source --csv` export by executed instructions and by stall samples."""
import csv, sys
fn = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
sec = int(sys.argv[3]) if len(sys.argv) > 3 else 0        # which kernel of a multi-kernel export
rows = list(csv.reader(open(fn)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[sec]; h = rows[hi]
end = next((i for i in range(hi + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
data = [r for r in rows[hi + 1:end] if len(r) == len(h)]
print(rows[hi - 1][:2] if hi else "")
ci = {k: h.index(k) for k in ("Source", "# Samples", "Instructions Executed", "Avg. Threads Executed")}
def num(x):
    try: return float(x)
    except Exception: return 0.0
tot_i = sum(num(r[ci["Instructions Executed"]]) for r in data); tot_s = sum(num(r[ci["# Samples"]]) for r in data)
print("total inst %.3g samples %.3g lines %d" % (tot_i, tot_s, len(data)))
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
for key in ("Instructions Executed", "# Samples"):
    print("---- top by", key)
    for r in sorted(data, key=lambda r: -num(r[ci[key]]))[:n]:
        st = sorted(((num(r[h.index(k)]), k) for k in stalls), reverse=True)[:2]
        print("%5.1f%% i %5.1f%% s thr %4.1f  %-70s %s" % (100 * num(r[ci["Instructions Executed"]]) / max(tot_i, 1), 100 * num(r[ci["# Samples"]]) / max(tot_s, 1),
              num(r[ci["Avg. Threads Executed"]]), r[ci["Source"]][:70], " ".join("%s:%d" % (k[6:], v) for v, k in st if v)))
