"""Share of executed instructions per innermost loop of a kernel in an `ncu --page source --csv` export (SASS view).
usage: ncu_loopshare.py source.csv [kernel section index]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1]))); sec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[sec]; h = rows[hi]
end = next((i for i in range(hi + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
data = [r for r in rows[hi + 1:end] if len(r) == len(h)]
print(rows[hi - 1][1][:90])
ia, isrc, iex, ism = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
num = lambda x: float(x) if x not in ("", None) else 0.0
addr = [int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia]) for r in data]
idx = {a: i for i, a in enumerate(addr)}
loops = []
for i, r in enumerate(data):
    m = re.search(r'BRA\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)', r[isrc])
    if m:
        t = int(m.group(1), 16)
        if t in idx and idx[t] < i: loops.append((idx[t], i))
inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
tot = sum(num(r[iex]) for r in data); tots = sum(num(r[ism]) for r in data)
acc = 0
for lo, hi2 in inner:
    ex = sum(num(r[iex]) for r in data[lo:hi2 + 1]); sm = sum(num(r[ism]) for r in data[lo:hi2 + 1])
    nmm = sum('VIADDMNMX' in r[isrc] for r in data[lo:hi2 + 1])
    if ex / max(tot, 1) < 0.003: continue
    acc += ex
    iters = num(data[hi2][iex])
    print("loop %5d..%5d  %4d instr  %3d VIADDMNMX  %5.1f%% of instr  %5.1f%% of samples  iterations %.3g" % (lo, hi2, hi2 - lo + 1, nmm, 100 * ex / tot, 100 * sm / max(tots, 1), iters))
print("inner loops listed: %.1f%% of %.3g instructions" % (100 * acc / tot, tot))
