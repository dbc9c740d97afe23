"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export per CUDA source line.
usage: ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv > x.csv; python tools/ncu_lines.py x.csv [top]"""
import csv, sys, collections
csv.field_size_limit(1 << 30)
path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path, errors="replace")))
fname, hdr = "?", None
agg = collections.defaultdict(lambda: [0, 0, ""])     # (file, line) -> [samples, instructions, text]
stalls = collections.defaultdict(lambda: collections.Counter())
cur = None
for r in rows:
    if not r:
        continue
    if r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
        st_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":                       # a CUDA source line row
        cur = (fname, int(r[0])); agg[cur][2] = r[1].strip()[:110]; continue
    if cur is None:
        continue
    try:
        agg[cur][0] += int(r[si]); agg[cur][1] += int(r[ii])
        for i, h in st_cols:
            v = int(r[i])
            if v: stalls[cur][h] += v
    except ValueError:
        pass
tot = sum(v[0] for v in agg.values()) or 1
print(f"total samples {tot}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    s = ", ".join(f"{h[6:]}={c}" for h, c in stalls[k].most_common(3))
    print(f"{100*v[0]/tot:5.1f}%  {v[1]:>9}  {k[0]}:{k[1]:<4} {v[2]}   [{s}]")
