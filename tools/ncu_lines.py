"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line (development aid).
usage: python tools/ncu_lines.py dump.csv [kernel-substring]"""
import csv, sys, collections
def toi(x):
    try:
        return int(x)
    except ValueError:
        return 0
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path":
        fpath = r[1]; fn = rows[i + 1][1]; hdr = rows[i + 2]; i += 3
        ix = {h: k for k, h in enumerate(hdr)}
        lines = []
        while i < len(rows) and rows[i] and rows[i][0] != "File Path":
            r = rows[i]
            if r[0] != "":
                lines.append(r)
            i += 1
        if want not in fn:
            continue
        tot = sum(toi(r[ix["# Samples"]]) for r in lines) or 1
        toti = sum(toi(r[ix["Instructions Executed"]]) for r in lines) or 1
        print(f"\n== {fn[:70]}  [{fpath.split('/')[-1]}] samples={tot} inst={toti}")
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        for r in sorted(lines, key=lambda r: -toi(r[ix["# Samples"]]))[:28]:
            smp = toi(r[ix["# Samples"]]); ins = toi(r[ix["Instructions Executed"]])
            top = sorted(((toi(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:3]
            print(f"{r[0]:>5} {100*smp/tot:5.1f}%smp {100*ins/toti:5.1f}%inst  {r[1].strip()[:78]:78s} " + " ".join(f"{n}={v}" for v, n in top if v))
    else:
        i += 1
