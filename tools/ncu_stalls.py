"""Top stall locations of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name K --launch-count 1` (development aid).
usage: python tools/ncu_stalls.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
print("total samples", tot)
for key in ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_not_selected", "stall_selected", "stall_no_inst", "stall_mio", "stall_lg", "stall_branch_resolving", "stall_dispatch"]:
    print(f"{key:26s} {sum(f(r, key) for r in data) / tot * 100:6.2f} %")
top = sorted(range(len(data)), key=lambda i: -f(data[i], "# Samples"))[:n]
for i in sorted(top):
    r = data[i]
    st = {k: f(r, k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
    best = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{r[ix['Address']][-5:]} {f(r, '# Samples') / tot * 100:5.2f}%  {r[ix['Source']][:80]:80s} {best[0][0]}={best[0][1]:.0f} {best[1][0]}={best[1][1]:.0f}  exec={r[ix['Instructions Executed']]}")
