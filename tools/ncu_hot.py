"""Hottest instructions (warp-stall samples) of one kernel of an .ncu-rep captured with --import-source on.
   python tools/ncu_hot.py rep.ncu-rep [kernel-id like :::2] [top N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"]
if kid:
    cmd += ["--kernel-id", kid]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; col = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
print(f"total samples {tot}")
agg = {s: sum(int(r[col[s]] or 0) for r in data) for s in stalls}
print("by reason:", ", ".join(f"{k[6:]} {v * 100 // max(tot, 1)}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 100 >= tot))
data.sort(key=lambda r: -int(r[col["# Samples"]] or 0))
for r in data[:top]:
    n = int(r[col["# Samples"]] or 0)
    why = sorted(((int(r[col[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{100 * n / tot:5.1f}%  {r[col['Source']][:110]:<110}  {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
