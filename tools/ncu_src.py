"""Print the hottest SASS lines (stall samples) of one kernel from an ncu report: python tools/ncu_src.py rep kernel [N]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several kernels may be concatenated; take the first block
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"): break
    body.append(r)
si = hdr.index("# Samples"); src = hdr.index("Source"); ie = hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si]) for r in body)
print(f"{kern}: {len(body)} SASS lines, {tot} samples, {sum(int(r[ie]) for r in body)} warp-instr")
agg = {hdr[i]: sum(int(r[i]) for r in body) for i in stalls}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
for n, r in sorted(enumerate(body), key=lambda nr: -int(nr[1][si]))[:top]:
    why = {hdr[i][6:]: int(r[i]) for i in stalls if int(r[i]) > 0}
    why = dict(sorted(why.items(), key=lambda kv: -kv[1])[:3])
    print(f"{n:5d} {int(r[si]):6d} {100*int(r[si])/max(tot,1):5.1f}%  {r[src].strip()[:70]:70s} {why}")
