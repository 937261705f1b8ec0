"""Stall samples of one kernel grouped into phases delimited by marker instructions (BAR.SYNC, UTCBAR, SYNCS try-wait loops):
python tools/ncu_phases.py rep kernel"""
import csv, subprocess, sys, io, re
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]; body = []
for r in rows[h + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"): break
    body.append(r)
si = hdr.index("# Samples"); src = hdr.index("Source"); ie = hdr.index("Instructions Executed")
stalls = [i for i, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
tot = sum(int(r[si]) for r in body)
seg_start = 0; acc = 0; why = {}; ninst = 0
def flush(n, label):
    global acc, why, seg_start, ninst
    top = dict(sorted(why.items(), key=lambda kv: -kv[1])[:4])
    print(f"[{seg_start:5d}-{n:5d}] {acc:6d} {100*acc/tot:5.1f}%  inst/warp~{ninst}  {label:40s} {top}")
    acc = 0; why = {}; seg_start = n + 1; ninst = 0
w0 = int(body[0][ie]) or 1
for n, r in enumerate(body):
    acc += int(r[si]); ninst += round(int(r[ie]) / w0, 1)
    for i in stalls:
        v = int(r[i])
        if v: why[hdr[i][6:]] = why.get(hdr[i][6:], 0) + v
    s = r[src].strip()
    if re.search(r"BAR\.SYNC|UTCBAR|TRYWAIT|EXIT", s):
        flush(n, s[:40])
flush(len(body), "end")
