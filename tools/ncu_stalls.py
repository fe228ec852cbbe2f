"""Total warp-stall samples by reason for one kernel of an ncu report.
usage: ncu_stalls.py <report.ncu-rep> <kernel-name-substring>"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; sections.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
sec = [x for x in sections if sys.argv[2] in x["name"]][0]
hdr, body = sec["rows"][0], sec["rows"][1:]
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = {hdr[c]: sum(int(r[c] or 0) for r in body) for c in cols}
s = sum(tot.values())
ie = hdr.index("Instructions Executed")
print("instructions", sum(int(r[ie] or 0) for r in body), "samples", s)
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        try:
            print(f"{k[6:]:24s} {100*v/s:6.2f}%")
        except BrokenPipeError:
            break
