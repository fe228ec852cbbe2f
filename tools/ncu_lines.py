"""Aggregate an ncu SASS source page by CUDA source line, using nvdisasm --print-line-info of the
same cubin (ncu's CSV export of the CUDA view carries no metrics).
usage: ncu_lines.py <report.ncu-rep> <lib.so> <kernel-symbol-substring> [top]"""
import csv, os, re, subprocess, sys, tempfile

rep, so, sym = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# instruction -> line list for the function
lines, cur, infn = [], None, False
for ln in dis:
    if ln.startswith(".text."):
        infn = sym in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines.append((cur, m.group(2).strip()))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# the page holds one section per profiled launch: "Kernel Name",<name> / header row / one row per SASS instruction
sections, cur_sec = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur_sec = {"name": r[1], "rows": []}
        sections.append(cur_sec)
    elif cur_sec is not None:
        cur_sec["rows"].append(r)
sec = [x for x in sections if sym in x["name"]][0]
hdr = sec["rows"][0]
body = sec["rows"][1:]
ie, ss, at = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
assert len(body) == len(lines), (len(body), len(lines))
agg = {}
tot_i = tot_s = 0
for (loc, txt), r in zip(lines, body):
    n, s = int(r[ie] or 0), int(r[ss] or 0)
    a = agg.setdefault(loc, [0, 0, 0.0, {}])
    a[0] += n; a[1] += s; a[2] += n * float(r[at] or 0)
    for c in stall_cols:
        v = int(r[c] or 0)
        if v:
            a[3][hdr[c]] = a[3].get(hdr[c], 0) + v
    tot_i += n; tot_s += s
src_cache = {}
def src(loc):
    if not loc: return ""
    f, l = loc
    for d in ("ngs_b200/csrc", "."):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            return src_cache[p][l - 1].strip()[:90] if l - 1 < len(src_cache[p]) else ""
    return ""
print(f"total warp-instructions {tot_i}, samples {tot_s}")
print(f"{'line':>22} {'inst%':>6} {'smpl%':>6} {'thr':>5}  top stalls | source")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    st = sorted(a[3].items(), key=lambda kv: -kv[1])[:3]
    sts = " ".join(f"{k[6:]}:{100*v/max(a[1],1):.0f}" for k, v in st)
    print(f"{str(loc[0])+':'+str(loc[1]) if loc else '?':>22} {100*a[0]/tot_i:6.2f} {100*a[1]/tot_s:6.2f} {a[2]/max(a[0],1):5.1f}  {sts:40s} | {src(loc)}")
