"""Writes profiles/<name>.md and profiles/inflate_traffic.json from an `ncu --set full` report of
tools/prof_run.py (kernels: facets, inflate_decode, inflate_resolve, crc32) and its launch-list log.
usage: ncu_report.py <report.ncu-rep> <prof_run log with the stats dict> <out.md>"""
import csv, json, os, re, subprocess, sys
rep, log, out_md = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); h, u = rows[0], rows[1]
col = h.index
m = re.findall(r"'compressed_bytes': (\d+), 'inflated_bytes': (\d+)", open(log).read())
C, D = int(m[-1][0]), int(m[-1][1])
keys = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__registers_per_thread", "regs/thread"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("smsp__inst_executed.sum", "warp instructions"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads / instruction"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX % of peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak")]
names = [r[col("Kernel Name")].split("(")[0] for r in rows[2:]]
out = ["# round 1 — ncu --set full of the four heaviest kernels of one step, B200", "",
       "Command: `ncu --set full --clock-control none --import-source on -k regex:\"inflate_|facets|crc32\" -s 5 -c 4 python tools/prof_run.py 12000000 1 2`",
       f"Input: 12 M-record WGS-shaped BAM, zlib-1: C = {C} compressed bytes, D = {D} inflated bytes, 55 k BGZF blocks (one decode launch, 0.7 of a wave).", "",
       "| metric | " + " | ".join(names) + " |", "|---|" + "---:|" * len(names)]
for k, label in keys:
    i = col(k); vals = []
    for r in rows[2:]:
        v = r[i]
        try: v = f"{float(v):.4g}"
        except ValueError: pass
        vals.append(v + (" " + u[i] if label in ("time", "dram read", "dram write") else ""))
    out.append(f"| {label} | " + " | ".join(vals) + " |")
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
tr = {}
for r, n in zip(rows[2:], names):
    tr[n] = sum(float(r[col(k)]) * scale[u[col(k)]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
dec, res = tr["inflate_decode_kernel"], tr["inflate_resolve_kernel"]
out += ["", "## Reading",
        f"* **inflate_decode_kernel** (dominant): DRAM traffic {dec/1e9:.2f} GB per launch vs algorithmic C + D = {(C+D)/1e9:.2f} GB "
        f"(+ D/8 bitmap): {dec/(C+D):.2f}x (16-byte stores are half sectors: fill reads, no re-reads of the input).  Issue slots ~62 % busy, ~21 of 32",
        "  threads active per instruction: bound by instruction issue and per-warp dependency latency, not by HBM (v1, a group of 16 lanes per block,",
        "  ran 3.4 threads per instruction: round1_v1_inflate_ncu.md).",
        f"* **inflate_resolve_kernel**: {res/1e9:.2f} GB DRAM (bitmap + tokens + sources read, match bytes written); issue ~70 % busy, ~19 threads per",
        "  instruction (matches that depend on earlier matches of the same 32-token batch wait a round).",
        "* **facets_kernel**: two CTAs of 7 warps per SM (one private 15 KB quality table per warp), issue ~50 % busy.",
        "* **crc32_kernel**: per-lane replicated table, 128-bit loads: 1.5 TB/s.",
        "", "Per-line stall attribution: round1_v2_decode_lines.txt (tools/ncu_lines.py, tools/ncu_stalls.py)."]
open(out_md, "w").write("\n".join(out) + "\n")
json.dump({"source": os.path.basename(out_md) + " (ncu --set full, 12M-record input)", "inflated_bytes": D, "compressed_bytes": C,
           "decode_dram_bytes_per_launch": dec, "decode_dram_bytes_per_inflated_byte": dec / D,
           "resolve_dram_bytes_per_inflated_byte": res / D},
          open(os.path.join(os.path.dirname(out_md), "inflate_traffic.json"), "w"), indent=1)
print("\n".join(out))
