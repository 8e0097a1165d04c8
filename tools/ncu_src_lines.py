"""Per-source-line shares (samples, warp instructions, lanes) of an .ncu-rep captured with --import-source on:
python tools/ncu_src_lines.py report.ncu-rep [file substring] [top n]"""
import csv, subprocess, sys
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, data = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        cur = r[1]
        hdr = None
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and cur and want in cur and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":
        data.append((cur.split("/")[-1], r))
isamp, iex, ith = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
ts = sum(int(r[isamp]) for _, r in data)
te = sum(int(r[iex]) for _, r in data)
print("samples %d, warp instructions %d" % (ts, te))
data.sort(key=lambda x: -int(x[1][iex]))
for f, r in data[:top]:
    e = int(r[iex])
    print("%5.1f%% instr %5.1f%% samples  lanes %4.1f  %s:%s  %s" % (
        100.0 * e / te, 100.0 * int(r[isamp]) / max(ts, 1), int(r[ith]) / max(e, 1), f, r[0], r[1].strip()[:90]))
