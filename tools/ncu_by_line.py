"""Per-source-line shares of an ncu capture: joins the SASS page of the report with nvdisasm's line info.
usage: python tools/ncu_by_line.py <sass_page.csv> <nvdisasm -g -c output> <kernel name substring> <source file>"""
import re, csv, collections, sys

def main():
    page, dis, kname, srcfile = sys.argv[1:5]
    lines = open(dis).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('//----') and kname in l][0]
    end = [i for i, l in enumerate(lines) if l.startswith('//----') and i > start][0]
    cur, inst = None, []
    for l in lines[start:end]:
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        if re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+.*?;', l):
            inst.append(cur)
    rows = list(csv.reader(open(page)))
    hdr, data = rows[1], rows[2:]
    isamp, iex, ith = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
    assert len(data) == len(inst), (len(data), len(inst))
    agg = collections.defaultdict(lambda: [0, 0, 0.0])
    ts = te = 0
    for k, r in zip(inst, data):
        s, e = int(r[isamp]), int(r[iex])
        a = agg[k]; a[0] += s; a[1] += e; a[2] += float(r[ith]) * e
        ts += s; te += e
    src = open(srcfile).read().splitlines()
    base = srcfile.split('/')[-1]
    print("samples %d, warp instructions %d" % (ts, te))
    for k, (s, e, th) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[5]) if len(sys.argv) > 5 else 30]:
        txt = src[k[1] - 1].strip()[:88] if k and k[0] == base else str(k)
        print("%-22s samples %5.1f%%  instr %5.1f%%  threads %4.1f  %s" % (("%s:%d" % k) if k else "-", 100 * s / ts, 100 * e / te, th / max(1, e), txt))

main()
