"""functions and lines of the samples of SMALT_B200_PROF files, static functions resolved with addr2line:
python tools/prof_resolve.py gpurun_out/prof_host_flat.txt [nfun [nline]]"""
import collections, subprocess, sys, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fn = sys.argv[1]
nfun = int(sys.argv[2]) if len(sys.argv) > 2 else 45
nline = int(sys.argv[3]) if len(sys.argv) > 3 else 30
per_lib = collections.defaultdict(collections.Counter)
other = collections.Counter()
for ln in open(fn):
    f = ln.split()
    if len(f) < 4:
        continue
    lib = f[1].split("/")[-1]
    if lib in ("libsmalt_b200_map.so", "libsmalt_b200.so"):
        per_lib[lib][f[2]] += int(f[0])
    else:
        other[(lib, f[3])] += int(f[0])
funs, lines = collections.Counter(), collections.Counter()
for lib, tot in per_lib.items():
    addrs = list(tot)
    out = subprocess.run(["addr2line", "-f", "-e", os.path.join(root, "smalt_b200", lib)] + addrs,
                         capture_output=True, text=True).stdout.split("\n")
    for i, a in enumerate(addrs):
        funs[(lib, out[2 * i])] += tot[a]
        lines[(out[2 * i], out[2 * i + 1].split("/")[-1])] += tot[a]
for k, c in other.items():
    funs[k] += c
n = sum(funs.values())
print("samples", n)
for (lib, name), c in funs.most_common(nfun):
    print("%5.1f%%  %-24s %s" % (100.0 * c / n, lib, name[:90]))
print()
for k, c in lines.most_common(nline):
    print("%5.1f%%  %s" % (100.0 * c / n, k))
