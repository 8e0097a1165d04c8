"""top functions of the sample files written by SMALT_B200_PROF (tools/prof_host.sh)"""
import collections, sys
for fn in sys.argv[1:]:
    tot = collections.Counter()
    for ln in open(fn):
        f = ln.split()
        if len(f) < 4:
            continue
        tot[(f[1].split("/")[-1], f[3])] += int(f[0])
    n = sum(tot.values())
    print("==", fn, "samples", n)
    for (lib, name), c in tot.most_common(40):
        print("%6.2f %%  %-28s %s" % (100.0 * c / n, lib, name))
