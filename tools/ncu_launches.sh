# launch list of the one-worker device pass (gpu__time_duration per launch, cold caches, serialised):
# per kernel name the number of launches and the summed / average duration
set +e
mkdir -p gpurun_out
TAG=${1:-r2}
CFG=${2:-c2}
READS=${3:-100000}
DB=${4:-32000}
B="python bench.py --config $CFG --steps 1 --warmup 1 --reads $READS --threads 1 --device-block $DB --no-cli --no-cpu-baseline --parity 0"
$B > gpurun_out/ncu_plain_$TAG.json 2> gpurun_out/ncu_plain_$TAG.err || { tail -5 gpurun_out/ncu_plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_l.log 2>&1
python - "$TAG" <<'P'
import csv, collections, sys
tag = sys.argv[1]
rows=list(csv.reader(l for l in open('gpurun_out/launches_%s.csv' % tag) if l.startswith('"')))
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); iu=hdr.index('Metric Unit')
t=collections.Counter(); n=collections.Counter()
for r in rows[1:]:
    v=float(r[iv].replace(',',''))
    if r[iu]=='ns': v/=1e3
    elif r[iu]=='ms': v*=1e3
    name=r[ik].split('(')[0]
    t[name]+=v; n[name]+=1
tot=sum(t.values())
with open('gpurun_out/launches_%s_summary.txt' % tag, 'w') as f:
    for k,v in t.most_common(30):
        line="%-44s n=%4d sum %9.1f us avg %8.1f us share %5.1f %%"%(k[:44],n[k],v,v/n[k],100*v/tot)
        print(line); f.write(line+"\n")
P
