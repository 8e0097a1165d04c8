set +e
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --reads 100000 --threads 1 --no-cli --no-cpu-baseline --no-paired"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_tmp.csv $B > gpurun_out/ncu_l.log 2>&1
python - <<'P'
import csv, collections
rows=list(csv.reader(l for l in open('gpurun_out/launches_tmp.csv') if l.startswith('"')))
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); iu=hdr.index('Metric Unit')
t=collections.Counter(); n=collections.Counter()
for r in rows[1:]:
    v=float(r[iv].replace(',',''))
    if r[iu]=='ns': v/=1e3
    elif r[iu]=='ms': v*=1e3
    name=r[ik].split('(')[0]
    t[name]+=v; n[name]+=1
for k,v in t.most_common(6): print("%-40s n=%3d avg %.1f us"%(k[:40],n[k],v/n[k]))
P
