OUT=gpurun_out/p1; mkdir -p $OUT
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_c5.csv python tools/profile_step.py c5 0 1 > $OUT/ncu_c5.log 2>&1
python - <<PY
import csv, collections
rows=list(csv.reader(open("$OUT/launches_c5.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hi]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[hi+1:]:
    if len(r)<=vi: continue
    n=r[ki].split("(")[0][:60]; t=float(r[vi].replace(",",""))
    tot[n]+=t; cnt[n]+=1
s=sum(tot.values())
for n,t in tot.most_common(): print("%-60s %6d launches %10.1f us %5.1f%%"%(n,cnt[n],t/1000 if t>1e6 else t, 100*t/s))
print("total", s)
PY
