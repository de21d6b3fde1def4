cd $GRAFT_REPO_ROOT
timeout 300 python scripts/gpu/frame_share.py 1 2 4 8 2>&1 | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2bd_share8.csv python scripts/gpu/frame_share.py 8 > gpurun_out/r2bd_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2bd_share8.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
names=[(r[ki][:60], float(r[vi].replace(',',''))) for r in rows[1:]]
# last render = trailing launches; print the tail of 80 launches aggregated
tail=names[-75:]
agg=collections.OrderedDict()
for n,v in tail:
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
for n,(c,v) in agg.items(): print("%-62s %3d %9.1f us" % (n,c,v/1000))
print("sum us", sum(v for _,v in tail)/1000)
PY
