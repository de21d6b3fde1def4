# final round-2 evidence: training-step launch list + --set full of one step (r2x.sh), then the launch list of one 800x800 frame
bash scripts/gpu/r2x.sh > gpurun_out/r2bh_x.log 2>&1
tail -5 gpurun_out/r2bh_x.log
cd $GRAFT_REPO_ROOT
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:k_march_round|k_composite_round|k_deform_forward|k_grid_forward|k_heads_forward|k_near_far|k_occupancy' --csv --log-file gpurun_out/r2_frame_launches.csv python scripts/frame_once.py > /tmp/ncu3.log 2>&1
tail -2 /tmp/ncu3.log
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_frame_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
names=[(r[ki].split('(')[0][:70], float(r[vi].replace(',',''))) for r in rows[1:]]
half=names[len(names)//2:]
agg=collections.OrderedDict()
for n,v in half:
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
for n,(c,v) in agg.items(): print("%-72s %3d %9.1f us" % (n,c,v/1000))
print("sum us", round(sum(v for _,v in half)/1000,1))
PY
