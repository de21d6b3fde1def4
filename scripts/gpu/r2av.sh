cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:k_deform_forward_umma -s 3 -c 1 -f -o /tmp/r2_deform_small python scripts/gpu/deform_small.py > /tmp/ncu_d.log 2>&1
tail -2 /tmp/ncu_d.log
ncu -i /tmp/r2_deform_small.ncu-rep --page source --print-source cuda,sass --csv > /tmp/src_d.csv 2>/dev/null
python scripts/ncu_lines.py /tmp/src_d.csv 28 > gpurun_out/r2_deform_small_hot_lines.txt 2>&1
cat gpurun_out/r2_deform_small_hot_lines.txt
ncu -i /tmp/r2_deform_small.ncu-rep --page raw --csv > /tmp/raw_d.csv 2>/dev/null
python scripts/summarize_ncu.py /tmp/raw_d.csv | tail -3
