# round-2 ncu evidence of the training step (eager launches, one GPU): launch list of ~6 steps + --set full of one whole step.
# The .ncu-rep stays in /tmp on the box (it exceeds what gpurun carries back); the raw CSV pages and summaries travel.
cd $GRAFT_REPO_ROOT
TAG=r2
OUT=gpurun_out
K='regex:k_select_frame|k_step_begin|k_march|k_deform|k_grid|k_heads|k_composite|k_wgrad|k_adam|k_mlp_tail|k_grad_finite|k_cast_pad|k_pack_umma|k_loss_scale|k_mse'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file $OUT/${TAG}_launches_train_steps.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > /tmp/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" -s ${NCU_SKIP:-60} -c ${NCU_COUNT:-22} -o /tmp/${TAG}_train -f python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extras > /tmp/ncu2.log 2>&1
tail -1 /tmp/ncu2.log | cut -c1-200
ncu -i /tmp/${TAG}_train.ncu-rep --page raw --csv > $OUT/${TAG}_train_raw.csv 2>/dev/null
python scripts/summarize_ncu.py $OUT/${TAG}_train_raw.csv > $OUT/${TAG}_train_step_ncu_full_summary.md
cat $OUT/${TAG}_train_step_ncu_full_summary.md | head -30
# source-level hot spots of the two heaviest kernels of our own (wgrad, deform backward)
ncu -i /tmp/${TAG}_train.ncu-rep --page source --print-source cuda,sass --csv -k regex:k_wgrad_umma > /tmp/src_wgrad.csv 2>/dev/null
python scripts/ncu_lines.py /tmp/src_wgrad.csv 12 > $OUT/${TAG}_wgrad_hot_lines.txt 2>&1
head -14 $OUT/${TAG}_wgrad_hot_lines.txt
ls -la $OUT/ | grep r2_
