tag=r4h
o=gpurun_out
mkdir -p $o
(timeout 600 python bench.py --steps 20 --warmup 5 2>$o/${tag}_bench.err | tail -1) > $o/${tag}_bench.json
(timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1) > $o/${tag}_reference_arm.json
(timeout 200 python tools/layer_profile.py 2>&1) > $o/${tag}_layer_table.txt
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $o/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $o/${tag}_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc|gn_apply|gn_finalize_chsum|stem_tc|head_tc" -c 44 -o /tmp/${tag}_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $o/${tag}_ncu_full.log 2>&1
ncu -i /tmp/${tag}_full.ncu-rep --page details > $o/${tag}_full_details.txt 2>&1
ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > $o/${tag}_full_raw.csv 2>&1
du -sh $o; ls -la $o | grep r4h; cut -c1-200 $o/${tag}_bench.json
