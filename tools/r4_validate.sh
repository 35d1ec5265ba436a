#!/bin/bash
# One-GPU validation of the tree as it stands: the whole -m gpu suite, smoke(), the bench line and the reference arm with
# the driver's arguments, the stand-alone elementwise bench, the in-graph layer table and the ncu captures that
# profiles/ summarises.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-r4g}
o=gpurun_out
mkdir -p $o
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > $o/${tag}_gpu_tests.log
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4) > $o/${tag}_smoke.log
(timeout 600 python bench.py --steps 20 --warmup 5 2>$o/${tag}_bench.err | tail -1) > $o/${tag}_bench.json
(timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1) > $o/${tag}_reference_arm.json
(timeout 200 python tools/elementwise_bench.py 2>&1 | tail -12) > $o/${tag}_elementwise.txt
(timeout 200 python tools/layer_profile.py 2>&1) > $o/${tag}_layer_table.txt
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $o/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $o/${tag}_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc|gn_apply|gn_finalize_chsum|stem_tc|head_tc" -c 44 \
  -o $o/${tag}_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $o/${tag}_ncu_full.log 2>&1
# the report itself is too large to travel back (gpurun_out is capped at 64 MiB): export text, drop the binary
ncu -i $o/${tag}_full.ncu-rep --page details > $o/${tag}_full_details.txt 2>&1
ncu -i $o/${tag}_full.ncu-rep --page raw --csv > $o/${tag}_full_raw.csv 2>&1
rm -f $o/${tag}_full.ncu-rep
du -sh $o
cat $o/${tag}_gpu_tests.log $o/${tag}_smoke.log; cut -c1-200 $o/${tag}_bench.json; cut -c1-300 $o/${tag}_reference_arm.json; cat $o/${tag}_elementwise.txt; tail -3 $o/${tag}_ncu_full.log
