#!/bin/bash
# ncu evidence of one round (run on a GPU box AFTER the same commands exited 0 without ncu):
#   launch list of the bench command (gpu__time_duration per launch) and one `--set full` capture of each hot kernel.
# usage: scripts/capture_profiles.sh <round tag, e.g. r2>     -> gpurun_out/ncu_<tag>/*
tag=${1:-r2}
out=gpurun_out/ncu_$tag
mkdir -p $out
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file $out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-loop > $out/launches_bench.out 2>&1
timeout 600 $NCU --set full --import-source on -k regex:physics_lanes -s 520 -c 1 -f -o $out/phys_4096 python scripts/prof_physics.py 4096 3 8 520 > $out/phys_4096.out 2>&1
timeout 600 $NCU --set full --import-source on -k regex:physics_compact -s 520 -c 1 -f -o $out/phys_262144 python scripts/prof_physics.py 262144 3 1001 520 > $out/phys_262144.out 2>&1
timeout 600 $NCU --set full --import-source on -k regex:ppo_tc_kernel -s 3 -c 1 -f -o $out/ppo_tc python scripts/prof_ppo.py 65536 3 0 > $out/ppo_tc.out 2>&1
WB_ITERATIONS=1 timeout 600 $NCU --set full --import-source on -k regex:physics_compact -s 520 -c 1 -f -o $out/phys_262144_it1 python scripts/sweep_physics.py 262144 3 517 1001 > $out/phys_262144_it1.out 2>&1
ls -la $out
