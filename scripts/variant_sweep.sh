#!/bin/bash
# Kernel-variant sweep behind wb_env_create's default choice (api_env.cu: default_lanes): every batch size from the state 512 env-steps
# into a random-action rollout, all candidate variants from ONE snapshot, bit-identical end states (scripts/sweep_physics.py).
# usage: scripts/variant_sweep.sh > profiles/variant_sweep_rN.log
for spec in "1 60 32 16 8" "148 60 32 16 8" "1024 60 32 16 8" "1480 60 32 16 8" "2048 50 32 16 8 4" "3072 50 16 8 4" "4096 40 16 8 4" "6144 40 16 8 4" \
            "7104 40 8 4 2" "8192 30 8 4 2" "12288 30 8 4 2 1001" "16384 24 8 4 2 1001" "23680 20 4 2 1 1001" "28416 20 4 2 1 1001" "32768 16 4 2 1 1001" \
            "49152 12 4 2 1 1001" "65536 10 2 1 1001" "131072 6 1 1001" "262144 4 1 1001"; do
  set -- $spec
  n=$1; k=$2; shift 2
  timeout 300 python scripts/sweep_physics.py $n $k 512 "$@" 2>&1 | grep "^n="
done
