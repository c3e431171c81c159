#!/bin/bash
# same-box A/B: scripts/ab.sh "<tags>" <sweep args...>   (tag "cur" = the in-tree build)
tags=$1; shift
for t in $tags; do
  if [ "$t" = cur ]; then unset WB_LIB_PATH; else export WB_LIB_PATH=$PWD/ppo-bipedalwalker_b200/lib/libwalker_b200_$t.so; fi
  echo "== $t: sweep $*"
  timeout 300 python scripts/sweep_physics.py "$@" 2>&1 | tail -8
done
