#!/bin/bash
# Build the WORKING-TREE tensor-core MLP TU with extra nvcc flags into ppo-bipedalwalker_b200/lib/libwalker_b200_<tag>.so
# usage: scripts/build_variant_tc.sh <tag> [nvcc flags, e.g. -DWB_TC_PROFILE]
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
unset CC CXX
L=ppo-bipedalwalker_b200/lib
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -ffp-contract=off -ftz=false -prec-div=true -prec-sqrt=true"
nvcc $FLAGS "$@" -c ppo-bipedalwalker_b200/csrc/mlp_tc.cu -o $L/mlp_tc_$tag.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $L/libwalker_b200_$tag.so $L/physics_lanes.o $L/api_env.o $L/physics_scene.o $L/api_scene.o \
  $L/mlp.o $L/mlp_tc_$tag.o $L/mlp_generic.o $L/api_policy.o -lcudart_static -lpthread -ldl -lrt
echo built $L/libwalker_b200_$tag.so
