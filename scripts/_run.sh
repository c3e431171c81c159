P=ppo-bipedalwalker_b200/lib/libwalker_b200_prev.so
python -m pytest tests/test_physics_gpu.py -x -q -m gpu -k "kernel_variants_bit_exact or ragged or rollout_with_resets or host_pin or zero_copy" > gpurun_out/pytest_v4.log 2>&1; tail -2 gpurun_out/pytest_v4.log
for rep in 1 2; do
echo "== new"; python scripts/sweep_physics.py 4096 20 68 8 16 4; python scripts/sweep_physics.py 65536 8 68 1001 1; python scripts/sweep_physics.py 16384 10 68 4
echo "== prev"; WB_LIB_PATH=$P python scripts/sweep_physics.py 4096 20 68 8 16 4; WB_LIB_PATH=$P python scripts/sweep_physics.py 65536 8 68 1001 1;  WB_LIB_PATH=$P python scripts/sweep_physics.py 16384 10 68 4
done > gpurun_out/sweep_m.log 2>&1
cat gpurun_out/sweep_m.log
