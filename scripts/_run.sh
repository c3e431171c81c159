P=ppo-bipedalwalker_b200/lib/libwalker_b200_prev.so
python -m pytest tests/test_physics_gpu.py -x -q -m gpu -k "kernel_variants_bit_exact or ragged" > gpurun_out/pytest_v3.log 2>&1; tail -2 gpurun_out/pytest_v3.log
for rep in 1 2; do
echo "== new"; python scripts/sweep_physics.py 262144 5 68 1001; python scripts/sweep_physics.py 65536 8 68 1001 1003
echo "== prev"; WB_LIB_PATH=$P python scripts/sweep_physics.py 262144 5 68 1001; WB_LIB_PATH=$P python scripts/sweep_physics.py 65536 8 68 1001 1003
done > gpurun_out/sweep_l.log 2>&1
cat gpurun_out/sweep_l.log
