python -m pytest tests/test_physics_gpu.py -x -q -m gpu -k "kernel_variants_bit_exact or ragged" > gpurun_out/pytest_v2.log 2>&1; tail -3 gpurun_out/pytest_v2.log
python scripts/sweep_physics.py 262144 5 68 1001 1011 1012 > gpurun_out/sweep_g.log 2>&1
python scripts/sweep_physics.py 65536 8 68 1001 1011 1003 >> gpurun_out/sweep_g.log 2>&1
python scripts/sweep_physics.py 16384 10 68 4 2 8 >> gpurun_out/sweep_g.log 2>&1
python scripts/sweep_physics.py 4096 20 68 8 16 4 >> gpurun_out/sweep_g.log 2>&1
cat gpurun_out/sweep_g.log
