python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_s3.log 2>&1; tail -3 gpurun_out/pytest_gpu_s3.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_s3.log 2>&1; tail -2 gpurun_out/smoke_s3.log
python bench.py --impl reference > gpurun_out/bench_ref_s3.json 2> gpurun_out/bench_ref_s3.err
python bench.py > gpurun_out/bench_s3.json 2> gpurun_out/bench_s3.err; tail -c 300 gpurun_out/bench_s3.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_s3.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch_s3.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:physics_compact -s 66 -c 1 -f -o gpurun_out/pc_s3 python scripts/prof_physics.py 262144 3 1001 64 > gpurun_out/ncu_pc_s3.log 2>&1
python scripts/sweep_physics.py 1 30 68 0 32 > gpurun_out/sweep_o.log 2>&1; python scripts/sweep_physics.py 2048 30 68 0 16 >> gpurun_out/sweep_o.log 2>&1; cat gpurun_out/sweep_o.log
