python -m pytest tests/test_physics_gpu.py -x -q -m gpu -k "zero_copy or ragged" > gpurun_out/pytest_zc.log 2>&1; tail -3 gpurun_out/pytest_zc.log
python bench.py --no-ppo > gpurun_out/bench_zc.json 2> gpurun_out/bench_zc.err
WB_NO_ZERO_COPY=1 python bench.py --no-ppo > gpurun_out/bench_nozc.json 2> gpurun_out/bench_nozc.err
python - <<'PY'
import json
for f in ("zc","nozc"):
    d=json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["e2e"])
PY
