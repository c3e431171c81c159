set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
python - <<'PY'
import time, numpy as np, sys
import __graft_entry__ as ge
wb = ge.load_package(); wb.init(0)
import torch
for n in (4096, 16384, 65536):
    env = wb.EnvBatch(n, floor_materials="Wood")
    rng = np.random.default_rng(0)
    a = torch.from_numpy(rng.uniform(-1,1,(n,4)).astype(np.float32)).cuda()
    obs = torch.empty(n,12,device='cuda'); rew=torch.empty(n,device='cuda'); done=torch.empty(n,dtype=torch.uint8,device='cuda')
    env.set_stream(torch.cuda.current_stream().cuda_stream)
    for _ in range(5): env.step_dev(a, obs, rew, done)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    K=20
    e0.record()
    for _ in range(K): env.step_dev(a, obs, rew, done)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/K
    print(f"n={n} {ms:.3f} ms/step  {n/ms*1e3:.3e} env-steps/s  done_frac={done.float().mean().item():.3f}")
PY
