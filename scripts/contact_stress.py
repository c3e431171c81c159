"""BASELINE.json configs[4]: contact-stress sweep -- 16384 walkers, 2048 per floor material (Ice ... SuperRubber), dropped with
initial spin (omega ~ U(-5, 5) per body) and driven by random actions.  Reports, per material, contacts per env-step
(candidate pairs with AABB overlap / SAT collision / contact points, from the kernel's trace hook on a 512-walker sample that
is also compared bit for bit with the CPU oracle) and the device-resident throughput of the whole batch."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge

wb = ge.load_package(); wb.init(0)
O = ge.load_oracle()
MATS = ["Ice", "Wood", "Paper", "Titanium", "Carpet", "Rubber", "Metal", "SuperRubber"]
N, PER = 16384, 2048
floors = [MATS[i // PER] for i in range(N)]
rng = np.random.default_rng(5)

def spun(envb, n):
    f, iv = envb.get_state()
    f[:, 78:83] = rng.uniform(-5, 5, (n, 5)).astype(np.float32)
    return f, iv

# ---- parity + contact statistics on a sample: 64 walkers per material
ns = 512
sf = [MATS[i // 64] for i in range(ns)]
env = wb.EnvBatch(ns, floor_materials=sf)
ref = O.EnvBatch(ns, floor=sf)
f, iv = spun(ref, ns)
ref.set_state(f, iv); env.set_state(f, iv)
stats = {m: {"aabb": 0, "sat": 0, "points": 0, "max_pairs_in_a_substep": 0} for m in MATS}
steps = 24
exact = True
for t in range(steps):
    a = rng.uniform(-1, 1, (ns, 4)).astype(np.float32)
    env.take_actions(a); ref.take_actions(a)
    pt, jt = env.debug_contacts(wb.DT_FRAME)
    rpt, rjt = ref.step_objects(O.DT_FRAME, 50, trace=True)
    for name in pt.dtype.names:
        exact &= bool(np.array_equal(pt[name].view(np.uint32), rpt[name].view(np.uint32)))
    env.observe(); ref.observe()
    for k, m in enumerate(MATS):
        sl = slice(k * 64, (k + 1) * 64)
        stats[m]["aabb"] += int(pt["aabb"][sl].sum()); stats[m]["sat"] += int(pt["sat"][sl].sum())
        stats[m]["points"] += int(pt["ncontacts"][sl].sum())
        stats[m]["max_pairs_in_a_substep"] = max(stats[m]["max_pairs_in_a_substep"], int(pt["sat"][sl].sum(axis=2).max()))
gf, giv = env.get_state(); rf, riv = ref.get_state()
exact &= bool(np.array_equal(gf.view(np.uint32), rf.view(np.uint32)) and np.array_equal(giv, riv))
per_step = {m: {k: (v / (64 * steps) if k != "max_pairs_in_a_substep" else v) for k, v in s.items()} for m, s in stats.items()}

# ---- throughput of the whole 16384-walker batch
big = wb.EnvBatch(N, floor_materials=floors)
f, iv = big.get_state()
f[:, 78:83] = rng.uniform(-5, 5, (N, 5)).astype(np.float32)
big.set_state(f, iv)
acts = [torch.from_numpy(rng.uniform(-1, 1, (N, 4)).astype(np.float32)).cuda() for _ in range(8)]
obs = torch.empty(N, 12, device="cuda"); rew = torch.empty(N, device="cuda"); done = torch.empty(N, dtype=torch.uint8, device="cuda")
big.set_stream(torch.cuda.current_stream().cuda_stream)
for i in range(16): big.step_dev(acts[i % 8], obs, rew, done)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
K = 32
e0.record()
for i in range(K): big.step_dev(acts[i % 8], obs, rew, done)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(json.dumps({"config": "contact-stress sweep: 16384 walkers, 2048 per floor material, initial spin U(-5,5), random actions (BASELINE configs[4])",
                  "bit_exact_vs_oracle_on_512_walker_sample": exact, "sample_env_steps": steps,
                  "per_env_step": per_step, "ms_per_step": ms, "env_steps_per_s": N / ms * 1e3, "kernel_variant": big.get_variant()}))
