"""Cycle budget of the compacting physics kernel per queue round (needs a -DWB_PHASE_PROFILE build:
scripts/build_variant.sh prof -DWB_PHASE_PROFILE; WB_LIB_PATH=.../libwalker_b200_prof.so python scripts/phase_profile.py N WARM K)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge

n, warm, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
wb = ge.load_package()
wb.init(0)
L = wb.lib()
env = wb.EnvBatch(n, floor_materials="Wood")
env.set_variant(int(os.environ.get("WB_VARIANT", "1001")))
rng = np.random.default_rng(0)
acts = [torch.from_numpy(rng.uniform(-1, 1, (n, 4)).astype(np.float32)).cuda() for _ in range(8)]
obs = torch.empty(n, 12, device="cuda")
rew = torch.empty(n, device="cuda")
done = torch.empty(n, dtype=torch.uint8, device="cuda")
env.set_stream(torch.cuda.current_stream().cuda_stream)
for i in range(warm):
    env.step_dev(acts[i % 8], obs, rew, done)
out = (C.c_uint64 * 24)()
L.wb_prof_read(out, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(K):
    env.step_dev(acts[i % 8], obs, rew, done)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
L.wb_prof_read(out, 0)
st1, w1, dr, w2, warps, items, rounds = [int(out[i]) for i in range(7)]
tot = st1 + w1 + dr + w2
ctas = warps / 8.0
print(f"n={n} {ms:.3f} ms/step; per warp-launch cycles: total {tot / warps:.0f}  stage1 {st1 / warps:.0f} ({100 * st1 / tot:.1f}%)  "
      f"wait-before-drain {w1 / warps:.0f} ({100 * w1 / tot:.1f}%)  drain {dr / warps:.0f} ({100 * dr / tot:.1f}%)  wait-after-drain {w2 / warps:.0f} ({100 * w2 / tot:.1f}%)")
print(f"rounds per CTA-launch {rounds / ctas:.1f}, items per round {items / max(rounds, 1):.1f}, cycles per round {tot / warps / (rounds / ctas):.0f}")
for k, name in ((0, "pole"), (1, "floor")):
    sat, rest, calls, packed = [int(out[8 + 4 * k + i]) for i in range(4)]
    if calls:
        lanes, coll = packed & 0xFFFFFFFF, packed >> 32
        print(f"drain {name}: {calls / K:.0f} warp-calls per launch, SAT {sat / calls:.0f} cycles, contacts+moves+impulses {rest / calls:.0f} cycles, "
              f"active lanes per call {lanes / calls:.1f}, colliding lanes per call {coll / calls:.1f}")
