"""Cycle budget of ppo_tc_kernel per phase of a 64-sample tile (needs a -DWB_TC_PROFILE build:
scripts/build_variant_tc.sh tcprof -DWB_TC_PROFILE; WB_LIB_PATH=.../libwalker_b200_tcprof.so python scripts/tc_phase_profile.py)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
wb = ge.load_package(); wb.init(0)
from ppo_bipedalwalker_b200._lib import check, lib, ptr
n = 65536
rng = np.random.default_rng(0)
hp = wb.default_hyperparams(); hp.batch_size = n
agent = wb.PPOAgent(hp=hp, seed=1, stream=torch.cuda.current_stream().cuda_stream)
dev = [torch.from_numpy(x).cuda() for x in (rng.normal(size=(n, 12)).astype(np.float32), (0.3 * rng.normal(size=(n, 4))).astype(np.float32),
       (-0.5 * rng.random((n, 4))).astype(np.float32), rng.normal(size=n).astype(np.float32), rng.normal(size=n).astype(np.float32))]
L = lib()
for _ in range(3): check(L.wb_ppo_grad_dev(agent._h, n, *[ptr(t) for t in dev]))
out = (C.c_uint64 * 16)()
L.wb_tc_prof_read(out, 1)
K = 10
for _ in range(K): check(L.wb_ppo_grad_dev(agent._h, n, *[ptr(t) for t in dev]))
L.wb_tc_prof_read(out, 0)
names = ["P0 stage X + sync", "P1 F1 MMA + wait", "P2 A1/C1 epilogue + sync", "P3 F2 MMA + wait", "P4 A2 epilogue + mu", "surrogate grad + g3v + sync",
         "P5 dW3 MMA + wait", "G2 epilogue + sync", "P6 B2/dW2/dB2 MMA + wait", "P7 G1/Gc1 epilogue + sync", "P8 dW1 MMA + wait"]
ctas = int(out[11]); tiles = 1024 * K
tot = sum(int(out[i]) for i in range(11))
print(f"cycles per tile (thread 0 of each CTA), {ctas} CTA-launches, {tiles} tiles: total {tot / tiles:.0f}")
for i, nm in enumerate(names): print(f"  {nm:32s} {int(out[i]) / tiles:8.0f}  {100 * int(out[i]) / tot:5.1f}%")
print(f"per CTA-launch: prologue {int(out[12]) / ctas:.0f} cycles, after the tile loop {int(out[13]) / ctas:.0f} cycles, tiles {tot / ctas:.0f} cycles")
