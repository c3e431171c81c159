"""List inputs where rcp_sqrt_rn differs from __frcp_rn(__fsqrt_rn(s)) (development probe)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
wb = ge.load_package(); wb.init(0)
from ppo_bipedalwalker_b200._lib import check, lib
def run(first, count):
    bad = C.c_uint64(0); fb = C.c_uint32(0)
    check(lib().wb_debug_rcp_sqrt_check(first, count, C.byref(bad), C.byref(fb)))
    return bad.value, fb.value
tot, _ = run(0, 1 << 32)
print("total mismatches", tot)
found = []
pos = 0
while len(found) < 40:
    n, fb = run(pos, (1 << 32) - pos)
    if n == 0: break
    found.append(fb)
    pos = fb + 1
for b in found:
    s = np.array([b], np.uint32).view(np.float32)[0]
    sq = np.float32(np.sqrt(np.float64(s)))
    r = np.float32(1.0 / np.float64(sq))
    print(f"bits 0x{b:08x} s={s!r} exp={(b>>23)&0xff} mant=0x{b&0x7fffff:06x} sqrt={sq!r} rcp={r!r}")
