"""Validates the tcgen05 GEMM building block (descriptor encodings, operand majorness, TMEM readback) on a B200."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
wb = ge.load_package(); wb.init(0)
from ppo_bipedalwalker_b200._lib import check, lib, ptr
rng = np.random.default_rng(0)
ok = True
for (M, N, K, amn, bmn) in [(128, 64, 64, 0, 0), (64, 64, 64, 0, 0), (64, 64, 64, 2, 2), (64, 128, 16, 2, 2), (64, 64, 64, 2, 3), (64, 64, 64, 0, 3), (64, 64, 64, 3, 3), (64, 72, 64, 3, 3), (128, 16, 64, 3, 3), (128, 8, 64, 3, 3), (64, 64, 32, 2, 0), (64, 128, 16, 0, 2), (64, 64, 64, 0, 2)]:
    for passes in (1, 3):
        A = rng.normal(size=(M, K)).astype(np.float32); B = rng.normal(size=(N, K)).astype(np.float32)
        D = np.zeros((M, N), np.float32)
        try:
            check(lib().wb_debug_tc_gemm(M, N, K, amn, bmn, passes, ptr(A), ptr(B), ptr(D)))
        except Exception as e:
            print("ERR", M, N, K, amn, bmn, passes, e); ok = False; continue
        ref = A.astype(np.float64) @ B.astype(np.float64).T
        err = np.abs(D - ref).max() / np.abs(ref).max()
        good = err < (2e-3 if passes == 1 else 4e-6)
        ok &= good
        print(f"M={M} N={N} K={K} a_mn={amn} b_mn={bmn} passes={passes}: rel err {err:.3e} {'OK' if good else 'BAD'}")
print("ALL OK" if ok else "FAILURES")
