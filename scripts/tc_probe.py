import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
wb = ge.load_package(); wb.init(0)
from ppo_bipedalwalker_b200._lib import check, lib, ptr
np.set_printoptions(linewidth=250, suppress=True)
M, K = 64, 8
A = np.zeros((M, K), np.float32); A[np.arange(8), np.arange(8)] = 1  # D[k][n] = B(n,k) as the hardware reads it
for N in (32, 64):
    B = np.zeros((N, K), np.float32)
    for bmn in (1, 0):
        for variant in (2, 3, 5, 7):
            D = np.zeros((M, N), np.float32)
            check(lib().wb_debug_tc_gemm(M, N, K, 0, bmn, 100 + variant, ptr(A), ptr(B), ptr(D)))
            print(f"N={N} b_mn={bmn} layout_type={variant-1}: float offsets read for B(n, k): rows = k (0..7), cols = n")
            print(D[:8].astype(int))
