"""Smallest possible run of the tcgen05 limb matmul (one tile, one k-step), printed against exact integer arithmetic."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CGB_MATMUL_IMPL"] = "tc"
import numpy as np
import torch

import cognn_b200

M, K, N = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (128, 32, 64))]
ctx = cognn_b200.Context(0)
rng = np.random.default_rng(1)
mode = sys.argv[4] if len(sys.argv) > 4 else "rand"
if mode == "small":
    A = rng.integers(0, 3, size=(M, K), dtype=np.uint64)
    B = rng.integers(0, 3, size=(K, N), dtype=np.uint64)
else:
    A = rng.integers(0, 1 << 64, size=(M, K), dtype=np.uint64)
    B = rng.integers(0, 1 << 64, size=(K, N), dtype=np.uint64)
want = np.zeros((M, N), dtype=np.uint64)
for k in range(K):
    want += A[:, k:k + 1] * B[k:k + 1, :]
dA = torch.from_numpy(A.view(np.int64)).cuda()
dB = torch.from_numpy(B.view(np.int64)).cuda()
got = ctx.matmul(dA, dB).cpu().numpy().view(np.uint64)
ok = np.array_equal(got, want)
print("tc_probe", M, K, N, mode, "OK" if ok else "MISMATCH", "mismatches:", int((got != want).sum()))
if not ok:
    idx = np.argwhere(got != want)[:5]
    for i, j in idx:
        print(" at", i, j, "got", hex(int(got[i, j])), "want", hex(int(want[i, j])))
sys.exit(0 if ok else 1)
