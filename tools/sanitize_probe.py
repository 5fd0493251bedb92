"""One small call of every kernel family, for `compute-sanitizer --tool memcheck python tools/sanitize_probe.py`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cognn_b200
from tests.util import power_law_csr, rand_u64, to_dev, to_np
from oracle import pyoracle as po

ctx = cognn_b200.Context(0)
rng = np.random.default_rng(0)
ok = True
for D in (3, 16, 70):
    rowptr, col = power_law_csr(rng, 700, 600, 9000)
    x, delta = rand_u64(rng, 600, D), rand_u64(rng, 700, D)
    csr = ctx.csr_create(to_dev(rowptr), to_dev(col), 600)
    ok &= np.array_equal(to_np(ctx.gather_sum(csr, to_dev(x), to_dev(delta))), po.gather_sum_csr(rowptr, col, x, delta))
    bufs = [torch.empty((350, D), dtype=torch.int64, device="cuda") for _ in range(2)]
    ctx.gather_sum_blocks(csr, to_dev(x), [b.data_ptr() for b in bufs], [0, 350, 700])
    ok &= np.array_equal(np.concatenate([to_np(b) for b in bufs]), po.gather_sum_csr(rowptr, col, x))
    ok &= np.array_equal(to_np(ctx.expand_rows(to_dev(col), to_dev(x))), po.expand_rows(col, x))
    exp = po.expand_rows(col, x)
    ok &= np.array_equal(to_np(ctx.segsum(to_dev(rowptr), to_dev(exp), True)), po.segsum(rowptr, exp, True))
    csr.destroy()
for impl, shapes in (("imad", [(130, 70, 9), (300, 1500, 33), (64, 3000, 100)]), ("tc", [(130, 70, 70), (300, 100, 128)])):
    os.environ["CGB_MATMUL_IMPL"] = impl
    for M, K, N in shapes:
        A, B = rand_u64(rng, M, K), rand_u64(rng, K, N)
        ok &= np.array_equal(to_np(ctx.matmul(to_dev(A), to_dev(B))), po.matmul(A, B))
        ok &= np.array_equal(to_np(ctx.matmul(to_dev(np.ascontiguousarray(A.T)), to_dev(B), transA=True)), po.matmul(A, B))
        U, V, Z = rand_u64(rng, M, K), rand_u64(rng, K, N), rand_u64(rng, M, N)
        ok &= np.array_equal(to_np(ctx.beaver_matmul_finish(to_dev(A), to_dev(B), to_dev(U), to_dev(V), to_dev(Z), 0, 16)),
                             po.beaver_matmul_finish(A, B, U, V, Z, 0, 16))
a, b = rand_u64(rng, 1001), rand_u64(rng, 1001)
ok &= np.array_equal(to_np(ctx.add(to_dev(a), to_dev(b))), a + b)
ok &= np.array_equal(to_np(ctx.sum_n([to_dev(a), to_dev(b), to_dev(a)])), a + b + a)
ok &= np.array_equal(to_np(ctx.scale_public(to_dev(a), 77, 1)), po.scale_public(a, 77, 16, 1))
ok &= np.array_equal(to_np(ctx.transpose(to_dev(a[:1000].reshape(40, 25)))), a[:1000].reshape(40, 25).T)
key = [1, 2, 3, 4, 5, 6, 7, 8]
ok &= np.array_equal(to_np(ctx.prg_fill(key, 3, 5, 1003)), po.prg_fill(key, 3, 5, 1003))
ok &= np.array_equal(to_np(ctx.prg_mask_sub(key, 3, 5, to_dev(a))), a - po.prg_fill(key, 3, 5, 1001))
e, fv, aa, bb, cc = rand_u64(rng, 50, 7), rand_u64(rng, 50), rand_u64(rng, 50, 7), rand_u64(rng, 50), rand_u64(rng, 50, 7)
ok &= np.array_equal(to_np(ctx.rowmul_beaver_finish(to_dev(e), to_dev(fv), to_dev(aa), to_dev(bb), to_dev(cc), 0, 16)),
                     po.rowmul_beaver_finish(e, fv, aa, bb, cc, 0, 16))
ctx.sync()
print("SANITIZE_PROBE", "OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
