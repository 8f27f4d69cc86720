"""Bring-up probe for the tcgen05 GEMM: prints relative errors for every operand-major combination
and descriptor variant (no asserts)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from dqn_zoo_b200 import _lib
import test_gpu_tc as t

rs = np.random.RandomState(0)
for (MI, NJ, R) in [(128, 64, 32), (128, 64, 64), (256, 128, 256)]:
  Am = rs.standard_normal((MI, R)).astype(np.float32)
  Bm = rs.standard_normal((NJ, R)).astype(np.float32)
  want = Am.astype(np.float64) @ Bm.astype(np.float64).T
  for variant in (0, 1):
    _lib.call('dz_test_tc_set_variant', variant)
    for tile_n in (32, 64):
      for name, (a, ak, b, bk, tr) in {'KK': (Am, 1, Bm, 1, False), 'K-MN': (Am, 1, np.ascontiguousarray(Bm.T), 0, False),
                                         'MN-K': (np.ascontiguousarray(Am.T), 0, Bm, 1, False),
                                         'MN-MN': (np.ascontiguousarray(Am.T), 0, np.ascontiguousarray(Bm.T), 0, False)}.items():
        got = t.run_tc(a, ak, b, bk, MI, NJ, R, 1, tile_n)
        err = t.rel(got, want)
        extra = ''
        if err > 1e-4:
          extra = ' got[0,:4]=%s want[0,:4]=%s nz=%d' % (np.round(got[0, :4], 3), np.round(want[0, :4], 3), int((got != 0).sum()))
        print('shape', (MI, NJ, R), 'variant', variant, 'tile', tile_n, name, 'rel=%.2e' % err, extra, flush=True)

print('--- accuracy vs reduction length (KK, tile 64)')
import torch
_lib.call('dz_test_tc_set_variant', 0)
for R in (32, 128, 256, 512, 1024, 3136):
  A = rs.standard_normal((128, R)).astype(np.float32)
  B = rs.standard_normal((64, R)).astype(np.float32)
  want = A.astype(np.float64) @ B.astype(np.float64).T
  fp32 = (torch.tensor(A) @ torch.tensor(B).T).numpy().astype(np.float64)
  for splits in (1, max(1, R // 128)):
    got = t.run_tc(A, 1, B, 1, 128, 64, R, splits, 64)
    print('R', R, 'splits', splits, 'tc rel=%.2e' % t.rel(got, want), 'torch-cpu fp32 rel=%.2e' % t.rel(fp32, want), 'mean signed rel bias=%.2e' % float(np.mean((got - want) / np.where(np.abs(want) > 1, want, np.inf))), flush=True)
