"""Runs the packed tcgen05 GEMM self-test on the ACTUAL operands of the IQN fc1 backward pass."""
import sys, os, ctypes as C
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT)
sys.path.insert(0, os.path.join(_ROOT, 'tests'))
import numpy as np, torch
import test_gpu_learner as T
import test_gpu_tc as G
from dqn_zoo_b200 import _lib

os.environ['DZ_PK_IQN'] = '0'   # the fp32-FMA path materialises the operands this probe feeds to the tcgen05 GEMM
spec, net, L, O, rs = T.make_case('iqn', 32, 84, seed=3)
arrs, batch, w, taus_o, taus_flat, noise_o, noise_flat = T.make_batch(spec, net, 32, rs)
L.update(*arrs, weights=w, taus=taus_flat, noise=noise_flat, apply_update=False)
torch.cuda.synchronize()

def buf(name, shape):
  p, n = C.c_void_p(), C.c_int64()
  _lib.call('dz_test_learner_buffer', L._h, name.encode(), C.byref(p), C.byref(n))
  t = torch.empty(shape, dtype=torch.float32, device='cuda')
  assert t.numel() == n.value
  C.cdll.LoadLibrary('libcudart.so').cudaMemcpy(C.c_void_p(t.data_ptr()), p, C.c_size_t(4 * n.value), 3)
  return t

hi = buf('iqn_hi', (2048, 3136)); dh1 = buf('dh1', (2048, 512))
Wf = L.view(L.online, 'fc1/w').clone()
print('hi: zeros %.3f max %.3e | dh1: zeros %.3f absmax %.3e absmean %.3e' % (
    (hi == 0).float().mean(), hi.max(), (dh1 == 0).float().mean(), dh1.abs().max(), dh1.abs().mean()))
want_w = (hi.double().T @ dh1.double()).cpu().numpy()
want_d = (dh1.double() @ Wf.double().T).cpu().numpy()
for run in ('4', '8'):
  os.environ['DZ_PK_RUN'] = run
  got = G.run_pgemm(hi.cpu().numpy(), 0, dh1.cpu().numpy(), 0, 3136, 512, 2048, 5)
  e = got - want_w
  print('RUN', run, 'wgrad rel %.3e' % G.rel(got, want_w), 'worst rows', np.argsort(-np.abs(e).max(1))[:6], 'max abs err %.3e' % np.abs(e).max(),
        'max |want| %.3e' % np.abs(want_w).max())
  got = G.run_pgemm(dh1.cpu().numpy(), 1, Wf.cpu().numpy(), 1, 2048, 3136, 512, 1)
  e = got - want_d
  print('RUN', run, 'dgrad rel %.3e' % G.rel(got, want_d), 'max abs err %.3e' % np.abs(e).max(), 'max |want| %.3e' % np.abs(want_d).max())
  # error structure: which rows/cols carry it
  print('   dgrad err by row-block of 128:', ['%.1e' % np.linalg.norm(e[i:i + 128]) for i in range(0, 2048, 256)])
