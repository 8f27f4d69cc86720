"""Prints the per-tensor gradient error of the IQN learner against the CPU oracle (84x84, B=32, 64 taus),
twice per configuration (bitwise repeatability), for several accumulation-run lengths / stream settings."""
import sys, os
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT)
sys.path.insert(0, os.path.join(_ROOT, 'tests'))
import numpy as np, torch
import test_gpu_learner as T

NAMES = ['conv3/w', 'embed/w', 'fc1/w', 'fc1/b', 'head/w']
for env in [{'DZ_PK_IQN': '0'}, {'DZ_PK_RUN': '8'}, {'DZ_PK_RUN': '8', 'DZ_NO_SIDE_STREAM': '1'}, {'DZ_PK_RUN': '6'},
            {'DZ_PK_RUN': '4'}, {'DZ_PK_RUN': '16'}]:
  for k in ('DZ_PK_IQN', 'DZ_PK_RUN', 'DZ_NO_SIDE_STREAM'):
    os.environ.pop(k, None)
  os.environ.update(env)
  spec, net, L, O, rs = T.make_case('iqn', 32, 84, seed=3)
  arrs, batch, w, taus_o, taus_flat, noise_o, noise_flat = T.make_batch(spec, net, 32, rs)
  loss, aux, grads = O.grads(batch, None, taus_o, noise_o)
  snaps = []
  for rep in range(3):
    L.update(*arrs, weights=w, taus=taus_flat, noise=noise_flat, apply_update=False)
    torch.cuda.synchronize()
    snaps.append({n: L.view(L.grads, n).cpu().numpy().copy() for n in L.tensors})
  same = all(np.array_equal(snaps[0][n], snaps[r][n]) for n in L.tensors for r in (1, 2))
  print(env, 'repeatable' if same else 'NOT REPEATABLE',
        ' '.join('%s %.2e' % (n, T.rel_err(snaps[0][n], grads[n].numpy())) for n in NAMES), flush=True)
  del L
