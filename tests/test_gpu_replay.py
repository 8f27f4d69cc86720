"""GPU: the CUDA-backed replay vs the golden vectors made from the reference's replay.py, the
reference's known-answer tables, and the oracle on longer random runs (bit-exact ids/indices/
probabilities; importance weights within 4 ulp because they go through pow())."""

import numpy as np
import pytest

from oracle import replay_oracle, scenarios
import replay_contract as rc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
  from dqn_zoo_b200 import replay
  return replay


@pytest.mark.parametrize('name', list(scenarios.ALL))
def test_device_reproduces_reference_golden(dev, name):
  rc.check_scenario(dev, name, 'device')


@pytest.mark.parametrize('fn', rc.CONTRACT, ids=lambda f: f.__name__)
def test_device_contract(dev, fn):
  fn(dev)


def test_device_matches_oracle_on_long_per_run(dev):
  res = {}
  for name, lib in (('oracle', replay_oracle), ('device', dev)):
    res[name] = scenarios.prioritized_replay_script(lib, capacity=1000, alpha=0.5, usp=1e-3, normalize=True, batch=32,
                                                    rounds=150, seed=41)
  for k in res['oracle']:
    if k == 'weights':
      np.testing.assert_allclose(res['device'][k], res['oracle'][k], rtol=1e-15, atol=0)
    else:
      np.testing.assert_array_equal(res['device'][k], res['oracle'][k], err_msg=k)


def test_synthetic_fill_matches_oracle_rows(dev):
  import ctypes as C
  import torch
  from dqn_zoo_b200 import _lib
  rep = dev.PrioritizedTransitionReplay(4096, dev.Transition(None, None, None, None, None), 0.5, lambda t: 0.5, 1e-3,
                                        True, np.random.RandomState(1))
  rep._store.allocate((84, 84, 4), np.uint8)
  v = rep.device_view()
  _lib.call('dz_replay_fill_synthetic', C.byref(v), 0, 4096, 7, 6, 0.99, torch.cuda.current_stream().cuda_stream)
  rows = np.array([0, 1, 17, 4095])
  obs, a, r, d = replay_oracle.synthetic_rows(7, rows, 84 * 84 * 4, 6)
  got = rep._store.obs[torch.as_tensor(rows, device='cuda')].cpu().numpy()
  np.testing.assert_array_equal(got[:, :, :84 * 84 * 4], obs)
  np.testing.assert_array_equal(rep._store.action[torch.as_tensor(rows, device='cuda')].cpu().numpy(), a)
  np.testing.assert_array_equal(rep._store.reward[torch.as_tensor(rows, device='cuda')].cpu().numpy(), r)
  np.testing.assert_array_equal(rep._store.discount[torch.as_tensor(rows, device='cuda')].cpu().numpy(), d)
