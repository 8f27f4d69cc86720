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


def test_get_state_and_gather_beyond_32767_rows(dev):
  """the gather put the row index on grid.y (<= 65535 blocks => <= 32767 rows) and `get_state()` gathered
  every live row in one call.  40k rows now round-trip through get_state()/set_state(); the uniform sampler takes
  batches above 1024 as the reference does (replay.py:76-82 has no limit)."""
  cap = 40000
  rep = dev.TransitionReplay(cap, dev.Transition(None, None, None, None, None), np.random.RandomState(2))
  dev.bulk_fill_synthetic(rep, (4, 4, 4), 13, 6)
  st = rep.get_state()
  assert len(st['storage']) == cap
  ids = np.array([i for i, _ in st['storage']])
  obs, a, r, d = replay_oracle.synthetic_rows(13, ids, 64, 6)
  got = np.stack([t.s_tm1 for _, t in st['storage']]).reshape(cap, -1)
  np.testing.assert_array_equal(got, obs[:, 0])
  np.testing.assert_array_equal(np.array([t.a_tm1 for _, t in st['storage']]), a)
  rep2 = dev.TransitionReplay(cap, dev.Transition(None, None, None, None, None), np.random.RandomState(2))
  rep2.set_state(st)
  rows = rep2.get([0, 32767, 32768, cap - 1])
  np.testing.assert_array_equal(np.stack([t.s_t for t in rows]).reshape(4, -1), obs[[0, 32767, 32768, cap - 1], 1])
  big = rep.sample(5000)
  assert big.s_tm1.shape == (5000, 4, 4, 4)
