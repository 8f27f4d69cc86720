"""GPU: the replay path at BASELINE.json's full size (capacity 1,000,000, batch 32) through
size-independent properties and a step-by-step comparison with the oracle:

  * bulk-filled device state == closed-form oracle state (ids, tree root, sampled ids bit-exact),
  * after 150 fused learner steps with priority write-back the 16 MiB device sum tree is bit-identical to
    the oracle's and internally consistent (every node == fl(left + right)),
  * uniform replay at 1M: sampled ids bit-exact, gathered rows byte-identical to the synthetic generator.

Observations are 44x44x4 (7.7 KB) instead of 84x84x4 so the 1M-row store is 15.5 GB, not 56 GB; nothing
in the replay path depends on the row length.
"""

import numpy as np
import pytest
import torch

from oracle import cpu_reference
from oracle import replay_oracle as ro

pytestmark = pytest.mark.gpu

CAP = 1000000
OBS = (44, 44, 4)


def test_prioritized_1m_fused_steps_match_oracle():
  from dqn_zoo_b200 import agent as ag
  from dqn_zoo_b200 import learner as dl
  from dqn_zoo_b200 import replay as dr
  seed = 3
  beta = lambda t: 0.5
  rep = dr.PrioritizedTransitionReplay(CAP, dr.Transition(None, None, None, None, None), 0.5, beta, 1e-3, True,
                                       np.random.RandomState(seed))
  dr.bulk_fill_synthetic(rep, OBS, seed, 6)
  assert rep.size == CAP and rep.capacity == CAP
  orep, _ = cpu_reference.build_replay('rainbow', CAP, 32, seed, obs_shape=OBS)
  orep._beta = beta
  # closed-form bookkeeping agrees
  st = rep._distribution.get_state()
  assert st['active_indices'][:3] == [CAP - 1, CAP - 2, CAP - 3] and st['id_to_index'][0] == CAP - 1
  tree0 = rep._distribution._sum_tree.get_state()['storage']
  assert tree0[1] == float(CAP) and tree0.shape[0] == 2 * (1 << 20)
  net = dl.NetworkSpec('rainbow', 6, obs_shape=OBS)
  agent = ag.Rainbow(preprocessor=lambda ts: ts, sample_network_input=np.zeros(OBS, np.uint8), network=net,
                     support=np.linspace(-10, 10, 51), optimizer=None,
                     transition_accumulator=dr.NStepTransitionAccumulator(3), replay=rep, batch_size=32,
                     min_replay_capacity_fraction=0.02, learn_period=16, target_network_update_period=32000,
                     rng_key=[0, seed], use_cuda_graph=True)
  L = agent.learner
  for step in range(150):
    agent.learn()
    ids, probs, w = orep.sample_ids(32)
    pri = L.priorities.cpu().numpy()            # synchronises
    np.testing.assert_array_equal(L.sampled_ids.cpu().numpy(), ids, err_msg='step %d' % step)
    np.testing.assert_allclose(L.sampled_weights.cpu().numpy(), w, rtol=1e-14)
    assert np.isfinite(pri).all() and (pri >= 0).all() and (pri <= 100).all()
    orep.update_priorities(ids, pri)
  agent.check_device_flags()
  tree = rep._distribution._sum_tree.get_state()['storage']
  otree = orep.get_state()['distribution']['sum_tree']['storage']
  np.testing.assert_array_equal(tree, otree)
  fl = 1 << 20
  assert np.array_equal(tree[1:fl], tree[2:2 * fl:2] + tree[3:2 * fl:2])     # every node = fl(left + right)
  assert agent.max_seen_priority >= float(pri.max())
  # gathered rows are the synthetic generator's bytes
  tr, ids2, _ = rep.sample(8)
  obs, a, r, d = ro.synthetic_rows(seed, ids2, int(np.prod(OBS)), 6)
  np.testing.assert_array_equal(tr.s_tm1.reshape(8, -1), obs[:, 0])
  np.testing.assert_array_equal(tr.s_t.reshape(8, -1), obs[:, 1])
  np.testing.assert_array_equal(tr.a_tm1, a)


def test_uniform_1m_sampling_matches_oracle_and_wraps():
  from dqn_zoo_b200 import replay as dr
  seed = 5
  rep = dr.TransitionReplay(CAP, dr.Transition(None, None, None, None, None), np.random.RandomState(seed))
  dr.bulk_fill_synthetic(rep, (8, 8, 4), seed, 6)
  orep, _ = cpu_reference.build_replay('dqn', CAP, 32, seed, obs_shape=(8, 8, 4))
  for _ in range(20):
    ids_d, _, _ = rep.sample_device(32)
    np.testing.assert_array_equal(ids_d.cpu().numpy(), orep.sample_ids(32))
  # 300 more adds wrap the ring: oldest ids leave, the swap-remove permutation follows the reference
  rs = np.random.RandomState(9)
  for k in range(300):
    o = rs.randint(0, 256, (8, 8, 4)).astype(np.uint8)
    item = (o, int(k % 6), float(k), 0.99, o)
    rep.add(dr.Transition(*item))
    orep.add(ro.Transition(*item))
  assert rep.size == CAP and list(rep.ids())[:2] == [300, 301]
  for _ in range(20):
    ids_d, _, _ = rep.sample_device(32)
    np.testing.assert_array_equal(ids_d.cpu().numpy(), orep.sample_ids(32))
  ok, msg = rep.check_valid()
  assert ok, msg


def test_prioritized_1m_at_the_baseline_geometry_84x84x4():
  """BASELINE.json configs[1] as the bench runs it: capacity 1M of 84x84x4 observations = 56.4 GB of HBM, row
  offsets up to slot * 2 * 28224 = 5.6e10 (> 2^32).  Checks the bytes the store hands out for slots on both sides of
  every 32-bit boundary (replay.py:706-723 `get`), the fused sampler's ids/weights against the oracle, and that the
  row-pointer table the learner reads in place (conv1 operand gather) resolves to those same bytes."""
  from dqn_zoo_b200 import agent as ag
  from dqn_zoo_b200 import learner as dl
  from dqn_zoo_b200 import replay as dr
  free, _ = torch.cuda.mem_get_info()
  if free < 62 * (1 << 30):
    pytest.skip('needs 62 GB of free HBM')
  obs_shape = (84, 84, 4)
  nbytes = int(np.prod(obs_shape))
  seed = 4
  beta = lambda t: 0.5
  rep = dr.PrioritizedTransitionReplay(CAP, dr.Transition(None, None, None, None, None), 0.5, beta, 1e-3, True,
                                       np.random.RandomState(seed))
  dr.bulk_fill_synthetic(rep, obs_shape, seed, 6)
  assert rep.size == CAP
  stride = rep._store.obs_stride if hasattr(rep._store, 'obs_stride') else nbytes
  # slots whose byte offset straddles k * 2^32 for every k the store reaches, plus both ends
  ids = [0, 1, CAP - 1, CAP - 2]
  k = 1
  while k * (1 << 32) < CAP * 2 * stride:
    s = (k * (1 << 32)) // (2 * stride)
    ids += [s - 1, s, s + 1]
    k += 1
  ids = np.array(sorted(set(i for i in ids if 0 <= i < CAP)), dtype=np.int64)
  assert len(ids) > 30
  got = rep.get(ids)
  obs, a, r, d = ro.synthetic_rows(seed, ids, nbytes, 6)
  np.testing.assert_array_equal(np.stack([t.s_tm1 for t in got]).reshape(len(ids), -1), obs[:, 0])
  np.testing.assert_array_equal(np.stack([t.s_t for t in got]).reshape(len(ids), -1), obs[:, 1])
  np.testing.assert_array_equal(np.array([t.a_tm1 for t in got]), a)
  np.testing.assert_array_equal(np.array([t.r_t for t in got]), r)
  # fused learner steps at this geometry: ids / weights bit-exact vs the oracle, tree stays consistent
  orep, _ = cpu_reference.build_replay('rainbow', CAP, 32, seed, obs_shape=obs_shape)
  orep._beta = beta
  net = dl.NetworkSpec('rainbow', 6, obs_shape=obs_shape)
  agent = ag.Rainbow(preprocessor=lambda ts: ts, sample_network_input=np.zeros(obs_shape, np.uint8), network=net,
                     support=np.linspace(-10, 10, 51), optimizer=None,
                     transition_accumulator=dr.NStepTransitionAccumulator(3), replay=rep, batch_size=32,
                     min_replay_capacity_fraction=0.02, learn_period=16, target_network_update_period=32000,
                     rng_key=[0, seed], use_cuda_graph=True)
  L = agent.learner
  for step in range(40):
    agent.learn()
    ids_o, _, w = orep.sample_ids(32)
    pri = L.priorities.cpu().numpy()
    np.testing.assert_array_equal(L.sampled_ids.cpu().numpy(), ids_o, err_msg='step %d' % step)
    np.testing.assert_allclose(L.sampled_weights.cpu().numpy(), w, rtol=1e-14)
    assert np.isfinite(pri).all()
    orep.update_priorities(ids_o, pri)
  agent.check_device_flags()
  # the sampled batch through the public API: bytes of rows anywhere in the 56 GB store
  tr, ids2, _ = rep.sample(32)
  obs2, a2, _, _ = ro.synthetic_rows(seed, ids2, nbytes, 6)
  np.testing.assert_array_equal(tr.s_tm1.reshape(32, -1), obs2[:, 0])
  np.testing.assert_array_equal(tr.s_t.reshape(32, -1), obs2[:, 1])
  np.testing.assert_array_equal(tr.a_tm1, a2)
  tree = rep._distribution._sum_tree.get_state()['storage']
  np.testing.assert_array_equal(tree, orep.get_state()['distribution']['sum_tree']['storage'])
