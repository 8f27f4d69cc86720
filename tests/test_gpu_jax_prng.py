"""GPU: jax.random.uniform on the device and the IQN agent's tau key chain against the KAT-pinned oracle."""

import ctypes as C

import numpy as np
import pytest
import torch

from oracle import jax_prng_oracle as jo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('counts', [[1], [2], [5], [2048], [2047, 3, 64], [64, 2048, 1, 7]])
def test_device_uniform_matches_oracle(counts):
  from dqn_zoo_b200 import jax_prng as jp
  rs = np.random.RandomState(len(counts))
  keys = rs.randint(0, 2 ** 32, size=(len(counts), 2), dtype=np.uint64).astype(np.uint32)
  du = jp.DeviceUniform(counts, 'cuda')
  out = torch.full((sum(counts) + 3,), -1.0, dtype=torch.float32, device='cuda')
  du.set_keys(keys)
  du.launch(out)
  torch.cuda.synchronize()
  got = out.cpu().numpy()
  off = 0
  for key, n in zip(keys, counts):
    np.testing.assert_array_equal(got[off:off + n], jo.uniform((int(key[0]), int(key[1])), (n,)))
    off += n
  assert (got[off:] == -1.0).all()


@pytest.mark.parametrize('graph', [False, True])
def test_iqn_taus_follow_the_reference_key_chain(graph):
  """update: rng_key, update_key = split(rng_key); _, k0, k1, k2 = split(update_key, 4); tau_i = uniform(k_i, (B, N)).
  act: rng_key, sample_key, _, _ = split(rng_key, 4); tau = uniform(sample_key, (1, N))   (iqn/agent.py:182-222)."""
  from dqn_zoo_b200 import agent as ag
  from dqn_zoo_b200 import learner as dl
  from dqn_zoo_b200 import parts
  from dqn_zoo_b200 import replay as dr
  B, N = 32, 64
  rep = dr.TransitionReplay(256, dr.Transition(None, None, None, None, None), np.random.RandomState(4))
  dr.bulk_fill_synthetic(rep, (84, 84, 4), 4, 6)
  agent = ag.Iqn(preprocessor=lambda ts: ts, sample_network_input=np.zeros((84, 84, 4), np.uint8), network=dl.NetworkSpec('iqn', 6),
                 optimizer=None, transition_accumulator=dr.TransitionAccumulator(), replay=rep, batch_size=B,
                 exploration_epsilon=lambda t: 0.0, min_replay_capacity_fraction=0.05, learn_period=4,
                 target_network_update_period=16, huber_param=1.0, tau_samples_policy=N, tau_samples_s_tm1=N, tau_samples_s_t=N,
                 rng_key=[0, 42], use_cuda_graph=graph, jax_prng_taus=True)
  L = agent.learner
  key = jo.prng_key(42)
  obs = np.random.RandomState(1).randint(0, 256, (84, 84, 4)).astype(np.uint8)
  for step in range(4):
    agent.learn()
    torch.cuda.synchronize()
    key, t0, t1, t2 = jo.iqn_update_taus(key, B, N, N, N)
    want = np.concatenate([t0.reshape(-1), t1.reshape(-1), t2.reshape(-1)])
    np.testing.assert_array_equal(L.taus.cpu().numpy()[:want.size], want)
    assert np.isfinite(float(L.loss.item()))
    if step == 1:       # an action selection in between advances the same key (4-way split)
      agent._act(parts.TimeStep(parts.StepType.MID, 0.0, 1.0, obs))
      key, ta = jo.iqn_act_taus(key, N)
      np.testing.assert_array_equal(L.taus.cpu().numpy()[:N], ta.reshape(-1))
  state = agent.get_state()
  assert tuple(int(v) for v in state['rng_key']['jax']) == key
