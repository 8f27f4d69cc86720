"""GPU: agents through the reference surface (parts.run_loop / step / get_state / set_state) and
the fused `_learn()` against the oracle replay + oracle learner, step by step."""

import copy

import numpy as np
import pytest
import torch

from oracle import learner_oracle as lo
from oracle import replay_oracle as ro

pytestmark = pytest.mark.gpu

OBS = (84, 84, 4)


def _unpack_noise(net, flat):
  from dqn_zoo_b200 import learner as dl
  out, pos = [], 0
  for _ in range(3):
    one = {}
    for name, n in dl.noise_vector_sizes(net):
      one[name] = torch.tensor(flat[pos:pos + n].copy())
      pos += (n + 3) // 4 * 4
    out.append(one)
  return out


def _oracle_replay_like(kind, cap, seed, prioritized, alpha, beta_fn):
  structure = ro.Transition(None, None, None, None, None)
  rs = np.random.RandomState(seed)
  if prioritized:
    rep = ro.PrioritizedTransitionReplay(cap, structure, alpha, beta_fn, 1e-3, True, rs)
  else:
    rep = ro.TransitionReplay(cap, structure, rs)
  obs, a, r, d = ro.synthetic_rows(seed, np.arange(cap), int(np.prod(OBS)), 6)
  for i in range(cap):
    item = ro.Transition(obs[i, 0].reshape(OBS), int(a[i]), float(r[i]), float(d[i]), obs[i, 1].reshape(OBS))
    rep.add(item, 1.0) if prioritized else rep.add(item)
  return rep


def _make_agent(kind, rep, seed, graph, **over):
  from dqn_zoo_b200 import agent as ag
  from dqn_zoo_b200 import learner as dl
  from dqn_zoo_b200 import replay as dr
  net = dl.NetworkSpec(kind, 6)
  common = dict(preprocessor=lambda ts: ts, sample_network_input=np.zeros(OBS, np.uint8), network=net, optimizer=None,
                transition_accumulator=dr.NStepTransitionAccumulator(3 if kind == 'rainbow' else 1), replay=rep,
                batch_size=32, min_replay_capacity_fraction=0.05, learn_period=4, target_network_update_period=16,
                rng_key=[0, seed], use_cuda_graph=graph)
  common.update(over)
  if kind == 'rainbow':
    return ag.Rainbow(support=np.linspace(-10, 10, 51), **common)
  if kind == 'iqn':
    return ag.Iqn(exploration_epsilon=lambda t: 0.1, huber_param=1.0, tau_samples_policy=64, tau_samples_s_tm1=64,
                  tau_samples_s_t=64, **common)
  return ag.AGENTS[kind](exploration_epsilon=lambda t: 0.1, grad_error_bound=1.0 / 32, **common)


@pytest.mark.parametrize('graph', [False, True])
def test_fused_rainbow_learn_matches_oracle_step_by_step(graph):
  from dqn_zoo_b200 import replay as dr
  cap, seed = 1024, 5
  beta = lambda t: 0.55
  rep = dr.PrioritizedTransitionReplay(cap, dr.Transition(None, None, None, None, None), 0.5, beta, 1e-3, True,
                                       np.random.RandomState(seed))
  dr.bulk_fill_synthetic(rep, OBS, seed, 6)
  agent = _make_agent('rainbow', rep, seed, graph)
  orep = _oracle_replay_like('rainbow', cap, seed, True, 0.5, beta)
  L = agent.learner
  spec = lo.NetSpec('rainbow', 6)
  O = lo.Learner(spec, L.get_params('online'), dtype=torch.float64)
  np.testing.assert_array_equal(rep.get_state()['distribution']['sum_tree']['storage'],
                                orep.get_state()['distribution']['sum_tree']['storage'])
  max_seen = 1.0
  for step in range(6):
    agent.learn()
    torch.cuda.synchronize()
    ids, probs, w = orep.sample_ids(32)
    np.testing.assert_array_equal(L.sampled_ids.cpu().numpy(), ids)                 # bit-exact ids
    np.testing.assert_allclose(L.sampled_weights.cpu().numpy(), w, rtol=1e-14)
    tr = ro._stack_fields(orep._structure, orep.get(ids))
    noise = _unpack_noise(L.net, L.noise.cpu().numpy())
    aux = O.update(lo.batch_from_numpy(*tr), torch.as_tensor(w), None, noise)
    assert abs(float(L.loss.item()) - float(aux['loss'])) <= 1e-4 * abs(float(aux['loss'])), step
    pri = L.priorities.cpu().numpy()
    np.testing.assert_allclose(pri, aux['priorities'].numpy(), rtol=2e-4, atol=1e-6)
    orep.update_priorities(ids, pri)   # same float32 priorities -> index path stays bit-comparable
    np.testing.assert_array_equal(rep.get_state()['distribution']['sum_tree']['storage'],
                                  orep.get_state()['distribution']['sum_tree']['storage'])
    max_seen = max(max_seen, float(pri.max()))
    assert agent.max_seen_priority == pytest.approx(max_seen, rel=0, abs=0)
  agent.check_device_flags()


def test_fused_dqn_uniform_ids_match_oracle():
  from dqn_zoo_b200 import replay as dr
  cap, seed = 777, 9
  rep = dr.TransitionReplay(cap, dr.Transition(None, None, None, None, None), np.random.RandomState(seed))
  dr.bulk_fill_synthetic(rep, OBS, seed, 6)
  agent = _make_agent('dqn', rep, seed, True)
  orep = _oracle_replay_like('dqn', cap, seed, False, 0.0, None)
  L = agent.learner
  O = lo.Learner(lo.NetSpec('dqn', 6), L.get_params('online'), dtype=torch.float64)
  for step in range(4):
    agent.learn()
    torch.cuda.synchronize()
    ids = orep.sample_ids(32)
    np.testing.assert_array_equal(L.sampled_ids.cpu().numpy(), ids)
    tr = ro._stack_fields(orep._structure, orep.get(ids))
    aux = O.update(lo.batch_from_numpy(*tr))
    assert abs(float(L.loss.item()) - float(aux['loss'])) <= 1e-4 * abs(float(aux['loss'])) + 1e-7


class _Env:
  """Deterministic dummy environment: random uint8 frames, episodes of 9..17 steps."""

  def __init__(self, seed):
    self.rs = np.random.RandomState(seed)
    self.left = 0

  def _obs(self):
    return self.rs.randint(0, 256, OBS).astype(np.uint8)

  def reset(self):
    from dqn_zoo_b200 import parts
    self.left = int(self.rs.randint(9, 18))
    return parts.TimeStep(parts.StepType.FIRST, None, None, self._obs())

  def step(self, action):
    from dqn_zoo_b200 import parts
    self.left -= 1
    last = self.left <= 0
    return parts.TimeStep(parts.StepType.LAST if last else parts.StepType.MID, float(self.rs.randint(-1, 2)),
                          0.0 if last else 0.99, self._obs())


class _RepeatEvery2:
  """Preprocessor stand-in with action repeat: passes every 2nd frame (and every LAST), else None."""

  def __init__(self):
    self.k = 0

  def reset(self):
    self.k = 0

  def __call__(self, ts):
    self.k += 1
    return ts if (self.k % 2 == 1 or ts.last()) else None


@pytest.mark.parametrize('kind', ['rainbow', 'dqn', 'iqn'])
def test_agent_runs_in_run_loop_and_state_round_trips(kind):
  from dqn_zoo_b200 import parts
  from dqn_zoo_b200 import replay as dr
  structure = dr.Transition(None, None, None, None, None)

  def make(seed):
    rs = np.random.RandomState(seed)
    if kind == 'rainbow':
      rep = dr.PrioritizedTransitionReplay(96, structure, 0.5, parts.LinearSchedule(0.4, 1.0, begin_t=10, end_t=400), 1e-3,
                                           True, rs)
    else:
      rep = dr.TransitionReplay(96, structure, rs)
    return _make_agent(kind, rep, 11, True, preprocessor=_RepeatEvery2(), min_replay_capacity_fraction=0.4)

  a = make(1)
  frames = 0
  actions = []
  for env, ts, agent, act in parts.run_loop(a, _Env(3), max_steps_per_episode=15):
    if act is not None:
      assert isinstance(act, int) and 0 <= act < 6
      actions.append(act)
    frames += 1
    if frames >= 260:
      break
  assert a._replay.size == 96 and a._learn_steps > 10
  assert isinstance(a.statistics['state_value'], float) and np.isfinite(a.statistics['state_value'])
  ok, msg = a._replay.check_valid()
  assert ok, msg
  a.check_device_flags()
  st = copy.deepcopy(a.get_state())
  assert set(st) >= {'rng_key', 'frame_t', 'opt_state', 'online_params', 'target_params', 'replay'}
  b = make(2)
  b.set_state(st)
  b._action = a._action
  b._preprocessor.k = a._preprocessor.k
  b._transition_accumulator = copy.deepcopy(a._transition_accumulator)
  b._replay._random_state.set_state(a._replay._random_state.get_state())
  # both continue on identical inputs: identical actions, parameters and replay contents
  rs = np.random.RandomState(77)
  def same_params(tag):
    pa, pb = a.learner.get_params(), b.learner.get_params()
    for name in pa:
      np.testing.assert_array_equal(pa[name], pb[name], err_msg='%s %s' % (tag, name))
    np.testing.assert_array_equal(a.learner.get_params('target')['conv1/w'], b.learner.get_params('target')['conv1/w'])

  same_params('after set_state')
  assert a._frame_t == b._frame_t and a._replay._t == b._replay._t
  for i in range(40):
    ts = parts.TimeStep(parts.StepType.MID, float(rs.randint(-1, 2)), 0.99, rs.randint(0, 256, OBS).astype(np.uint8))
    act_a, act_b = a.step(ts), b.step(ts)
    same_params('step %d' % i)
    assert act_a == act_b, (i, act_a, act_b, a.statistics, b.statistics)
  torch.cuda.synchronize()
  pa, pb = a.learner.get_params(), b.learner.get_params()
  for name in pa:
    np.testing.assert_array_equal(pa[name], pb[name], err_msg=name)
  ids_a = [i for i, _ in a._replay.get_state()['storage']]
  ids_b = [i for i, _ in b._replay.get_state()['storage']]
  assert ids_a == ids_b


@pytest.mark.parametrize('kind', ['dqn', 'rainbow'])
def test_train_eval_iteration_with_trackers_actor_and_checkpoint(kind, tmp_path):
  """The structure of dqn/run_atari.py:237-290 on a dummy environment: train phase -> eval actor takes the online
  parameters -> statistics -> CSV row -> checkpoint; then a second process restores and continues identically."""
  import collections
  import itertools
  from dqn_zoo_b200 import agent as ag
  from dqn_zoo_b200 import learner as dl
  from dqn_zoo_b200 import parts
  from dqn_zoo_b200 import replay as dr
  from dqn_zoo_b200 import reporting
  structure = dr.Transition(None, None, None, None, None)

  def build(path):
    rs = np.random.RandomState(2)
    if kind == 'rainbow':
      rep = dr.PrioritizedTransitionReplay(64, structure, 0.5, lambda t: 0.5, 1e-3, True, rs)
    else:
      rep = dr.TransitionReplay(64, structure, rs)
    train = _make_agent(kind, rep, 13, False, preprocessor=_RepeatEvery2(), min_replay_capacity_fraction=0.25)
    evaluator = ag.EpsilonGreedyActor(preprocessor=_RepeatEvery2(), network=dl.NetworkSpec(kind, 6),
                                      exploration_epsilon=0.05, rng_key=[0, 99])
    ck = reporting.FileCheckpoint(path)
    ck.state.iteration = 0
    ck.state.train_agent = train
    ck.state.eval_agent = evaluator
    ck.state.random_state = rs
    ck.state.writer = reporting.CsvWriter(str(tmp_path / ('results_%s.csv' % kind)))
    return ck

  def iteration(ck, seed):
    st = ck.state
    train_seq = itertools.islice(parts.run_loop(st.train_agent, _Env(seed), max_steps_per_episode=30), 120)
    train_stats = reporting.generate_statistics(reporting.make_default_trackers(st.train_agent), train_seq)
    st.eval_agent.network_params = st.train_agent.online_params if seed % 2 else st.train_agent.learner
    eval_seq = itertools.islice(parts.run_loop(st.eval_agent, _Env(seed + 100), max_steps_per_episode=30), 60)
    eval_stats = reporting.generate_statistics(reporting.make_default_trackers(st.eval_agent), eval_seq)
    st.writer.write(collections.OrderedDict([('iteration', st.iteration), ('train_episode_return', train_stats['episode_return']),
                                             ('eval_episode_return', eval_stats['episode_return']),
                                             ('train_num_episodes', train_stats['num_episodes']),
                                             ('train_state_value', train_stats['state_value'])]))
    st.iteration += 1
    return train_stats, eval_stats

  path = str(tmp_path / ('ck_%s.pkl' % kind))
  ck = build(path)
  t0, e0 = iteration(ck, 1)
  assert t0['num_steps'] == 120 and e0['num_steps'] == 60 and np.isfinite(t0['state_value'])
  got = ck.state.eval_agent.network_params
  want = ck.state.train_agent.online_params
  for mod in want:
    for leaf in want[mod]:
      np.testing.assert_array_equal(got[mod][leaf], want[mod][leaf])
  ck.save()
  t1, e1 = iteration(ck, 2)                 # continue in the same "process"
  ck2 = build(path)                         # a fresh one restores and must reproduce that iteration
  assert ck2.can_be_restored()
  ck2.restore()
  assert ck2.state.iteration == 1
  t1b, e1b = iteration(ck2, 2)
  for key in ('episode_return', 'num_episodes', 'num_steps_over_episodes', 'state_value'):
    assert t1[key] == t1b[key] or (np.isnan(t1[key]) and np.isnan(t1b[key])), key
  assert e1['episode_return'] == e1b['episode_return'] or (np.isnan(e1['episode_return']) and np.isnan(e1b['episode_return']))
  a = ck.state.train_agent.get_state()
  b = ck2.state.train_agent.get_state()
  for name in a['online_params']:
    np.testing.assert_array_equal(a['online_params'][name], b['online_params'][name])
  rows = open(str(tmp_path / ('results_%s.csv' % kind))).read().strip().splitlines()
  assert rows[0].startswith('iteration,train_episode_return') and len(rows) == 4   # header + it0 + it1 + it1 again


@pytest.mark.parametrize('kind', ['dqn', 'double_q', 'c51', 'qrdqn', 'iqn', 'rainbow'])
def test_batched_acting_equals_per_stream_acting(kind):
  """dz_learner_act_batch (E environment streams in one enqueue, epsilon-greedy on the device) against E calls of the
  single-observation q_values path (dqn/agent.py:121-131): same q-values, first-argmax actions, and the documented
  exploration rule action = u0 < eps ? floor(u1 * A) : argmax."""
  from dqn_zoo_b200 import learner as dl
  rs = np.random.RandomState(5)
  E, A = 7, 6
  L = dl.Learner(dl.NetworkSpec(kind, A), batch_size=8)
  from oracle import learner_oracle as lo
  L.set_params(lo.init_params(lo.NetSpec(kind, A), 3), also_target=True)
  obs = torch.as_tensor(rs.randint(0, 256, size=(E, 84, 84, 4)).astype(np.uint8), device='cuda')
  taus = noise = None
  if kind == 'iqn':
    taus = torch.as_tensor(rs.uniform(size=(E, 64)).astype(np.float32), device='cuda')
  if kind == 'rainbow':
    L.generate_randomness(11)
    torch.cuda.synchronize()
    noise = L.noise.clone()
  explore = torch.as_tensor(rs.uniform(size=(2, E)).astype(np.float32), device='cuda')
  eps = 0.4
  actions, q = L.act_batch(obs, epsilon=eps, explore=explore, taus=taus, noise=noise)
  torch.cuda.synchronize()
  actions, q = actions.cpu().numpy(), q.cpu().numpy()
  u = explore.cpu().numpy()
  for e in range(E):
    q1 = L.q_values(obs[e], taus=None if taus is None else taus[e], noise=noise).cpu().numpy()
    np.testing.assert_allclose(q[e], q1, rtol=2e-6, atol=1e-6)
    want = min(int(u[1, e] * A), A - 1) if u[0, e] < eps else int(np.argmax(q[e]))
    assert actions[e] == want, (e, actions[e], want)
  greedy, _ = L.act_batch(obs, taus=taus, noise=noise)
  assert np.array_equal(greedy.cpu().numpy(), np.argmax(q, axis=1))


def test_batched_actor_runs_many_streams_with_one_copy_per_tick():
  from dqn_zoo_b200 import agent as agent_lib
  from dqn_zoo_b200 import learner as dl
  from oracle import learner_oracle as lo
  L = dl.Learner(dl.NetworkSpec('dqn', 6), batch_size=32)
  L.set_params(lo.init_params(lo.NetSpec('dqn', 6), 2), also_target=True)
  actor = agent_lib.BatchedEpsilonGreedyActor(L, 32, exploration_epsilon=0.05, rng_key=[0, 9])
  obs = torch.randint(0, 256, (32, 84, 84, 4), dtype=torch.uint8, device='cuda')
  a = actor.step(obs)
  assert a.shape == (32,) and a.dtype == np.int32 and a.min() >= 0 and a.max() < 6
  q = actor.q_values.cpu().numpy()
  agree = (a == np.argmax(q, axis=1)).mean()
  assert agree > 0.8   # epsilon = 0.05


def test_debug_timeline_stamps_every_kernel_of_a_graph_replayed_step():
  """dz_debug_timeline: every kernel appends (globaltimer, launch geometry) after its griddepcontrol.wait, also under CUDA
  graph replay — the instrument behind tools/step_timeline.py and bench.py's roofline timing.  One stamp per launch
  (kernels with several (0,0,z) blocks stamp once per z), monotone within a step, and nothing is written once removed."""
  from dqn_zoo_b200 import _lib
  from dqn_zoo_b200 import replay as dr
  rep = dr.TransitionReplay(512, dr.Transition(None, None, None, None, None), np.random.RandomState(4))
  dr.bulk_fill_synthetic(rep, OBS, 4, 6)
  ag = _make_agent('dqn', rep, 4, True)
  for _ in range(4):
    ag.learn()
  torch.cuda.synchronize()
  ag._use_graph = False
  c0 = _lib.lib.dz_launch_count()
  ag.learn()
  torch.cuda.synchronize()
  launches = int(_lib.lib.dz_launch_count() - c0)
  ag._use_graph = True
  tl = torch.zeros(2 + 2 * 4000, dtype=torch.int64, device='cuda')
  _lib.call('dz_debug_timeline', tl.data_ptr())
  steps = 5
  for _ in range(steps):
    ag.learn()
  torch.cuda.synchronize()
  _lib.call('dz_debug_timeline', 0)
  t = tl.cpu().numpy()
  n = int(t[0] & 0xffffffff)
  assert n >= launches * steps and n % steps == 0, (n, launches, steps)
  ts = np.sort(t[2:2 + 2 * n:2])
  assert np.all(np.diff(ts) >= 0) and ts[-1] - ts[0] < 50_000_000     # < 50 ms for five steps
  ag.learn()
  torch.cuda.synchronize()
  assert int(tl.cpu().numpy()[0] & 0xffffffff) == n
