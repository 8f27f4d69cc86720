"""CPU: trackers, CSV writer and checkpoint (reference behaviour: dqn_zoo/parts.py:125-333, 448-541; the expected
numbers below restate parts_test.py's scenarios)."""

import collections
import csv
import math

import numpy as np
import pytest

from dqn_zoo_b200 import parts
from dqn_zoo_b200 import reporting

F, M, L = parts.StepType.FIRST, parts.StepType.MID, parts.StepType.LAST


def ts(step_type, reward=None):
  return parts.TimeStep(step_type, reward, None if step_type == F else 1.0, None)


def test_episode_tracker_conventions():
  t = reporting.EpisodeTracker()
  with pytest.raises(RuntimeError):
    t.get()
  t.reset()
  g = t.get()
  assert math.isnan(g['episode_return']) and math.isnan(g['mean_episode_return']) and g['num_episodes'] == 0
  for x in [ts(F), ts(M, 1.0), ts(M, 2.0)]:
    t.step(None, x, None, None)
  g = t.get()   # no complete episode yet: episode_return is the running return
  assert g['episode_return'] == 3.0 and g['current_episode_return'] == 3.0 and math.isnan(g['mean_episode_return'])
  assert g['current_episode_step'] == 3 and g['num_steps_since_reset'] == 3 and g['num_steps_over_episodes'] == 0
  t.step(None, ts(L, 4.0), None, None)
  for x in [ts(F), ts(M, 10.0), ts(L, 1.0), ts(F), ts(M, 5.0)]:
    t.step(None, x, None, None)
  g = t.get()
  assert g['num_episodes'] == 2 and g['mean_episode_return'] == pytest.approx((7.0 + 11.0) / 2)
  assert g['episode_return'] == g['mean_episode_return'] and g['current_episode_return'] == 5.0
  assert g['num_steps_over_episodes'] == 7 and g['current_episode_step'] == 2 and g['num_steps_since_reset'] == 9
  with pytest.raises(ValueError):
    t.step(None, ts(F), None, None)   # FIRST in the middle of an episode


def test_step_rate_tracker_and_generate_statistics():
  class Agent:
    statistics = {'state_value': 2.0}
  seq = [(None, ts(F), Agent(), 0), (None, ts(M, 1.0), Agent(), 1), (None, ts(L, 1.0), Agent(), 0)]
  stats = reporting.generate_statistics(reporting.make_default_trackers(Agent()), seq)
  assert stats['num_steps'] == 3 and stats['step_rate'] > 0 and stats['episode_return'] == 2.0
  assert stats['state_value'] == pytest.approx(2.0)
  r = reporting.StepRateTracker()
  r.reset()
  assert math.isnan(r.get()['step_rate'])


def test_unbiased_average_matches_closed_form():
  class Agent:
    def __init__(self):
      self.statistics = {'x': 0.0}
  a = Agent()
  tr = reporting.UnbiasedExponentialWeightedAverageAgentTracker(step_size=0.1, initial_agent=a)
  tr.reset()
  values = [3.0, -1.0, 4.0, 1.0, 5.0]
  for v in values:
    a.statistics = {'x': v}
    tr.step(None, None, a, None)
  w = np.array([0.9 ** (len(values) - 1 - i) for i in range(len(values))])
  assert tr.get()['x'] == pytest.approx(float((w * values).sum() / w.sum()))
  tr.reset()
  assert tr.get() == {'x': 0.0} and tr.trace == 0.0


def test_csv_writer_is_resumable(tmp_path):
  path = str(tmp_path / 'sub' / 'results.csv')
  w = reporting.CsvWriter(path)
  w.write(collections.OrderedDict([('iteration', 0), ('frame', 100), ('eval_episode_return', 1.5)]))
  state = w.get_state()
  w2 = reporting.CsvWriter(path)
  w2.set_state(state)
  w2.write(collections.OrderedDict([('iteration', 1), ('frame', 200), ('eval_episode_return', 2.5)]))
  rows = list(csv.DictReader(open(path)))
  assert [r['frame'] for r in rows] == ['100', '200'] and list(rows[0].keys()) == ['iteration', 'frame', 'eval_episode_return']
  reporting.NullWriter().write({'a': 1})


def test_file_checkpoint_round_trip(tmp_path):
  class Thing:
    def __init__(self, v):
      self.v = v
    def get_state(self):
      return {'v': self.v}
    def set_state(self, s):
      self.v = s['v']
  null = reporting.NullCheckpoint()
  null.state.iteration = 3
  null.save()
  assert not null.can_be_restored() and null.state.iteration == 3
  ck = reporting.FileCheckpoint(str(tmp_path / 'ck' / 'state.pkl'))
  assert not ck.can_be_restored()
  ck.state.iteration = 7
  ck.state.agent = Thing(np.arange(4))
  ck.state.random_state = np.random.RandomState(5)
  ck.state.random_state.uniform(size=3)
  want_next = np.random.RandomState(5)
  want_next.uniform(size=3)
  ck.save()
  # a fresh process: objects are re-created by the driver, then restored in place
  ck2 = reporting.FileCheckpoint(str(tmp_path / 'ck' / 'state.pkl'))
  ck2.state.iteration = 0
  ck2.state.agent = Thing(None)
  ck2.state.random_state = np.random.RandomState(0)
  assert ck2.can_be_restored()
  ck2.restore()
  assert ck2.state.iteration == 7 and np.array_equal(ck2.state.agent.v, np.arange(4))
  assert ck2.state.random_state.uniform() == want_next.uniform()
  with pytest.raises(AttributeError):
    ck2.state.missing
