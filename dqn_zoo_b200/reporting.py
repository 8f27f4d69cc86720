"""Run-loop reporting and persistence around the agent surface (reference `dqn_zoo/parts.py`).

Behavioural contract kept from the reference so that its plotting notebook and run drivers work unchanged:
  * `generate_statistics(trackers, sequence)` — parts.py:125-146: reset every tracker, feed every
    `(environment, timestep, agent, action)` item, merge the trackers' dicts (first tracker wins on key clashes,
    as `collections.ChainMap` does).
  * `EpisodeTracker` — parts.py:149-247: the seven keys `mean_episode_return, current_episode_return, episode_return,
    num_episodes, num_steps_over_episodes, current_episode_step, num_steps_since_reset` with the same conventions
    (`episode_return` falls back to the running return until one episode has completed; NaN before any step).
  * `StepRateTracker` — parts.py:250-287: `step_rate, num_steps, duration`.
  * `UnbiasedExponentialWeightedAverageAgentTracker` — parts.py:290-333 (Sutton & Barto's unbiased constant-step-size
    trick on `agent.statistics`).
  * `CsvWriter` / `NullWriter` — parts.py:448-504: one header row from the first dict's keys, append mode, resumable.
  * `NullCheckpoint` / `AttributeDict` — parts.py:507-541.  `FileCheckpoint` is the working implementation behind the
    same three methods (`save`, `can_be_restored`, `restore`) that the reference leaves as a placeholder.

Everything here is host-side bookkeeping; nothing touches the GPU."""

import collections
import csv
import math
import os
import pickle
import tempfile
import timeit
from typing import Any, Iterable, Mapping, Optional, Sequence


def generate_statistics(trackers: Sequence[Any], timestep_action_sequence: Iterable[Any]) -> Mapping[str, Any]:
  for t in trackers:
    t.reset()
  for environment, timestep, agent, action in timestep_action_sequence:
    for t in trackers:
      t.step(environment, timestep, agent, action)
  merged = {}
  for t in reversed(list(trackers)):     # earlier trackers take precedence, like ChainMap(*dicts)
    merged.update(t.get())
  return merged


class EpisodeTracker:
  """Episode returns and step counts."""

  def __init__(self):
    self._ready = False

  def reset(self) -> None:
    self._ready = True
    self._steps = 0                # num_steps_since_reset
    self._steps_in_done = 0        # num_steps_over_episodes
    self._returns = []             # completed episodes
    self._rewards = []             # rewards of the episode in progress
    self._episode_step = 0

  def step(self, environment, timestep, agent, action) -> None:
    del environment, agent, action
    if not self._ready:
      raise RuntimeError('reset() must be called before first call to step().')
    if timestep.first():
      if self._rewards:
        raise ValueError('Current episode reward list should be empty.')
      if self._episode_step != 0:
        raise ValueError('Current episode step should be zero.')
    else:
      self._rewards.append(timestep.reward)
    self._steps += 1
    self._episode_step += 1
    if timestep.last():
      self._returns.append(sum(self._rewards))
      self._rewards = []
      self._steps_in_done += self._episode_step
      self._episode_step = 0

  def get(self) -> Mapping[str, Any]:
    if not self._ready:
      raise RuntimeError('reset() must be called before first call to get().')
    running = sum(self._rewards)
    if self._returns:
      mean = sum(self._returns) / len(self._returns)
      current, shown = running, mean
    else:
      mean = math.nan
      current = running if self._steps > 0 else math.nan
      shown = current
    return {
        'mean_episode_return': mean,
        'current_episode_return': current,
        'episode_return': shown,
        'num_episodes': len(self._returns),
        'num_steps_over_episodes': self._steps_in_done,
        'current_episode_step': self._episode_step,
        'num_steps_since_reset': self._steps,
    }


class StepRateTracker:
  """Steps per second since the last reset."""

  def __init__(self):
    self._steps = None
    self._t0 = None

  def reset(self) -> None:
    self._steps = 0
    self._t0 = timeit.default_timer()

  def step(self, environment, timestep, agent, action) -> None:
    del environment, timestep, agent, action
    self._steps += 1

  def get(self) -> Mapping[str, float]:
    if self._steps is None:
      raise RuntimeError('reset() must be called before first call to get().')
    duration = timeit.default_timer() - self._t0
    return {'step_rate': self._steps / duration if self._steps > 0 else math.nan, 'num_steps': self._steps,
            'duration': duration}


class UnbiasedExponentialWeightedAverageAgentTracker:
  """Bias-corrected exponential average of `agent.statistics` (a flat mapping of floats)."""

  def __init__(self, step_size: float, initial_agent):
    self._initial = dict(initial_agent.statistics)
    self._alpha = step_size
    self.reset()

  def reset(self) -> None:
    self.trace = 0.0
    self._avg = dict(self._initial)

  def step(self, environment, timestep, agent, action) -> None:
    del environment, timestep, action
    self.trace = (1 - self._alpha) * self.trace + self._alpha
    beta = self._alpha / self.trace
    assert 0 <= beta <= 1
    stats = agent.statistics
    if beta == 1:
      self._avg = dict(stats)
    else:
      self._avg = {k: (1 - beta) * self._avg[k] + beta * stats[k] for k in self._avg}

  def get(self) -> Mapping[str, float]:
    return self._avg


def make_default_trackers(initial_agent) -> Sequence[Any]:
  return [EpisodeTracker(), StepRateTracker(),
          UnbiasedExponentialWeightedAverageAgentTracker(step_size=1e-3, initial_agent=initial_agent)]


class NullWriter:
  def write(self, *args, **kwargs) -> None:
    pass

  def close(self) -> None:
    pass


class CsvWriter:
  """Appends one row per `write(OrderedDict)`; the first call fixes the columns."""

  def __init__(self, fname: str):
    folder = os.path.dirname(fname)
    if folder and not os.path.exists(folder):
      os.makedirs(folder)
    self._fname = fname
    self._header_written = False
    self._fieldnames = None

  def write(self, values: Mapping[str, Any]) -> None:
    if self._fieldnames is None:
      self._fieldnames = list(values.keys())
    with open(self._fname, 'a', newline='') as f:
      w = csv.DictWriter(f, fieldnames=self._fieldnames)
      if not self._header_written:
        w.writeheader()
        self._header_written = True
      w.writerow(values)

  def close(self) -> None:
    pass

  def get_state(self) -> Mapping[str, Any]:
    return {'header_written': self._header_written, 'fieldnames': self._fieldnames}

  def set_state(self, state: Mapping[str, Any]) -> None:
    self._header_written = state['header_written']
    self._fieldnames = state['fieldnames']


class AttributeDict(dict):
  """dict with attribute access (`state.iteration = 3`)."""

  def __getattr__(self, key):
    try:
      return self[key]
    except KeyError as e:
      raise AttributeError(key) from e

  def __setattr__(self, key, value):
    self[key] = value

  def __delattr__(self, key):
    del self[key]


class NullCheckpoint:
  """Checkpointing disabled: same surface, no effect."""

  def __init__(self):
    self.state = AttributeDict()

  def save(self) -> None:
    pass

  def can_be_restored(self) -> bool:
    return False

  def restore(self) -> None:
    pass


class FileCheckpoint:
  """Working checkpoint behind NullCheckpoint's interface.

  The run driver registers live objects in `checkpoint.state` (`state.train_agent = agent`, `state.iteration = 0`,
  `state.random_state = np.random.RandomState(...)`, `state.writer = CsvWriter(...)`, as dqn/run_atari.py:237-256 does).
  `save()` pickles a snapshot — `get_state()` of every entry that has one (agents, replay, writers), numpy
  `RandomState.get_state()` for random states, the value itself otherwise — atomically (write + rename).
  `restore()` pushes the snapshot back INTO the registered objects (`set_state`) and overwrites plain values, so the
  objects the driver already holds continue from the checkpoint."""

  def __init__(self, path: str):
    self._path = path
    self.state = AttributeDict()

  @staticmethod
  def _snapshot(value):
    if hasattr(value, 'get_state') and hasattr(value, 'set_state'):
      return ('stateful', value.get_state())
    return ('value', value)

  def save(self) -> None:
    payload = {k: self._snapshot(v) for k, v in self.state.items()}
    folder = os.path.dirname(os.path.abspath(self._path))
    os.makedirs(folder, exist_ok=True)
    fd, tmp = tempfile.mkstemp(dir=folder, suffix='.tmp')
    try:
      with os.fdopen(fd, 'wb') as f:
        pickle.dump(payload, f, protocol=pickle.HIGHEST_PROTOCOL)
      os.replace(tmp, self._path)
    except BaseException:
      if os.path.exists(tmp):
        os.remove(tmp)
      raise

  def can_be_restored(self) -> bool:
    return os.path.exists(self._path)

  def restore(self) -> None:
    with open(self._path, 'rb') as f:
      payload = pickle.load(f)
    for key, (kind, value) in payload.items():
      if kind == 'stateful':
        if key not in self.state:
          raise KeyError('checkpoint entry %r has no registered object to restore into' % key)
        self.state[key].set_state(value)
      else:
        self.state[key] = value
