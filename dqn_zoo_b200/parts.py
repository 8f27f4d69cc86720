"""The agent surface and run loop of `dqn_zoo/parts.py`, kept verbatim in behaviour.

Only what the hot path's drop-in boundary needs: `Agent` (parts.py:42-67), `run_loop`
(parts.py:70-122), `LinearSchedule` (parts.py:414-430) and minimal `dm_env` stand-ins
(dm_env is not installed in this image; any object with `.first()/.mid()/.last()`,
`.observation/.reward/.discount` and `._replace(step_type=...)` works).
"""

import abc
import enum
from typing import Any, Iterable, Mapping, NamedTuple, Optional, Tuple

Action = int


class StepType(enum.IntEnum):
  FIRST = 0
  MID = 1
  LAST = 2


class TimeStep(NamedTuple):
  """dm_env.TimeStep stand-in."""
  step_type: Any
  reward: Any
  discount: Any
  observation: Any

  def first(self):
    return self.step_type == StepType.FIRST

  def mid(self):
    return self.step_type == StepType.MID

  def last(self):
    return self.step_type == StepType.LAST


class Agent(abc.ABC):
  """Agent interface (parts.py:42-67)."""

  @abc.abstractmethod
  def step(self, timestep) -> Action:
    """Selects action given timestep and potentially learns."""

  @abc.abstractmethod
  def reset(self) -> None:
    """Resets the agent's episodic state such as frame stack and action repeat."""

  @abc.abstractmethod
  def get_state(self) -> Mapping[str, Any]:
    """Retrieves agent state as a dictionary (e.g. for serialization)."""

  @abc.abstractmethod
  def set_state(self, state: Mapping[str, Any]) -> None:
    """Sets agent state from a (potentially de-serialized) dictionary."""

  @property
  @abc.abstractmethod
  def statistics(self) -> Mapping[str, float]:
    """Returns current agent statistics as a dictionary."""


def run_loop(agent, environment, max_steps_per_episode: int = 0, yield_before_reset: bool = False):
  """Alternates environment and agent steps (parts.py:70-122).

  Yields `(environment, timestep_t, agent, a_t)`; truncates an episode at
  `max_steps_per_episode` by relabelling the timestep LAST (parts.py:115-117) and gives the
  agent one extra step on LAST whose action is ignored (parts.py:119-122).
  """
  while True:
    if yield_before_reset:
      yield environment, None, agent, None
    steps = 0
    agent.reset()
    ts = environment.reset()
    while True:
      action = agent.step(ts)
      yield environment, ts, agent, action
      steps += 1
      ts = environment.step(action)
      if max_steps_per_episode > 0 and steps >= max_steps_per_episode:
        assert steps == max_steps_per_episode
        ts = ts._replace(step_type=StepType.LAST)
      if ts.last():
        agent.step(ts)
        yield environment, ts, agent, None
        break


class LinearSchedule:
  """Linear schedule (parts.py:414-430): begin_value until begin_t, then linear to end_value."""

  def __init__(self, begin_value, end_value, begin_t, end_t=None, decay_steps=None):
    if (end_t is None) == (decay_steps is None):
      raise ValueError('Exactly one of end_t, decay_steps must be provided.')
    self._decay_steps = decay_steps if end_t is None else end_t - begin_t
    self._begin_t = begin_t
    self._begin_value = begin_value
    self._end_value = end_value

  def __call__(self, t):
    frac = min(max(t - self._begin_t, 0), self._decay_steps) / self._decay_steps
    return (1 - frac) * self._begin_value + frac * self._end_value
