"""CPU oracle for the replay half of the hot path (TEST INFRASTRUCTURE ONLY).

This is a numpy restatement of the algorithms in the reference's
`dqn_zoo/replay.py`.  It exists so the CUDA path can be checked against a
CPU implementation that travels to the GPU box (the reference itself does not).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import it; the product package `dqn_zoo_b200` never does.

Parity status: PINNED.  `oracle/gen_golden.py` imports the reference's own
`replay.py` (in the build container, where /root/reference exists) and writes
`tests/golden/replay_*.npz`; `tests/test_oracle_replay.py` checks this file
against those vectors and against the reference's known-answer tables
(`replay_test.py:939-953`, `:468`, `:762-772`, `:209-244`).

Every function cites the reference lines it restates.  The data structures are
deliberately different from the reference's (flat numpy arrays instead of
dict/list/OrderedDict) because they double as the specification of the HBM
layout used by the CUDA implementation.
"""

from __future__ import annotations

import collections
import math
from typing import Any, Callable, Iterable, List, Mapping, NamedTuple, Optional, Sequence, Tuple

import numpy as np


class Transition(NamedTuple):
  """Flat replay item, field order as `replay.py:36-41`."""
  s_tm1: Any
  a_tm1: Any
  r_t: Any
  discount_t: Any
  s_t: Any


# ----------------------------------------------------------------------------
# R3 / R4: scalar helpers
# ----------------------------------------------------------------------------


def power_keep_zero(base, exponent):
  """`replay.py:203-208` (`_power`): base**exponent except 0**0 -> 0.

  dtype follows numpy: a float32 array stays float32 (this is what
  `update_priorities` sees, the priorities having come back from the device as
  float32), python/np.float64 input is evaluated in float64 (the `add` path).
  For exponent 0.5 numpy evaluates `**` as a correctly rounded sqrt in the
  array's dtype, which is what the CUDA path does with `sqrt.rn`.
  """
  b = np.asarray(base)
  if exponent == 0.5 or b.dtype != np.float32:
    out = b ** exponent  # numpy's `**` fast-paths 0.5 to a correctly rounded sqrt
  else:
    # Canonical float32 definition for non-sqrt exponents:
    # round_f32(pow_f64(x, (double)(float)alpha)).  numpy's own SIMD powf is
    # library dependent (+-1ulp), so the oracle pins this one.
    out = np.power(b.astype(np.float64), np.float64(np.float32(exponent))).astype(np.float32)
  return np.where(b == 0.0, np.zeros_like(out), out)


def importance_sampling_weights(probabilities, uniform_probability, exponent, normalize):
  """`replay.py:211-243`: w = (u/p)**beta, optionally divided by the batch max."""
  if not 0.0 <= exponent <= 1.0:
    raise ValueError('Require 0 <= exponent <= 1.')
  if not 0.0 <= uniform_probability <= 1.0:
    raise ValueError('Expected 0 <= uniform_probability <= 1.')
  w = (uniform_probability / np.asarray(probabilities, dtype=np.float64)) ** exponent
  if normalize:
    w = w / w.max()
  if not np.isfinite(w).all():
    raise ValueError('Weights are not finite: %s.' % w)
  return w


# ----------------------------------------------------------------------------
# R1: sum tree
# ----------------------------------------------------------------------------


class SumTree:
  """Array-embedded binary sum tree, float64 (`replay.py:246-426`).

  node i has children 2i and 2i+1, root is node 1, leaves start at
  `first_leaf` (a power of two >= size).  Every internal node is always
  *recomputed* as fl(left + right), never delta-updated, so the tree is a pure
  function of its leaves (`replay.py:284-290`, `:394-404`).
  """

  def __init__(self):
    self._n = 0
    self._first_leaf = 0
    self._nodes = np.zeros(0, dtype=np.float64)

  # -- sizes ---------------------------------------------------------------
  @property
  def size(self):
    return self._n

  @property
  def capacity(self):
    return self._first_leaf

  @property
  def values(self):
    return self._nodes[self._first_leaf:self._first_leaf + self._n]

  def root(self):
    """`replay.py:315-317`: NaN when empty."""
    return self._nodes[1] if self._n > 0 else np.nan

  # -- bulk (re)build --------------------------------------------------------
  def _rebuild(self, leaves):
    """`replay.py:394-404` (`_set_values`): write leaves, zero the rest, resum."""
    fl = self._first_leaf
    k = len(leaves)
    assert k <= fl
    self._nodes[fl:fl + k] = leaves
    self._nodes[fl + k:] = 0.0
    lo = fl
    while lo > 1:  # One vectorised pass per level, bottom-up.
      half = lo // 2
      self._nodes[half:lo] = self._nodes[lo:2 * lo:2] + self._nodes[lo + 1:2 * lo:2]
      lo = half
    if len(self._nodes):
      self._nodes[0] = 0.0

  def _reshape(self, size, leaves):
    """`replay.py:361-392` (`_initialize`)."""
    assert size >= 0
    if size < self._n:
      keep = self.values[:size].copy() if leaves is None else leaves
      self._n = size
      self._rebuild(keep)
    elif size <= self._first_leaf:
      self._n = size
      if leaves is not None:
        self._rebuild(leaves)
    else:
      cap = 1
      while cap < size:
        cap *= 2
      keep = self.values.copy() if leaves is None else leaves
      self._nodes = np.empty(2 * cap, dtype=np.float64)
      self._first_leaf = cap
      self._n = size
      self._rebuild(keep)

  def resize(self, size):
    """`replay.py:267-269`."""
    self._reshape(size, None)

  def set_all(self, values):
    """`replay.py:292-297`."""
    v = np.asarray(values, dtype=np.float64)
    if not np.isfinite(v).all() or (v < 0.0).any():
      raise ValueError('Values must be finite positive numbers.')
    self._reshape(len(v), v)

  # -- point ops -------------------------------------------------------------
  def get(self, indices):
    """`replay.py:271-276`."""
    idx = np.asarray(indices)
    if idx.size and not ((0 <= idx) & (idx < self._n)).all():
      raise IndexError('index out of range, expect 0 <= index < %s' % self._n)
    return self.values[idx]

  def set(self, indices, values):
    """`replay.py:278-290`: duplicates -> last write wins, then resum paths."""
    v = np.asarray(values)
    if not np.isfinite(v).all() or (v < 0.0).any():
      raise ValueError('value must be finite and positive.')
    idx = np.asarray(indices, dtype=np.int64)
    self.values[idx] = v
    nodes = self._nodes
    for leaf in idx + self._first_leaf:
      p = int(leaf) >> 1
      while p > 0:
        nodes[p] = nodes[2 * p] + nodes[2 * p + 1]
        p >>= 1

  def query_one(self, target):
    """`replay.py:406-426`: smallest index whose inclusive prefix sum > target."""
    if not 0.0 <= target < self.root():
      raise ValueError('Require 0 <= target < total sum.')
    nodes = self._nodes
    i = 1
    while i < self._first_leaf:
      left = nodes[2 * i]
      if target < left:
        i = 2 * i
      else:
        target = target - left
        i = 2 * i + 1
    return i - self._first_leaf

  def query(self, targets):
    """`replay.py:299-313`."""
    return [self.query_one(t) for t in targets]

  # -- state -----------------------------------------------------------------
  def get_state(self):
    """`replay.py:334-340`."""
    return {'size': self._n, 'storage': self._nodes, 'first_leaf': self._first_leaf}

  def set_state(self, state):
    """`replay.py:342-346`."""
    self._n = state['size']
    self._nodes = state['storage']
    self._first_leaf = state['first_leaf']

  def check_valid(self):
    """`replay.py:348-359`."""
    if len(self._nodes) != 2 * self._first_leaf:
      return False, 'first_leaf should be half the size of storage.'
    if not 0 <= self._n <= self._first_leaf:
      return False, 'Require 0 <= self.size <= self.capacity.'
    fl = self._first_leaf
    if fl > 1:
      inner = self._nodes[1:fl]
      sums = self._nodes[2:2 * fl:2] + self._nodes[3:2 * fl:2]
      bad = np.nonzero(inner != sums)[0]
      if bad.size:
        return False, 'Non-leaf node %d should be sum of child nodes.' % (bad[0] + 1)
    return True, ''


# ----------------------------------------------------------------------------
# R6: uniform distribution + replay
# ----------------------------------------------------------------------------


class UniformDistribution:
  """Uniform sampling over user ids with swap-remove bookkeeping (`replay.py:44-117`)."""

  def __init__(self, random_state):
    self._rs = random_state
    self._slots: List[int] = []     # dense list of ids (the reference's `_ids`)
    self._where = {}                # id -> position in _slots

  def add(self, ids):
    """`replay.py:52-61`."""
    for i in ids:
      if i in self._where:
        raise IndexError('Cannot add ID %d, it already exists.' % i)
    for i in ids:
      self._where[i] = len(self._slots)
      self._slots.append(i)

  def remove(self, ids):
    """`replay.py:63-74`: the LAST id moves into the hole."""
    for i in ids:
      if i not in self._where:
        raise IndexError('Cannot remove ID %d, it does not exist.' % i)
    for i in ids:
      hole = self._where.pop(i)
      tail = self._slots.pop()
      if tail != i:
        self._slots[hole] = tail
        self._where[tail] = hole

  def sample(self, size):
    """`replay.py:76-82`: one `randint(size, size=B)` draw."""
    picks = self._rs.randint(self.size, size=size)
    return np.asarray([self._slots[p] for p in picks], dtype=np.int64)

  def ids(self):
    return self._where.keys()

  @property
  def size(self):
    return len(self._slots)

  def get_state(self):
    """`replay.py:93-98`."""
    return {'ids': self._slots, 'id_to_index': self._where}

  def set_state(self, state):
    """`replay.py:100-103`."""
    self._slots = state['ids']
    self._where = state['id_to_index']

  def check_valid(self):
    """`replay.py:105-117`."""
    if len(self._slots) != len(self._where):
      return False, 'ids and id_to_index should be the same size.'
    if len(set(self._slots)) != len(self._slots):
      return False, 'IDs should be unique.'
    for pos, i in enumerate(self._slots):
      if self._where.get(i) != pos:
        return False, 'ID %d should map to itself.' % i
    return True, ''


def _stack_fields(structure, items):
  """`replay.py:162-165` / `:718-722`: transpose then np.stack per field."""
  cols = list(zip(*items))
  return type(structure)(*[np.stack(c, axis=0) for c in cols])


class TransitionReplay:
  """Uniform replay with oldest-out eviction (`replay.py:120-200`)."""

  def __init__(self, capacity, structure, random_state, encoder=None, decoder=None):
    self._capacity = capacity
    self._structure = structure
    self._enc = encoder or (lambda s: s)
    self._dec = decoder or (lambda s: s)
    self._distribution = UniformDistribution(random_state)
    self._items = collections.OrderedDict()
    self._t = 0

  def add(self, item):
    """`replay.py:142-151`."""
    if len(self._items) == self._capacity:
      old, _ = self._items.popitem(last=False)
      self._distribution.remove([old])
    self._distribution.add([self._t])
    self._items[self._t] = self._enc(item)
    self._t += 1

  def get(self, ids):
    """`replay.py:153-156`."""
    return [self._dec(self._items[i]) for i in ids]

  def sample_ids(self, size):
    return self._distribution.sample(size)

  def sample(self, size):
    """`replay.py:158-165`."""
    return _stack_fields(self._structure, self.get(self.sample_ids(size)))

  def ids(self):
    return self._items.keys()

  @property
  def size(self):
    return len(self._items)

  @property
  def capacity(self):
    return self._capacity

  def get_state(self):
    """`replay.py:181-187`."""
    return {'storage': list(self._items.items()), 't': self._t,
            'distribution': self._distribution.get_state()}

  def set_state(self, state):
    """`replay.py:189-193`."""
    self._items = collections.OrderedDict(state['storage'])
    self._t = state['t']
    self._distribution.set_state(state['distribution'])

  def check_valid(self):
    """`replay.py:195-200`."""
    if self._t < len(self._items):
      return False, 't should be >= storage size.'
    if set(self._items.keys()) != set(self._distribution.ids()):
      return False, 'IDs in storage and distribution do not match.'
    return self._distribution.check_valid()


# ----------------------------------------------------------------------------
# R2: prioritized distribution
# ----------------------------------------------------------------------------


class PrioritizedDistribution:
  """Proportional prioritized sampling of ids (`replay.py:429-651`).

  Slot bookkeeping restated:
    * `_free` is a stack of unused tree indices; allocation pops from the END
      (`replay.py:457,499`), a removed index is pushed back (`:533`).
    * `_live` is the dense list of indices in use and `_live_pos[idx]` its
      position; removal swaps the last live index into the hole (`:519-531`).
  """

  def __init__(self, priority_exponent, uniform_sample_probability, random_state,
               min_capacity=0, max_capacity=None):
    if priority_exponent < 0.0:
      raise ValueError('Require priority_exponent >= 0.')
    if not 0.0 <= uniform_sample_probability <= 1.0:
      raise ValueError('Require 0 <= uniform_sample_probability <= 1.')
    if max_capacity is not None and max_capacity < min_capacity:
      raise ValueError('Require max_capacity >= min_capacity.')
    if min_capacity < 0:
      raise ValueError('Require min_capacity >= 0.')
    self._alpha = priority_exponent
    self._usp = uniform_sample_probability
    self._max_capacity = max_capacity
    self._rs = random_state
    self._tree = SumTree()
    self._tree.resize(min_capacity)
    self._idx_of = {}     # id -> tree index
    self._id_at = {}      # tree index -> id
    self._free = list(range(min_capacity))
    self._live: List[int] = []
    self._live_pos = {}

  def ensure_capacity(self, capacity):
    """`replay.py:463-473`."""
    if self._max_capacity is not None and capacity > self._max_capacity:
      raise ValueError('capacity %d cannot exceed max_capacity %d' % (capacity, self._max_capacity))
    if capacity <= self._tree.size:
      return
    self._free.extend(range(self._tree.size, capacity))
    self._tree.resize(capacity)

  def add_priorities(self, ids, priorities):
    """`replay.py:475-507`."""
    for i in ids:
      if i in self._idx_of:
        raise IndexError('ID %d already exists.' % i)
    want = self.size + len(ids)
    if self._max_capacity is not None and want > self._max_capacity:
      raise ValueError('Cannot add IDs as max capacity would be exceeded.')
    if want > self.capacity:
      grown = max(want, 2 * self.capacity)
      if self._max_capacity is not None:
        grown = min(self._max_capacity, grown)
      self.ensure_capacity(grown)
    got = []
    for i in ids:
      idx = self._free.pop()
      self._live_pos[idx] = len(self._live)
      self._live.append(idx)
      self._idx_of[i] = idx
      self._id_at[idx] = i
      got.append(idx)
    self._tree.set(got, power_keep_zero(priorities, self._alpha))

  def remove_priorities(self, ids):
    """`replay.py:509-534`."""
    gone = [self._idx_of[i] for i in ids]  # KeyError if absent, as the reference
    for i, idx in zip(ids, gone):
      del self._idx_of[i]
      del self._id_at[idx]
      hole = self._live_pos.pop(idx)
      tail = self._live.pop()
      if tail != idx:
        self._live[hole] = tail
        self._live_pos[tail] = hole
    self._free.extend(gone)
    self._tree.set(gone, np.zeros(len(gone), dtype=np.float64))

  def update_priorities(self, ids, priorities):
    """`replay.py:536-545`."""
    where = []
    for i in ids:
      if i not in self._idx_of:
        raise IndexError('ID %d does not exist.' % i)
      where.append(self._idx_of[i])
    self._tree.set(where, power_keep_zero(priorities, self._alpha))

  def sample_indices(self, size):
    """`replay.py:547-577` up to (indices, probabilities); RNG draw order is
    randint -> [uniform iff root != 0] -> uniform."""
    if self.size == 0:
      raise RuntimeError('No IDs to sample.')
    uni = np.asarray([self._live[j] for j in self._rs.randint(self.size, size=size)], dtype=np.int64)
    root = self._tree.root()
    if root == 0.0:
      pri = uni
    else:
      targets = self._rs.uniform(size=size) * root
      pri = np.asarray(self._tree.query(targets), dtype=np.int64)
    pick_uniform = self._rs.uniform(size=size) < self._usp
    idx = np.where(pick_uniform, uni, pri)
    one_over_n = np.asarray(1.0 / self.size)
    leaf = self._tree.get(idx)
    frac = np.full_like(leaf, fill_value=one_over_n) if root == 0.0 else leaf / root
    probs = (1.0 - self._usp) * frac + self._usp * one_over_n
    return idx, probs

  def sample(self, size):
    """`replay.py:547-583`."""
    idx, probs = self.sample_indices(size)
    ids = np.asarray([self._id_at[int(k)] for k in idx], dtype=np.int64)
    return ids, probs

  def get_exponentiated_priorities(self, ids):
    """`replay.py:585-590`."""
    return self._tree.get(np.asarray([self._idx_of[i] for i in ids], dtype=np.int64))

  def ids(self):
    return self._idx_of.keys()

  @property
  def capacity(self):
    return self._tree.size

  @property
  def size(self):
    return len(self._idx_of)

  def get_state(self):
    """`replay.py:606-615` (same keys)."""
    return {
        'sum_tree': self._tree.get_state(),
        'id_to_index': self._idx_of,
        'index_to_id': self._id_at,
        'inactive_indices': self._free,
        'active_indices': self._live,
        'active_indices_location': self._live_pos,
    }

  def set_state(self, state):
    """`replay.py:617-624`."""
    self._tree.set_state(state['sum_tree'])
    self._idx_of = state['id_to_index']
    self._id_at = state['index_to_id']
    self._free = state['inactive_indices']
    self._live = state['active_indices']
    self._live_pos = state['active_indices_location']

  def check_valid(self):
    """`replay.py:626-651`."""
    if len(self._idx_of) != len(self._id_at):
      return False, 'ID to index maps are not the same size.'
    for i, idx in self._idx_of.items():
      if self._id_at.get(idx) != i:
        return False, 'ID %d should map to itself.' % i
    if len(set(self._free)) != len(self._free):
      return False, 'Inactive indices should be unique.'
    if len(set(self._live)) != len(self._live):
      return False, 'Active indices should be unique.'
    if set(self._live) != set(self._id_at.keys()):
      return False, 'Active indices should match index to ID mapping keys.'
    if sorted(self._free + self._live) != list(range(self._tree.size)):
      return False, 'Inactive and active indices should partition all indices.'
    for pos, idx in enumerate(self._live):
      if self._live_pos.get(idx) != pos:
        return False, 'Active index location %d not correct for index %d.' % (pos, idx)
    return self._tree.check_valid()


# ----------------------------------------------------------------------------
# R5: prioritized replay
# ----------------------------------------------------------------------------


class PrioritizedTransitionReplay:
  """Proportional PER with oldest-out eviction (`replay.py:654-768`)."""

  def __init__(self, capacity, structure, priority_exponent, importance_sampling_exponent,
               uniform_sample_probability, normalize_weights, random_state,
               encoder=None, decoder=None):
    self._capacity = capacity
    self._structure = structure
    self._enc = encoder or (lambda s: s)
    self._dec = decoder or (lambda s: s)
    self._distribution = PrioritizedDistribution(
        priority_exponent, uniform_sample_probability, random_state,
        min_capacity=capacity, max_capacity=capacity)
    self._beta = importance_sampling_exponent
    self._normalize = normalize_weights
    self._items = collections.OrderedDict()
    self._t = 0

  def add(self, item, priority):
    """`replay.py:690-699`."""
    if len(self._items) == self._capacity:
      old, _ = self._items.popitem(last=False)
      self._distribution.remove_priorities([old])
    self._distribution.add_priorities([self._t], [priority])
    self._items[self._t] = self._enc(item)
    self._t += 1

  def get(self, ids):
    return [self._dec(self._items[i]) for i in ids]

  def sample_ids(self, size):
    """`replay.py:710-717`: (ids, probabilities, weights) without the gather."""
    ids, probs = self._distribution.sample(size)
    w = importance_sampling_weights(probs, 1.0 / self.size, self.importance_sampling_exponent,
                                    self._normalize)
    return ids, probs, w

  def sample(self, size):
    """`replay.py:706-723`."""
    ids, _, w = self.sample_ids(size)
    return _stack_fields(self._structure, self.get(ids)), ids, w

  def update_priorities(self, ids, priorities):
    """`replay.py:725-730`."""
    self._distribution.update_priorities(ids, np.asarray(priorities))

  @property
  def size(self):
    return len(self._items)

  @property
  def capacity(self):
    return self._capacity

  @property
  def importance_sampling_exponent(self):
    """`replay.py:742-745`: schedule evaluated at the number of adds."""
    return self._beta(self._t)

  def get_state(self):
    """`replay.py:747-754`."""
    return {'storage': list(self._items.items()), 't': self._t,
            'distribution': self._distribution.get_state()}

  def set_state(self, state):
    """`replay.py:756-760`."""
    self._items = collections.OrderedDict(state['storage'])
    self._t = state['t']
    self._distribution.set_state(state['distribution'])

  def check_valid(self):
    """`replay.py:762-768`."""
    if self._t < len(self._items):
      return False, 't should be >= storage size.'
    if set(self._items.keys()) != set(self._distribution.ids()):
      return False, 'IDs in storage and distribution do not match.'
    return self._distribution.check_valid()


# ----------------------------------------------------------------------------
# R7: transition accumulators (host-side, insert time)
# ----------------------------------------------------------------------------


def n_step_fold(one_steps):
  """`replay.py:808-824` (`_build_n_step_transition`), float64 python scalars."""
  ret = 0.0
  disc = 1.0
  for tr in one_steps:
    ret += disc * tr.r_t
    disc *= tr.discount_t
  return Transition(one_steps[0].s_tm1, one_steps[0].a_tm1, ret, disc, one_steps[-1].s_t)


class NStepTransitionAccumulator:
  """`replay.py:827-892`.  n=1 behaves as `TransitionAccumulator` (`:771-805`)."""

  def __init__(self, n):
    self._window = collections.deque(maxlen=n)
    self.reset()

  def reset(self):
    self._window.clear()
    self._prev = None
    self._prev_action = None

  def step(self, timestep_t, a_t):
    out = []
    if timestep_t.first():
      self.reset()
    if self._prev is None:
      if not timestep_t.first():
        raise ValueError('Expected FIRST timestep, got %s.' % str(timestep_t))
      self._prev, self._prev_action = timestep_t, a_t
      return out
    self._window.append(Transition(self._prev.observation, self._prev_action, timestep_t.reward,
                                   timestep_t.discount, timestep_t.observation))
    self._prev, self._prev_action = timestep_t, a_t
    if timestep_t.last():
      while self._window:       # n, n-1, ..., 1-step, all ending at s_T (`:873-877`)
        out.append(n_step_fold(list(self._window)))
        self._window.popleft()
    elif len(self._window) == self._window.maxlen:
      out.append(n_step_fold(list(self._window)))
    return out


class TransitionAccumulator(NStepTransitionAccumulator):
  """`replay.py:771-805`: the 1-step special case (pinned by `replay_test.py:264-280`)."""

  def __init__(self):
    super().__init__(1)


# ----------------------------------------------------------------------------
# Synthetic replay contents shared by bench.py, tests and the CUDA fill kernel
# ----------------------------------------------------------------------------

_M64 = (1 << 64) - 1


def _mix64(x):
  """splitmix64 finaliser on uint64 numpy arrays (wraps mod 2**64)."""
  x = np.asarray(x, dtype=np.uint64)
  with np.errstate(over='ignore'):
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    x = x ^ (x >> np.uint64(31))
  return x


def synthetic_rows(seed, rows, obs_bytes, num_actions, discount=0.99):
  """Contents of synthetic transitions `rows` (array of ids), matching the CUDA
  kernel `dz_replay_fill_synthetic` bit for bit (SURVEY §8(d) synthetic inputs).

  obs bytes: 8-byte word w of observation o (0 = s_tm1, 1 = s_t) of row r is
  mix64(seed*0x9E3779B97F4A7C15 + (r*2+o)*(obs_bytes/8) + w), little endian.
  action = h % A; reward in {-1,0,1} with p = .05/.9/.05; discount = 0.99 w.p. .99 else 0.
  """
  rows = np.asarray(rows, dtype=np.uint64)
  words = obs_bytes // 8
  assert obs_bytes % 8 == 0
  with np.errstate(over='ignore'):
    base = np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
    w = np.arange(words, dtype=np.uint64)[None, None, :]
    o = np.arange(2, dtype=np.uint64)[None, :, None]
    ctr = base + (rows[:, None, None] * np.uint64(2) + o) * np.uint64(words) + w
    obs = _mix64(ctr).view(np.uint8).reshape(len(rows), 2, obs_bytes)
    h = _mix64(base + np.uint64(0xD1B54A32D192ED03) + rows * np.uint64(4))
    a = (h % np.uint64(num_actions)).astype(np.int32)
    h2 = _mix64(base + np.uint64(0xD1B54A32D192ED03) + rows * np.uint64(4) + np.uint64(1))
    u = (h2 >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    r = np.where(u < 0.05, -1.0, np.where(u < 0.95, 0.0, 1.0))
    h3 = _mix64(base + np.uint64(0xD1B54A32D192ED03) + rows * np.uint64(4) + np.uint64(2))
    u3 = (h3 >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    d = np.where(u3 < 0.99, discount, 0.0)
  return obs, a, r, d
