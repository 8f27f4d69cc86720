"""Replay contract checks, written once and run against (a) the CPU oracle and
(b) the CUDA-backed `dqn_zoo_b200.replay`.  They restate what the reference's
`replay_test.py` pins (cited per function) plus equality with the golden
vectors produced from the reference itself (oracle/gen_golden.py).
"""

import copy
import os

import numpy as np
import pytest

from oracle import scenarios

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# weights go through pow(); CUDA's double pow is not correctly rounded, so the
# device path is allowed 4 ulp there.  Everything else must be bit-exact.
WEIGHT_RTOL = {'oracle': 0.0, 'device': 1e-15}


def check_scenario(lib, name, kind):
  want = np.load(os.path.join(GOLDEN, name + '.npz'))
  got = scenarios.ALL[name](lib)
  assert set(got.keys()) == set(want.files)
  for k in want.files:
    g, w = np.asarray(got[k]), want[k]
    assert g.shape == w.shape, (name, k, g.shape, w.shape)
    if k in ('weights',) and WEIGHT_RTOL[kind] > 0:
      np.testing.assert_allclose(g, w, rtol=WEIGHT_RTOL[kind], atol=0, err_msg='%s/%s' % (name, k))
    else:
      np.testing.assert_array_equal(g, w, err_msg='%s/%s' % (name, k))


# --- SumTree: replay_test.py:820-1045 -------------------------------------------------------


def sumtree_empty(lib):
  t = lib.SumTree()
  assert t.check_valid()[0]
  assert t.size == 0 and np.isnan(t.root())
  t.resize(0)
  assert np.isnan(t.root())
  t.resize(1)
  assert t.check_valid()[0] and t.root() == 0


def sumtree_resize_semantics(lib):
  t = lib.SumTree()
  t.resize(3)
  assert t.size == 3
  for i in range(3):
    assert t.get([i])[0] == 0
  t = lib.SumTree()
  t.set_all([4.0, 5.0, 3.0, 2.0])
  assert t.capacity == 4
  t = lib.SumTree()
  t.set_all([4.0, 5.0, 3.0, 2.0, 9])
  assert t.capacity == 8
  vals = [4.0, 5.0, 3.0]
  t = lib.SumTree()
  t.set_all(vals)
  np.testing.assert_array_equal(vals, np.asarray(t.values))
  t.resize(8)  # grow: keep, zero the rest (:880-890)
  np.testing.assert_array_equal(vals + [0.0] * 5, np.asarray(t.values))
  assert t.check_valid()[0]
  t = lib.SumTree()
  vals = [4.0, 5.0, 3.0, 8.0, 2.0]
  t.set_all(vals)
  t.resize(3)  # shrink (:892-901)
  np.testing.assert_array_equal(vals[:3], np.asarray(t.values))
  assert t.check_valid()[0] and t.root() == 12.0
  t = lib.SumTree()
  t.set_all(vals)
  t.resize(7)  # between size and capacity (:903-911)
  np.testing.assert_array_equal(vals + [0.0, 0.0], np.asarray(t.values))
  assert t.check_valid()[0]


def sumtree_get_set(lib):
  t = lib.SumTree()
  t.resize(3)
  for bad in (-1, 3):
    with pytest.raises(IndexError):
      t.get([bad])
  t = lib.SumTree()
  t.set_all([4.0, 5.0, 3.0, 9.0])
  np.testing.assert_array_equal([5.0, 9.0], np.asarray(t.get([1, 3])))
  t.set([2], [99])
  np.testing.assert_array_equal([4, 5, 99, 9], np.asarray(t.values))
  t.set([2, 0], [7, 88])
  np.testing.assert_array_equal([88, 5, 7, 9], np.asarray(t.values))
  t.set([1, 1], [1.0, 2.0])  # duplicates: last write wins (numpy fancy assign)
  np.testing.assert_array_equal([88, 2, 7, 9], np.asarray(t.values))
  assert t.root() == 106.0 and t.check_valid()[0]


QUERY_TABLE = [(0, 0.0), (0, 3.0 - 0.1), (1, 3.0), (1, 4.0 - 0.1), (2, 4.0), (2, 6.0 - 0.1), (3, 6.0),
               (3, 11.0 - 0.1)]  # replay_test.py:939-953


def sumtree_query_known_answers(lib):
  t = lib.SumTree()
  t.set_all([3.0, 1.0, 2.0, 5.0])
  for want, target in QUERY_TABLE:
    assert list(t.query([target])) == [want]
  np.testing.assert_array_equal([0, 1, 2], np.asarray(t.query([2.9, 3.0, 4])))
  for bad in (-1.0, 11.0, 12.0, t.root()):
    with pytest.raises(ValueError):
      t.query([bad])
  assert abs(t.root() - 11.0) < 1e-12


def sumtree_never_returns_zero_leaf(lib):
  vals = np.array([0, 1, 0, 0, 3, 0, 2, 0, 3, 0], dtype=np.float64)  # replay_test.py:978-987
  t = lib.SumTree()
  t.set_all(vals)
  zero = set(np.nonzero(vals == 0)[0].tolist())
  for target in [0, 0.1, 0.9, 1, 1.1, 3.9, 4, 4.1, 5.9, 6, 6.1, 8.9, 8.999999]:
    assert t.query([target])[0] not in zero


def sumtree_rejects_bad_values(lib):
  t = lib.SumTree()
  t.set_all([0, 1, 2])
  for bad in (-1, np.nan, np.inf):
    with pytest.raises(ValueError):
      t.set([1], [bad])
    with pytest.raises(ValueError):
      t.set_all([1, bad])


def sumtree_matches_naive_prefix_sums(lib, seeds=range(10)):
  """NaiveSumTree equivalence, replay_test.py:1048-1161: O(n) cumsum oracle."""
  for seed in seeds:
    rs = np.random.RandomState(seed)
    t = lib.SumTree()
    leaves = np.zeros(0)
    for _ in range(40):
      op = rs.randint(4)
      if op == 0 or len(leaves) == 0:
        leaves = np.abs(rs.standard_cauchy(int(rs.randint(1, 50))))
        t.set_all(leaves)
      elif op == 1:
        n = int(rs.randint(1, 60))
        leaves = np.concatenate([leaves[:n], np.zeros(max(0, n - len(leaves)))])
        t.resize(n)
      elif op == 2:
        k = int(rs.randint(1, 8))
        idx = rs.randint(len(leaves), size=k)
        vals = np.abs(rs.standard_cauchy(k))
        leaves = leaves.copy()
        leaves[idx] = vals
        t.set(idx, vals)
      else:
        st = t.get_state()
        t2 = lib.SumTree()
        t2.set_state({k: (np.array(v) if isinstance(v, np.ndarray) else v) for k, v in st.items()})
        t = t2
      np.testing.assert_array_equal(leaves, np.asarray(t.values))
      ok, msg = t.check_valid()
      assert ok, msg
      if leaves.sum() > 0:
        # exact-arithmetic targets: integers-only leaves would be exact; here just check the
        # defining property against float cumsum with a tolerance-free formulation:
        targets = rs.uniform(size=6) * t.root()
        got = np.asarray(t.query(targets))
        csum = np.cumsum(leaves)
        for g, tg in zip(got, targets):
          assert leaves[g] > 0
          # prefix before g <= target (+eps) < prefix through g (+eps)
          lo = csum[g] - leaves[g]
          assert lo <= tg + 1e-9 * csum[-1] and tg < csum[g] + 1e-9 * csum[-1]


# --- PrioritizedDistribution: replay_test.py:429-744 ------------------------------------------


def distribution_ids_order(lib):
  """replay_test.py:437-468: ids() order after add/remove is [2,3,5,4,7,1,0]-style dict order."""
  d = lib.PrioritizedDistribution(priority_exponent=1.0, uniform_sample_probability=0.0,
                                  random_state=np.random.RandomState(1), min_capacity=0, max_capacity=None)
  d.add_priorities([2, 3, 5], [1.0, 1.0, 1.0])
  d.add_priorities([4, 7], [1.0, 1.0])
  d.add_priorities([1, 0], [1.0, 1.0])
  assert list(d.ids()) == [2, 3, 5, 4, 7, 1, 0]
  with pytest.raises(IndexError):
    d.add_priorities([3], [1.0])
  with pytest.raises(IndexError):
    d.update_priorities([99], [1.0])
  assert d.check_valid()[0]


def distribution_zero_priorities_sample_uniformly(lib):
  """replay_test.py:662-667."""
  d = lib.PrioritizedDistribution(priority_exponent=1.0, uniform_sample_probability=0.0,
                                  random_state=np.random.RandomState(1), min_capacity=4, max_capacity=4)
  d.add_priorities([0, 1, 2, 3], [0.0, 0.0, 0.0, 0.0])
  ids, probs = d.sample(200)
  assert set(np.asarray(ids).tolist()) == {0, 1, 2, 3}
  np.testing.assert_array_equal(np.asarray(probs), np.full(200, 0.25))


def distribution_sample_statistics(lib):
  """replay_test.py:669-697: empirical frequencies follow (1-usp) p/sum + usp/n, rtol 1e-2."""
  rs = np.random.RandomState(1)
  d = lib.PrioritizedDistribution(priority_exponent=1.0, uniform_sample_probability=0.3, random_state=rs,
                                  min_capacity=5, max_capacity=5)
  pri = np.array([1.0, 2.0, 3.0, 4.0, 10.0])
  d.add_priorities(list(range(5)), pri)
  counts = np.zeros(5)
  for _ in range(4):
    ids, probs = d.sample(50000)
    counts += np.bincount(np.asarray(ids), minlength=5)
  want = 0.7 * pri / pri.sum() + 0.3 / 5
  np.testing.assert_allclose(counts / counts.sum(), want, rtol=2e-2)


# --- PrioritizedTransitionReplay / TransitionReplay: replay_test.py:39-174, 747-817 -----------


def _tiny_item(lib, k):
  o = np.full(scenarios.OBS_SHAPE, k % 251, dtype=np.uint8)
  return lib.Transition(s_tm1=o, a_tm1=k % 6, r_t=float(k), discount_t=0.5, s_t=o + 1)


def per_eviction_keeps_newest(lib):
  """replay_test.py:762-772: after 2*capacity adds only the newest `capacity` ids remain."""
  cap = 10
  rep = lib.PrioritizedTransitionReplay(cap, lib.Transition(None, None, None, None, None), 1.0, lambda t: 1.0, 0.0,
                                        False, np.random.RandomState(1))
  for k in range(2 * cap):
    rep.add(_tiny_item(lib, k), priority=1.0 + k)
  assert rep.size == cap and rep.capacity == cap
  st = rep.get_state()
  assert [i for i, _ in st['storage']] == list(range(cap, 2 * cap))
  tr, ids, w = rep.sample(64)
  assert np.asarray(ids).min() >= cap and np.asarray(ids).max() < 2 * cap
  np.testing.assert_array_equal(np.asarray(tr.r_t), np.asarray(ids, dtype=np.float64))
  assert tr.s_tm1.shape == (64,) + scenarios.OBS_SHAPE and tr.s_tm1.dtype == np.uint8
  assert np.asarray(tr.a_tm1).dtype == np.int64 and np.asarray(tr.r_t).dtype == np.float64
  assert rep.check_valid()[0]


def per_state_roundtrip(lib):
  """replay_test.py:774-787 style: set_state(get_state()) continues identically."""
  def make(seed):
    return lib.PrioritizedTransitionReplay(12, lib.Transition(None, None, None, None, None), 0.5, lambda t: 0.7, 0.2,
                                           True, np.random.RandomState(seed))
  a = make(3)
  for k in range(30):
    a.add(_tiny_item(lib, k), priority=1.0 + (k % 5))
  _, ids, _ = a.sample(8)
  a.update_priorities(ids, np.linspace(0.0, 3.0, 8).astype(np.float32))
  st = copy.deepcopy(a.get_state())  # the reference hands out live references
  b = make(99)
  b.set_state(st)
  # same RNG stream for both from here on
  sa, sb = np.random.RandomState(5), np.random.RandomState(5)
  _rebind_rng(a, sa)
  _rebind_rng(b, sb)
  for k in range(30, 40):
    a.add(_tiny_item(lib, k), priority=2.0)
    b.add(_tiny_item(lib, k), priority=2.0)
    ta, ia, wa = a.sample(8)
    tb, ib, wb = b.sample(8)
    np.testing.assert_array_equal(np.asarray(ia), np.asarray(ib))
    np.testing.assert_array_equal(np.asarray(wa), np.asarray(wb))
    np.testing.assert_array_equal(ta.s_t, tb.s_t)
  assert b.check_valid()[0]


def _rebind_rng(rep, rs):
  """Points a replay (oracle or device) and its distribution at a new RandomState."""
  for obj in (rep, getattr(rep, '_distribution', None)):
    if obj is None:
      continue
    for name in ('_random_state', '_rs'):
      if hasattr(obj, name):
        setattr(obj, name, rs)


def uniform_swap_remove_permutation(lib):
  """SURVEY §8(a) R6 probe: C=6 after 10 adds the distribution's id list is [5,6,7,8,4,9]; the
  storage ids stay contiguous and sorted (replay_test.py:129-147)."""
  rep = lib.TransitionReplay(6, lib.Transition(None, None, None, None, None), np.random.RandomState(1))
  for k in range(10):
    rep.add(_tiny_item(lib, k))
  st = rep.get_state()
  assert list(st['distribution']['ids']) == [5, 6, 7, 8, 4, 9]
  assert sorted(st['distribution']['ids']) == list(range(4, 10))
  assert list(rep.ids()) == list(range(4, 10))
  tr = rep.sample(32)
  assert tr.s_tm1.shape == (32,) + scenarios.OBS_SHAPE
  assert rep.check_valid()[0]


def nstep_known_answers(lib):
  """replay_test.py:209-244: n=3 returns/discount products; :282-323 LAST flush."""
  ts = scenarios._TS
  acc = lib.NStepTransitionAccumulator(3)
  r = [None, 0.5, 1.0, -2.0, 4.0]
  d = [None, 0.9, 0.8, 0.7, 0.0]
  out = []
  for t in range(5):
    st = 0 if t == 0 else (2 if t == 4 else 1)
    out.append(list(acc.step(ts(st, r[t], d[t], t), a_t=10 + t)))
  assert out[0] == [] and out[1] == [] and out[2] == []
  (tr,) = out[3]
  assert (tr.s_tm1, tr.a_tm1, tr.s_t) == (0, 10, 3)
  assert tr.r_t == 0.0 + 1.0 * 0.5 + 0.9 * 1.0 + (0.9 * 0.8) * -2.0
  assert tr.discount_t == 1.0 * 0.9 * 0.8 * 0.7
  flush = out[4]
  assert [(x.s_tm1, x.s_t) for x in flush] == [(1, 4), (2, 4), (3, 4)]
  assert flush[2].r_t == 4.0 and flush[2].discount_t == 0.0
  assert flush[0].r_t == 0.0 + 1.0 + 0.8 * -2.0 + (0.8 * 0.7) * 4.0
  with pytest.raises(ValueError):
    lib.NStepTransitionAccumulator(2).step(ts(1, 0.0, 1.0, 0), 0).__iter__().__next__()


CONTRACT = [sumtree_empty, sumtree_resize_semantics, sumtree_get_set, sumtree_query_known_answers,
            sumtree_never_returns_zero_leaf, sumtree_rejects_bad_values, sumtree_matches_naive_prefix_sums,
            distribution_ids_order, distribution_zero_priorities_sample_uniformly, distribution_sample_statistics,
            per_eviction_keeps_newest, per_state_roundtrip, uniform_swap_remove_permutation, nstep_known_answers]
