"""CPU: the preprocessing oracle against its pins — Pillow (run here), numpy's tensordot (run here) and the
reference's golden vector (processors_test.py:405-475)."""

import hashlib
from fractions import Fraction

import numpy as np
import pytest

from oracle import processors_oracle as po

GOLDEN_INPUT_HASHES = [
    '250557b2184381fc2ec541fc313127050098fce825a6e98a728c2993874db300',
    'db8054ca287971a0e1264bfbc5642233085f1b27efbca9082a29f5be8a24c552',
    '7016e737a257fcdb77e5f23daf96d94f9820bd7361766ca7b1401ec90984ef71',
    '356dfcf0c6eaa4e2b5e80f4611375c0131435cc22e6a413b573818d7d084e9b2',
    '73078bedd438422ad1c3dda6718aa1b54f6163f571d2c26ed714c515a6372159',
]
GOLDEN_OUTPUT_HASH = '0d158a8f45aa09aa6fad0354d2eb1fc0e3f57add88e772f3b71f54819d8200aa'


def golden_inputs():
  rs = np.random.RandomState(seed=1)
  return [rs.randint(0, 256, size=(210, 160, 3), dtype=np.uint8) for _ in range(5)]


def test_reference_golden_vector():
  """processors_test.py:405-475: five fixed timesteps through processors.atari(); hash of the stacked observation."""
  rgb = golden_inputs()
  assert [hashlib.sha256(o).hexdigest() for o in rgb] == GOLDEN_INPUT_HASHES
  p = po.AtariPreprocessor()
  steps = [(po.FIRST, None, None), (po.MID, 0.5, 0.9), (po.MID, 0.2, 0.9), (po.MID, 0, 0.9), (po.MID, 0.1, 0.9)]
  outs = [p(st, r, d, (o, 3)) for (st, r, d), o in zip(steps, rgb)]
  assert outs[0] is not None and outs[0][0] == po.FIRST and outs[0][1] is None and outs[0][2] is None
  assert outs[1] is None and outs[2] is None and outs[3] is None
  step_type, reward, discount, obs = outs[4]
  assert step_type == po.MID
  assert reward == pytest.approx(0.5 + 0.2 + 0.0 + 0.1)
  assert discount == pytest.approx(0.9 ** 4 * 0.99)
  assert obs.shape == (84, 84, 4) and obs.dtype == np.uint8
  assert hashlib.sha256(obs.flatten()).hexdigest() == GOLDEN_OUTPUT_HASH


@pytest.mark.parametrize('shape', [(210, 160, 84, 84), (210, 160, 110, 84), (100, 100, 84, 84), (64, 48, 84, 84),
                                   (210, 160, 42, 42), (84, 84, 84, 84), (250, 160, 84, 84), (17, 23, 5, 7)])
def test_resize_matches_pillow(shape):
  Image = pytest.importorskip('PIL.Image')
  h, w, oh, ow = shape
  rs = np.random.RandomState(h * w + oh)
  for _ in range(3):
    img = rs.randint(0, 256, size=(h, w), dtype=np.uint8)
    want = np.array(Image.fromarray(img).resize((ow, oh), Image.Resampling.BILINEAR), dtype=np.uint8)
    assert np.array_equal(po.resize_bilinear_u8(img, oh, ow), want)
  for const in (0, 255):   # saturation: fixed-point weights sum to 1 << 22 +- a few units
    img = np.full((h, w), const, dtype=np.uint8)
    want = np.array(Image.fromarray(img).resize((ow, oh), Image.Resampling.BILINEAR), dtype=np.uint8)
    assert np.array_equal(po.resize_bilinear_u8(img, oh, ow), want)


def test_rgb2y_is_the_plain_left_to_right_float64_sum():
  """Exact rational check of the canonical rounding order on a sample that includes every rounding-order case of
  one red plane, plus the statement about numpy in this container (differences only at near-integer lumas)."""
  g, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing='ij')
  w = [Fraction(x) for x in po.LUMA]
  total_diff = 0
  for r in (0, 1, 77, 128, 200, 255):
    arr = np.stack([np.full_like(g, r), g, b], axis=-1)
    got = po.rgb2y(arr)
    blas = np.tensordot(arr, po.LUMA, (-1, 0)).astype(np.uint8)
    diff = np.argwhere(got != blas)
    total_diff += len(diff)
    for i, j in diff:      # where the BLAS order disagrees, the exact luma is (numerically) an integer
      exact = r * w[0] + int(g[i, j]) * w[1] + int(b[i, j]) * w[2]
      assert abs(exact - round(exact)) < Fraction(1, 10 ** 9)
    for i, j in list(diff[:50]) + [(0, 0), (255, 255), (13, 200)]:
      gg, bb = int(g[i, j]), int(b[i, j])
      t = float(Fraction(float(r * w[0])) + Fraction(float(gg * w[1])))     # fl(fl(r wr) + fl(g wg))
      t = float(Fraction(t) + Fraction(float(bb * w[2])))
      assert int(t) == int(got[i, j])
  assert total_diff < 6 * 65536 // 500


def test_pooling_uses_the_last_two_slots_and_zero_padding():
  rs = np.random.RandomState(3)
  a = rs.randint(0, 256, size=(210, 160, 3), dtype=np.uint8)
  b = rs.randint(0, 256, size=(210, 160, 3), dtype=np.uint8)
  assert np.array_equal(po.pooled_gray_resized(a, b), po.resize_bilinear_u8(po.rgb2y(np.maximum(a, b)), 84, 84))
  assert np.array_equal(po.pooled_gray_resized(None, b), po.resize_bilinear_u8(po.rgb2y(b), 84, 84))


def test_episode_cadence_and_last_padding():
  """Diagram of processors.py:432-437: F | M M M M | M M L -> outputs at F, at the 4th M, and at L (padded)."""
  rs = np.random.RandomState(4)
  frames = [rs.randint(0, 256, size=(32, 24, 3), dtype=np.uint8) for _ in range(8)]
  p = po.AtariPreprocessor(resize_shape=(8, 8))
  types = [po.FIRST] + [po.MID] * 6 + [po.LAST]
  outs = []
  for t, f in zip(types, frames):
    outs.append(p(t, None if t == po.FIRST else 1.0, None if t == po.FIRST else (0.0 if t == po.LAST else 1.0), (f, 3)))
  emitted = [i for i, o in enumerate(outs) if o is not None]
  assert emitted == [0, 4, 7]
  assert outs[4][0] == po.MID and outs[4][1] == 1.0 and outs[4][2] == pytest.approx(0.99)   # sum 4 clipped to 1
  assert outs[7][0] == po.LAST and outs[7][2] == 0.0
  # the LAST buffer is [M, M, L, pad]: the pooled pair is (L frame, zeros)
  want_last = po.pooled_gray_resized(frames[7], None, 8, 8)
  assert np.array_equal(outs[7][3][..., 2], want_last)
  assert np.array_equal(outs[7][3][..., 0], po.pooled_gray_resized(None, frames[0], 8, 8))
  assert np.array_equal(outs[7][3][..., 3], np.zeros((8, 8), np.uint8))
  with pytest.raises(RuntimeError):
    p(po.MID, 0.0, 1.0, (frames[0], 3))
  p.reset()
  assert p(po.FIRST, None, None, (frames[0], 3)) is not None


def test_life_loss_zeroes_the_discount():
  rs = np.random.RandomState(5)
  f = rs.randint(0, 256, size=(32, 24, 3), dtype=np.uint8)
  p = po.AtariPreprocessor(resize_shape=(8, 8))
  p(po.FIRST, None, None, (f, 3))
  outs = [p(po.MID, 0.0, 1.0, (f, 3 if i != 2 else 2)) for i in range(4)]
  assert outs[3][2] == 0.0


# ---- scalar half: the scenarios of processors_test.py, through the oracle AND the product's host state machine ----

def _machines(**kwargs):
  """(name, feed(step_type, reward, discount, lives) -> None | (type, reward, discount)) for both implementations.

  The product class (dqn_zoo_b200.processors) only touches CUDA when it sees pixels; its emission rule and scalar
  aggregation are plain host code and are driven here directly."""
  from dqn_zoo_b200 import processors as dev

  frame = np.zeros((8, 16, 3), np.uint8)
  oracle = po.AtariPreprocessor(resize_shape=(4, 4), **kwargs)

  def feed_oracle(st, r, d, lives=3):
    out = oracle(st, r, d, (frame, lives))
    return None if out is None else out[:3]

  product = dev.BatchedAtariPreprocessor(num_streams=1, resize_shape=(4, 4), **kwargs)
  stream = product._streams[0]

  def feed_product(st, r, d, lives=3):
    # the part of BatchedAtariPreprocessor.step() that precedes the pixel upload / kernel launch
    if product._life_loss:
      lost = st == po.MID and lives < stream.lives
      stream.lives = lives
      if lost:
        d = 0.0
    if stream.index >= product._repeats:
      stream.index = 0
      stream.slots = [None] * product._repeats
    stream.slots[stream.index] = (dev.StepType(st), r, d)
    stream.index += 1
    if not product._should_emit(stream):
      return None
    t, rr, dd = product._reduce_scalars(stream)
    return int(t), rr, dd

  def reset_both():
    oracle.reset()
    stream.reset()

  return [('oracle', feed_oracle), ('product', feed_product)], reset_both


def test_emission_cadence_table():
  """processors_test.py:74-93: F emits at once, then every 4th MID, and LAST emits immediately."""
  machines, _ = _machines()
  seq = [(0, True)] + [(1, False), (1, False), (1, False), (1, True)] * 2 + [(1, False), (2, True)]
  for name, feed in machines:
    for st, expected in seq:
      out = feed(st, None if st == 0 else 0.0, None if st == 0 else 1.0)
      assert (out is not None) == expected, (name, st)


def test_errors_without_reset_and_with_two_boundaries():
  """processors_test.py:95-137."""
  for kwargs in ({}, {'num_action_repeats': 3}):
    machines, _ = _machines(**kwargs)
    for name, feed in machines:
      feed(0, None, None)
      feed(1, 0.0, 1.0)
      feed(2, 0.0, 0.0)
      with pytest.raises(RuntimeError, match='Should have reset'):
        feed(0, None, None)
  machines, _ = _machines(num_action_repeats=3)
  for name, feed in machines:            # [F, M, F] inside one buffer
    feed(0, None, None)                  # slot 2 -> emitted, buffer restarts
    feed(0, None, None)                  # slot 0: a second FIRST without a LAST in between is accepted alone ...
    feed(1, 0.0, 1.0)
    with pytest.raises(RuntimeError, match='at most one FIRST or LAST'):
      feed(2, 0.0, 0.0)                  # ... but FIRST and LAST in the same buffer are not


@pytest.mark.parametrize('types,discounts,lives,expected', [
    ('fmmmmm', 'n11111', '333333', 'n11111'),
    ('fmmmmm', 'n11111', '333222', 'n11011'),
    ('fmmmmm', 'n11111', '332211', 'n10101'),
])
def test_life_loss_zeroes_exactly_the_losing_step(types, discounts, lives, expected):
  """processors_test.py:204-253 (ZeroDiscountOnLifeLoss), observed through the emitted discount products."""
  machines, _ = _machines(num_action_repeats=2, additional_discount=1.0, max_abs_reward=None)
  code = {'f': 0, 'm': 1, 'l': 2}
  for name, feed in machines:
    got = []
    for t, d, lv in zip(types, discounts, lives):
      out = feed(code[t], None if t == 'f' else 8.0, None if d == 'n' else float(d), int(lv))
      if out is not None:
        got.append(out[2])
    # repeats of 2: emissions at f, then after each pair of m; the product over a pair is 0 iff a loss fell in it
    want = [None]
    per_step = [None if e == 'n' else float(e) for e in expected][1:]
    for i in range(0, len(per_step) - 1, 2):
      want.append(per_step[i] * per_step[i + 1])
    assert got == want, (name, got, want)


@pytest.mark.parametrize('rewards,clip,expected', [([1, 2, 3, 0], None, 6), ([1, -2, 3, 0], None, 2), ([1, 2, 3, 0], 2, 2),
                                                   ([-1, -2, 0.5, 0], 2, -2), ([0.5, 0.2, 0, 0.1], 1.0, 0.8)])
def test_reward_sum_then_clip(rewards, clip, expected):
  """processors_test.py:281-349: rewards are summed over the action repeats, then clipped."""
  machines, _ = _machines(max_abs_reward=clip)
  for name, feed in machines:
    feed(0, None, None)
    out = None
    for r in rewards:
      out = feed(1, r, 1.0)
    assert out is not None and out[0] == po.MID
    assert out[1] == pytest.approx(expected), name
    assert out[2] == pytest.approx(0.99)


def test_first_has_no_reward_or_discount_and_last_is_reported():
  """processors_test.py:255-279 (reduce_step_type) via the pipeline: 000F -> FIRST, MML0 -> LAST, MMMM -> MID."""
  machines, _ = _machines()
  for name, feed in machines:
    out = feed(0, None, None)
    assert out == (po.FIRST, None, None), name
    for _ in range(3):
      assert feed(1, 1.0, 1.0) is None
    assert feed(1, 1.0, 1.0)[0] == po.MID
    feed(1, 1.0, 1.0)
    feed(1, 1.0, 1.0)
    out = feed(2, 1.0, 0.0)
    assert out[0] == po.LAST and out[2] == 0.0 and out[1] == 1.0   # 3 summed, clipped to 1


def test_vectorized_scalar_state_machine_equals_the_per_stream_one():
  """VectorScalars (numpy array code over n streams) against the per-stream host state machine of
  BatchedAtariPreprocessor.step() on desynchronised random episodes: FIRST/LAST boundaries, idle ticks, life losses,
  resets after LAST, None rewards/discounts of FIRST timesteps — every emission identical (type, reward, discount, which
  pooled frames exist, stack fill)."""
  from dqn_zoo_b200 import processors as dev
  n, R = 9, 4
  rs = np.random.RandomState(4)
  product = dev.BatchedAtariPreprocessor(num_streams=n, resize_shape=(4, 4), additional_discount=0.99, max_abs_reward=1.0)
  vs = dev.VectorScalars(n, R, 0.99, 1.0, True, 4)
  FIRST, MID, LAST = int(dev.StepType.FIRST), int(dev.StepType.MID), int(dev.StepType.LAST)
  need_first = np.ones(n, bool)
  lives = np.full(n, 3)
  emissions = 0
  for tick in range(600):
    active = rs.uniform(size=n) < 0.85
    st = np.full(n, MID); rw = np.zeros(n); dc = np.ones(n)
    for e in range(n):
      if not active[e]:
        continue
      if need_first[e]:
        st[e], rw[e], dc[e], lives[e] = FIRST, np.nan, np.nan, 3
        need_first[e] = False
      else:
        st[e] = LAST if rs.uniform() < 0.03 else MID
        rw[e] = float(rs.choice([0.0, 1.0, -1.0, 2.5, -3.0]))
        dc[e] = 0.0 if st[e] == LAST else 1.0
        if rs.uniform() < 0.05 and lives[e] > 0:
          lives[e] -= 1
    out = vs.tick(st, rw, dc, lives, active)
    for e in range(n):
      if not active[e]:
        assert not out['emit'][e]
        continue
      s = product._streams[e]
      r = None if np.isnan(rw[e]) else float(rw[e])
      d = None if np.isnan(dc[e]) else float(dc[e])
      if product._life_loss:
        lost = st[e] == MID and lives[e] < s.lives
        s.lives = int(lives[e])
        if lost:
          d = 0.0
      if s.index >= R:
        s.index = 0; s.slots = [None] * R; s.has_frame = [False] * R
      s.slots[s.index] = (dev.StepType(int(st[e])), r, d)
      if s.index - (R - 2) >= 0:
        s.has_frame[s.index] = True
      assert out['pooled_slot'][e] == s.index - (R - 2) if s.index - (R - 2) >= 0 else out['pooled_slot'][e] < 0
      s.index += 1
      emit = product._should_emit(s)
      assert bool(out['emit'][e]) == emit, (tick, e)
      if emit:
        emissions += 1
        t, rr, dd = product._reduce_scalars(s)
        assert int(t) == int(out['step_type'][e])
        assert (rr is None and np.isnan(out['reward'][e])) or rr == out['reward'][e]
        assert (dd is None and np.isnan(out['discount'][e])) or dd == out['discount'][e]
        assert bool(out['a_ok'][e]) == (s.has_frame[R - 2] and s.slots[R - 2] is not None)
        assert bool(out['b_ok'][e]) == (s.has_frame[R - 1] and s.slots[R - 1] is not None)
        assert int(out['count'][e]) == s.count
        s.count = min(s.count + 1, 4)
        if int(t) == LAST:                    # the run loop resets the processor after a LAST timestep
          s.reset(); vs.reset([e]); need_first[e] = True
  assert emissions > 800
