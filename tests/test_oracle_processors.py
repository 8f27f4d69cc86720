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
