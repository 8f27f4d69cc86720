"""CPU: threefry2x32 / jax.random restatements against published known answers; the product's host key handling and the
host-compiled device function against the oracle."""

import ctypes as C

import numpy as np
import pytest

from oracle import jax_prng_oracle as jo

KATS = [((0x0, 0x0), (0x0, 0x0), (0x6b200159, 0x99ba4efe)),
        ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1cb996fc, 0xbb002be7)),
        ((0x13198a2e, 0x03707344), (0x243f6a88, 0x85a308d3), (0xc4923a9c, 0x483df7a0))]


def test_threefry_known_answers():
  """Random123 known-answer vectors (the ones jax's random_test.py::testThreefry2x32 asserts)."""
  for key, ctr, want in KATS:
    assert jo.threefry2x32(key, ctr) == want


def test_documented_jax_values():
  assert jo.prng_key(0) == (0, 0) and jo.prng_key(42) == (0, 42)
  assert jo.split(jo.prng_key(0)) == [(4146024105, 967050713), (2718843009, 1272950319)]
  u = jo.uniform(jo.prng_key(0), ())
  assert u.dtype == np.float32 and repr(float(u)).startswith('0.41845703')
  assert np.float32(u) == np.float32(0.41845703)


def test_uniform_range_and_odd_sizes():
  for n in (1, 2, 3, 7, 64, 2047, 2048):
    u = jo.uniform((123, 456), (n,))
    assert u.shape == (n,) and (u >= 0).all() and (u < 1).all()
  # an odd request is the even one with the zero counter appended and the last word dropped
  a = jo.random_bits((5, 6), 5)
  half = 3
  want = [jo.threefry2x32((5, 6), (i, [3, 4, 0][i]))[0] for i in range(half)] + \
         [jo.threefry2x32((5, 6), (i, [3, 4, 0][i]))[1] for i in range(half)][:2]
  assert a.tolist() == want


def test_host_compiled_device_function_matches_known_answers():
  from dqn_zoo_b200 import _lib
  out = (C.c_uint32 * 2)()
  rs = np.random.RandomState(0)
  cases = [(k, c) for k, c, _ in KATS] + [(tuple(int(x) for x in rs.randint(0, 2 ** 32, 2, dtype=np.uint64)),
                                           tuple(int(x) for x in rs.randint(0, 2 ** 32, 2, dtype=np.uint64))) for _ in range(200)]
  for key, ctr in cases:
    _lib.call('dz_test_threefry2x32', key[0], key[1], ctr[0], ctr[1], C.cast(out, C.c_void_p))
    assert (out[0], out[1]) == jo.threefry2x32(key, ctr)


def test_product_host_keys_match_oracle():
  from dqn_zoo_b200 import jax_prng as jp
  rs = np.random.RandomState(1)
  for _ in range(20):
    key = tuple(int(x) for x in rs.randint(0, 2 ** 32, 2, dtype=np.uint64))
    for num in (2, 3, 4):
      assert [tuple(int(v) for v in row) for row in jp.split(key, num)] == jo.split(key, num)
    for shape in ((), (1, 5), (32, 64), (7,)):
      np.testing.assert_array_equal(jp.uniform(key, shape), jo.uniform(key, shape))
  assert tuple(jp.prng_key(42)) == jo.prng_key(42)
  key = jo.prng_key(7)
  new_o, t0, t1, t2 = jo.iqn_update_taus(key, 4, 8, 5, 7)
  new_p, sample = jp.iqn_update_keys(np.array(key, dtype=np.uint32))
  assert tuple(int(v) for v in new_p) == new_o
  for k, t in zip(sample, (t0, t1, t2)):
    np.testing.assert_array_equal(jp.uniform(k, t.shape), t)
  new_o, ta = jo.iqn_act_taus(key, 6)
  new_p, sample = jp.iqn_act_keys(np.array(key, dtype=np.uint32))
  assert tuple(int(v) for v in new_p) == new_o
  np.testing.assert_array_equal(jp.uniform(sample[0], (1, 6)), ta)
