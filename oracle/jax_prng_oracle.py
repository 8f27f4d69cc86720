"""CPU restatement of the jax.random primitives the reference's IQN agent uses (iqn/agent.py:45-50, 182-190, 207, 222).
TEST INFRASTRUCTURE ONLY (imported by tests/ and smoke()).

jax (pinned 0.3.10, docker_requirements.txt:15) is a third-party dependency that is not under /root/reference and is
not installable here, so the algorithm is restated from its published definition and PINNED to published known answers
(tests/test_jax_prng.py):
  * threefry2x32, 20 rounds (Salmon et al. 2011; jax/_src/prng.py `threefry_2x32`): the three Random123 known-answer
    vectors that jax's own test-suite checks — (key 0 0, ctr 0 0) -> 6b200159 99ba4efe; (all ones) -> 1cb996fc bb002be7;
    (13198a2e 03707344; 243f6a88 85a308d3) -> c4923a9c 483df7a0;
  * `jax.random.split(PRNGKey(0))` = [[4146024105, 967050713], [2718843009, 1272950319]] and
    `jax.random.uniform(PRNGKey(0))` = 0.41845703 (values printed in the jax documentation).
Not restated (cannot be pinned without jax): truncated_normal (XLA's erf_inv), categorical/epsilon-greedy sampling,
Haiku's `next_rng_key` order.

Everything here is deliberately scalar/naive Python over uint32 so that it shares no code with the product."""

import numpy as np

M32 = 0xFFFFFFFF
ROTATIONS = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
  return ((x << r) | (x >> (32 - r))) & M32


def threefry2x32(key, counter):
  """One block: key (k0, k1), counter (c0, c1) -> (o0, o1), all Python ints < 2^32."""
  k0, k1 = int(key[0]) & M32, int(key[1]) & M32
  ks = (k0, k1, k0 ^ k1 ^ 0x1BD11BDA)
  x0 = (int(counter[0]) + ks[0]) & M32
  x1 = (int(counter[1]) + ks[1]) & M32
  for i in range(5):
    for r in ROTATIONS[i % 2]:
      x0 = (x0 + x1) & M32
      x1 = _rotl(x1, r)
      x1 ^= x0
    x0 = (x0 + ks[(i + 1) % 3]) & M32
    x1 = (x1 + ks[(i + 2) % 3] + i + 1) & M32
  return x0, x1


def random_bits(key, n):
  """jax `threefry_random_bits(key, 32, (n,))` = threefry_2x32(key, iota(n)): counters split into two halves."""
  counts = list(range(n))
  odd = n % 2
  if odd:
    counts.append(0)
  half = len(counts) // 2
  lo, hi = [], []
  for i in range(half):
    a, b = threefry2x32(key, (counts[i], counts[i + half]))
    lo.append(a)
    hi.append(b)
  out = lo + hi
  if odd:
    out = out[:-1]
  return np.array(out, dtype=np.uint32)


def prng_key(seed):
  """jax.random.PRNGKey for a non-negative 32-bit seed: [0, seed] (x64 disabled)."""
  return (0, int(seed) & M32)


def split(key, num=2):
  bits = random_bits(key, 2 * num)
  return [(int(bits[2 * i]), int(bits[2 * i + 1])) for i in range(num)]


def uniform(key, shape):
  """jax.random.uniform(key, shape, float32, 0, 1)."""
  n = int(np.prod(shape)) if len(shape) else 1
  bits = random_bits(key, n)
  floats = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
  return np.maximum(np.float32(0.0), floats).reshape(shape)


def iqn_update_taus(rng_key, batch, n_tm1, n_policy, n_t):
  """iqn/agent.py:207 + 182-190: -> (new agent key, tau_tm1, tau_t_selector, tau_t)."""
  new_key, update_key = split(rng_key, 2)
  _, k0, k1, k2 = split(update_key, 4)
  return new_key, uniform(k0, (batch, n_tm1)), uniform(k1, (batch, n_policy)), uniform(k2, (batch, n_t))


def iqn_act_taus(rng_key, n_policy):
  """iqn/agent.py:220-222: -> (new agent key, tau_t [1, n_policy])."""
  new_key, sample_key, _apply_key, _policy_key = split(rng_key, 4)
  return new_key, uniform(sample_key, (1, n_policy))
