"""JAX-compatible key handling for the tau samples of the IQN agent (iqn/agent.py:45-50,182-190).

Keys are split on the host (a handful of threefry blocks per learner step — vectorised numpy over uint32); the
draws themselves are produced on the device by `dz_jax_uniform` (csrc/dz_jaxprng.cu) from keys that sit in device
memory, so the launch can be part of the captured CUDA graph of the learner step.  Pinned to published known answers
of threefry2x32 / jax.random (tests/test_jax_prng.py); what cannot be pinned without jax (truncated normal,
epsilon-greedy sampling, Haiku key order) is not claimed."""

import ctypes as C
from typing import Sequence, Tuple

import numpy as np
import torch

from dqn_zoo_b200 import _lib

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _threefry(key, c0, c1):
  """Vectorised threefry2x32: key (2,) uint32, counters uint32 arrays of equal shape."""
  k0, k1 = np.uint32(key[0]), np.uint32(key[1])
  ks = (k0, k1, np.uint32(k0 ^ k1 ^ np.uint32(0x1BD11BDA)))
  with np.errstate(over='ignore'):
    x0 = c0.astype(np.uint32) + ks[0]
    x1 = c1.astype(np.uint32) + ks[1]
    for i in range(5):
      for r in _ROT[i % 2]:
        x0 = x0 + x1
        x1 = (x1 << np.uint32(r)) | (x1 >> np.uint32(32 - r))
        x1 = x1 ^ x0
      x0 = x0 + ks[(i + 1) % 3]
      x1 = x1 + ks[(i + 2) % 3] + np.uint32(i + 1)
  return x0, x1


def random_bits(key, n: int) -> np.ndarray:
  counts = np.arange(n + (n & 1), dtype=np.uint32)
  if n & 1:
    counts[-1] = 0
  half = counts.size // 2
  a, b = _threefry(key, counts[:half], counts[half:])
  return np.concatenate([a, b])[:n]


def prng_key(seed: int) -> np.ndarray:
  """jax.random.PRNGKey(seed) for a non-negative 32-bit seed."""
  return np.array([0, int(seed) & 0xFFFFFFFF], dtype=np.uint32)


def split(key, num: int = 2) -> np.ndarray:
  """jax.random.split: uint32 [num, 2]."""
  return random_bits(np.asarray(key, dtype=np.uint32), 2 * num).reshape(num, 2)


def uniform(key, shape: Sequence[int]) -> np.ndarray:
  """jax.random.uniform(key, shape) on the host (small draws, checks)."""
  n = int(np.prod(shape)) if len(shape) else 1
  bits = random_bits(np.asarray(key, dtype=np.uint32), n)
  return (((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)).reshape(shape)


class DeviceUniform:
  """Up to four `jax.random.uniform` blocks per launch, keys fed through a small device buffer."""

  def __init__(self, counts: Sequence[int], device):
    if not 1 <= len(counts) <= 4:
      raise ValueError('1..4 blocks')
    self._counts = (C.c_int64 * len(counts))(*[int(c) for c in counts])
    self._n = len(counts)
    self.total = int(sum(counts))
    self._host = torch.zeros(2 * self._n, dtype=torch.int32).pin_memory()
    self.keys = torch.zeros(2 * self._n, dtype=torch.int32, device=device)
    self._done = torch.cuda.Event()
    self._pending = False

  def set_keys(self, keys) -> None:
    """keys: uint32 [nblocks, 2]; async H2D (call outside graph capture)."""
    if self._pending:
      self._done.synchronize()
    self._host.numpy().view(np.uint32)[:] = np.asarray(keys, dtype=np.uint32).reshape(-1)
    self.keys.copy_(self._host, non_blocking=True)
    self._done.record()
    self._pending = True

  def launch(self, out: torch.Tensor) -> None:
    """Enqueue the draws into `out` (float32, >= total elements) on the current stream."""
    if out.dtype != torch.float32 or out.numel() < self.total:
      raise ValueError('output tensor too small')
    _lib.call('dz_jax_uniform', self.keys.data_ptr(), C.cast(self._counts, C.c_void_p), self._n, out.data_ptr(),
              torch.cuda.current_stream().cuda_stream)


def iqn_update_keys(rng_key) -> Tuple[np.ndarray, np.ndarray]:
  """iqn/agent.py:207 + 182: `rng_key, update_key = split(rng_key)`; `_, *sample_keys = split(update_key, 4)`.
  Returns (new agent key [2], sample keys [3, 2])."""
  both = split(rng_key, 2)
  return both[0].copy(), split(both[1], 4)[1:].copy()


def iqn_act_keys(rng_key) -> Tuple[np.ndarray, np.ndarray]:
  """iqn/agent.py:220: `rng_key, sample_key, apply_key, policy_key = split(rng_key, 4)`.
  Returns (new agent key [2], sample key [1, 2])."""
  four = split(rng_key, 4)
  return four[0].copy(), four[1:2].copy()
