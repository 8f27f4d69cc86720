"""CPU restatement of the image half of the reference's Atari preprocessing.  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and the CPU leg of tools/bench_preprocess.py — never by the
product path (dqn_zoo_b200/processors.py runs the CUDA kernel and fails loudly without it).

Follows /root/reference/dqn_zoo/processors.py:
  * max-pool over the last `num_pooled_frames` raw frames ............ processors.py:485-487
  * rgb2y: tensordot with [0.299, 0.587, 1 - (0.299 + 0.587)] in float64, then astype(uint8) .. :367-371
  * resize: PIL `Image.resize((w, h), BILINEAR)` on the uint8 image ...... :374-388
  * frame stack: deque(maxlen=4) -> trailing zero pad -> stack on the last axis .. :492-500
  * scalar half (reward sum + clip, discount product * 0.99, step-type reduction, action-repeat
    cadence) ........................................................... :54-66, 121-215, 290-365, 399-505

Pillow's resize is a third-party dependency that is not under /root/reference (docker_requirements pins
Pillow==9.0.1); its algorithm (libImaging/Resample.c, unchanged since 3.x for 8-bit images) is restated in
`resample_coeffs` / `resize_bilinear_u8`: two passes (horizontal, then vertical), per-output-pixel windows
[xmin, xmax) from center +- support with support = max(scale, 1), triangle weights normalised in double,
converted to fixed point with PRECISION_BITS = 32 - 8 - 2 = 22, accumulated in int32 starting from
1 << 21 and shifted down, clipped to [0, 255]; the intermediate image between the passes is uint8.

PINNED: tests/test_oracle_processors.py checks (a) `resize_bilinear_u8` against the Pillow installed in the
build container (12.2.0) on random images, (b) that `rgb2y` and this container's np.tensordot differ only on
colours whose exact luma is within 1e-9 of an integer (rounding-order cases, see rgb2y), and
(c) the whole pipeline against the reference's own golden vector: the SHA-256 of the processed observation in
processors_test.py:405-475 (`0d158a8f...00aa`, inputs regenerated from RandomState(1) and checked against the
five input hashes the reference test lists)."""

import collections
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2

# processors.py:370 — the third weight is computed, not 0.114
LUMA = (0.299, 0.587, 1 - (0.299 + 0.587))


def rgb2y(array):
  """uint8 [H, W, 3] -> uint8 [H, W] (processors.py:367-371: float64 tensordot, then astype(uint8) = truncation).

  The reference delegates the three-term dot product to numpy's BLAS.  The exact value is an integer plus
  O(1e-14) for about one colour in 1000, so the ORDER of the float64 roundings decides those pixels, and that
  order depends on the BLAS build: the golden vector of processors_test.py:405-475 is reproduced by the plain
  left-to-right evaluation fl(fl(fl(r*wr) + fl(g*wg)) + fl(b*wb)) with separately rounded products, whereas the
  numpy in this build container (2.3.5, OpenBLAS dgemv) evaluates fma(b, wb, fma(r, wr, fl(g*wg))), differs from
  it on 522 of the 2^24 colours and FAILS the reference's own golden test.  Canonical here (and in the CUDA kernel,
  with __dmul_rn/__dadd_rn so nothing is contracted): the order that reproduces the reference's golden vector."""
  a = np.asarray(array)
  r = np.multiply(a[..., 0].astype(np.float64), LUMA[0])
  g = np.multiply(a[..., 1].astype(np.float64), LUMA[1])
  b = np.multiply(a[..., 2].astype(np.float64), LUMA[2])
  return np.add(np.add(r, g), b).astype(np.uint8)


def _triangle(x):
  x = -x if x < 0.0 else x
  return 1.0 - x if x < 1.0 else 0.0


def resample_coeffs(in_size, out_size):
  """Pillow precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter.

  Returns (bounds int32 [out, 2] = (xmin, count), kk int32 [out, ksize], ksize)."""
  scale = in_size / out_size
  filterscale = max(scale, 1.0)
  support = 1.0 * filterscale
  ksize = int(math.ceil(support)) * 2 + 1
  bounds = np.zeros((out_size, 2), dtype=np.int32)
  kk = np.zeros((out_size, ksize), dtype=np.int32)
  ss = 1.0 / filterscale
  for xx in range(out_size):
    center = (xx + 0.5) * scale
    xmin = int(center - support + 0.5)
    xmin = max(xmin, 0)
    xmax = int(center + support + 0.5)
    xmax = min(xmax, in_size)
    count = xmax - xmin
    w = [_triangle((x + xmin - center + 0.5) * ss) for x in range(count)]
    ww = 0.0
    for v in w:
      ww += v
    if ww != 0.0:
      w = [v / ww for v in w]
    bounds[xx] = (xmin, count)
    for x, v in enumerate(w):
      kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
  return bounds, kk, ksize


def _resample_axis1(img, bounds, kk):
  """One pass along axis 1 of a uint8 [rows, in] image -> uint8 [rows, out]."""
  rows = img.shape[0]
  out = np.empty((rows, bounds.shape[0]), dtype=np.uint8)
  src = img.astype(np.int64)
  for xx in range(bounds.shape[0]):
    xmin, count = int(bounds[xx, 0]), int(bounds[xx, 1])
    acc = np.full(rows, 1 << (PRECISION_BITS - 1), dtype=np.int64)
    for x in range(count):
      acc += src[:, xmin + x] * int(kk[xx, x])
    out[:, xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
  return out


def resize_bilinear_u8(img, out_h, out_w):
  """PIL.Image.fromarray(img).resize((out_w, out_h), BILINEAR) for a 2-D uint8 image."""
  img = np.ascontiguousarray(img, dtype=np.uint8)
  in_h, in_w = img.shape
  tmp = img
  if in_w != out_w:
    bh, kh, _ = resample_coeffs(in_w, out_w)
    tmp = _resample_axis1(tmp, bh, kh)
  if in_h != out_h:
    bv, kv, _ = resample_coeffs(in_h, out_h)
    tmp = np.ascontiguousarray(_resample_axis1(np.ascontiguousarray(tmp.T), bv, kv).T)
  return tmp


def pooled_gray_resized(frame_prev, frame_last, out_h=84, out_w=84):
  """max over the two frames (None = zero padding), grayscale, resize."""
  a = np.zeros_like(frame_last) if frame_prev is None else frame_prev
  b = np.zeros_like(a) if frame_last is None else frame_last
  return resize_bilinear_u8(rgb2y(np.maximum(a, b)), out_h, out_w)


FIRST, MID, LAST = 0, 1, 2


class AtariPreprocessor:
  """The whole `processors.atari()` pipeline on plain tuples (step_type, reward, discount, (rgb, lives)).

  Returns None or (step_type, reward, discount, observation uint8 [out_h, out_w, num_stacked])."""

  def __init__(self, additional_discount=0.99, max_abs_reward=1.0, resize_shape=(84, 84), num_action_repeats=4,
               num_pooled_frames=2, zero_discount_on_life_loss=True, num_stacked_frames=4):
    assert num_pooled_frames == 2
    self._gamma = additional_discount
    self._clip = max_abs_reward
    self._shape = resize_shape
    self._repeats = num_action_repeats
    self._life_loss = zero_discount_on_life_loss
    self._stack = num_stacked_frames
    self.reset()

  def reset(self):
    self._lives = None
    self._index = (-1) % self._repeats          # FixedPaddedBuffer(length, initial_index=-1)
    self._buffer = [None] * self._repeats
    self._since_first = None
    self._should_reset = False
    self._frames = collections.deque(maxlen=self._stack)

  def __call__(self, step_type, reward, discount, observation):
    rgb, lives = observation
    if self._life_loss:                          # processors.py:254-260
      lost = step_type == MID and lives < self._lives
      self._lives = lives
      if lost:
        discount = 0.0
    if self._index >= self._repeats:             # processors.py:121-150
      self._index = 0
      self._buffer = [None] * self._repeats
    self._buffer[self._index] = (step_type, reward, discount, rgb)
    self._index += 1
    if not self._emit():
      return None
    example = next(v for v in self._buffer if v is not None)
    zero = (0, 0.0 if example[1] is not None else 0, 0.0 if example[2] is not None else 0, None)
    slots = [zero if v is None else v for v in self._buffer]
    out_type = MID                               # processors.py:267-289
    for v in slots:
      if v[0] == 0:
        out_type = FIRST
        break
      if v[0] == LAST:
        out_type = LAST
        break
    rewards = [v[1] for v in slots]
    if None in rewards:
      out_reward = None
    else:
      out_reward = sum(rewards)
      if self._clip:
        out_reward = max(min(out_reward, self._clip), -self._clip)
    discounts = [v[2] for v in slots]
    if None in discounts:
      out_discount = None
    else:
      out_discount = 1
      for d in discounts:
        out_discount *= d
      out_discount = self._gamma * out_discount
    shape = example[3].shape
    frames = [np.zeros(shape, np.uint8) if v[3] is None else v[3] for v in slots[-2:]]
    gray = rgb2y(np.maximum(frames[0], frames[1]))
    if self._shape:
      gray = resize_bilinear_u8(gray, *self._shape)
    self._frames.append(gray)
    stack = list(self._frames)
    stack = stack + [np.zeros_like(stack[0])] * (self._stack - len(stack))
    return out_type, out_reward, out_discount, np.stack(stack, axis=-1)

  def _emit(self):                               # processors.py:165-215
    if self._should_reset:
      raise RuntimeError('Should have reset.')
    main = MID
    for v in self._buffer:
      if v is None:
        continue
      if v[0] in (FIRST, LAST):
        if main in (FIRST, LAST):
          raise RuntimeError('Expected at most one FIRST or LAST.')
        main = v[0]
    if self._since_first is None and main != FIRST:
      raise RuntimeError('After reset first timestep should be FIRST.')
    if main == FIRST:
      self._since_first = 0
      return True
    if main == LAST:
      self._since_first = None
      self._should_reset = True
      return True
    self._since_first += 1
    return self._since_first % self._repeats == 0
