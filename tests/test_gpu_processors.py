"""GPU: device Atari preprocessing (dqn_zoo_b200/processors.py + csrc/dz_preprocess.cu) against the reference's golden
vector (processors_test.py:405-475) and, step by step, against the CPU oracle."""

import hashlib

import numpy as np
import pytest
import torch

from oracle import processors_oracle as po
from test_oracle_processors import GOLDEN_INPUT_HASHES, GOLDEN_OUTPUT_HASH, golden_inputs

pytestmark = pytest.mark.gpu


def ts(step_type, reward, discount, rgb, lives=3):
  from dqn_zoo_b200 import parts
  return parts.TimeStep(step_type=parts.StepType(step_type), reward=reward, discount=discount, observation=(rgb, lives))


def test_reference_golden_vector_on_device():
  from dqn_zoo_b200 import processors
  rgb = golden_inputs()
  assert [hashlib.sha256(o).hexdigest() for o in rgb] == GOLDEN_INPUT_HASHES
  processor = processors.atari()
  steps = [(0, None, None), (1, 0.5, 0.9), (1, 0.2, 0.9), (1, 0, 0.9), (1, 0.1, 0.9)]
  processed = None
  for (st, r, d), o in zip(steps, rgb):
    processed = processor(ts(st, r, d, o))
  assert processed is not None
  assert processed.step_type == 1
  assert processed.reward == pytest.approx(0.5 + 0.2 + 0.0 + 0.1)
  assert processed.discount == pytest.approx(0.9 ** 4 * 0.99)
  assert processed.observation.dtype == np.uint8 and processed.observation.shape == (84, 84, 4)
  assert hashlib.sha256(processed.observation.flatten()).hexdigest() == GOLDEN_OUTPUT_HASH


def random_episode(rs, length, shape, life_loss_at=None):
  frames = [rs.randint(0, 256, size=shape, dtype=np.uint8) for _ in range(length)]
  out = []
  for i, f in enumerate(frames):
    st = 0 if i == 0 else (2 if i == length - 1 else 1)
    reward = None if st == 0 else float(rs.choice([-3.0, -1.0, 0.0, 0.5, 1.0, 2.0]))
    discount = None if st == 0 else (0.0 if st == 2 else 1.0)
    lives = 3 if life_loss_at is None or i < life_loss_at else 2
    out.append((st, reward, discount, f, lives))
  return out


def same(a, b):
  if a is None or b is None:
    return a is None and b is None
  ok = int(a[0]) == int(b.step_type)
  ok &= (a[1] is None and b.reward is None) or (a[1] is not None and b.reward is not None and a[1] == b.reward)
  ok &= (a[2] is None and b.discount is None) or (a[2] is not None and b.discount is not None and a[2] == b.discount)
  obs = b.observation.cpu().numpy() if torch.is_tensor(b.observation) else b.observation
  return ok and np.array_equal(a[3], obs)


@pytest.mark.parametrize('device_obs', [False, True])
def test_episodes_match_oracle_step_by_step(device_obs):
  from dqn_zoo_b200 import processors
  rs = np.random.RandomState(11)
  dev = processors.atari(device_observations=device_obs)
  ref = po.AtariPreprocessor()
  emitted = 0
  for length, loss_at in [(1 + 4 * 3, None), (7, 3), (2, None), (18, 9), (5, None)]:
    dev.reset()
    ref.reset()
    for st, r, d, f, lives in random_episode(rs, length, (210, 160, 3), loss_at):
      want = ref(st, r, d, (f, lives))
      got = dev(ts(st, r, d, f, lives))
      assert same(want, got), (length, st)
      emitted += want is not None
  assert emitted >= 12


def test_other_geometry_and_stack_depth():
  from dqn_zoo_b200 import processors
  rs = np.random.RandomState(12)
  kwargs = dict(resize_shape=(42, 50), num_action_repeats=3, num_stacked_frames=3, max_abs_reward=None, additional_discount=0.9)
  dev = processors.atari(**kwargs)
  ref = po.AtariPreprocessor(**kwargs)
  for length in (10, 4):
    dev.reset()
    ref.reset()
    for st, r, d, f, lives in random_episode(rs, length, (100, 96, 3)):
      assert same(ref(st, r, d, (f, lives)), dev(ts(st, r, d, f, lives)))


def test_many_streams_in_one_launch():
  from dqn_zoo_b200 import processors
  rs = np.random.RandomState(13)
  n = 5
  dev = processors.BatchedAtariPreprocessor(num_streams=n)
  refs = [po.AtariPreprocessor() for _ in range(n)]
  episodes = [random_episode(rs, 9 + 2 * e, (210, 160, 3), life_loss_at=(4 if e % 2 else None)) for e in range(n)]
  for t in range(max(len(ep) for ep in episodes)):
    batch, wants = [], []
    for e in range(n):
      if t < len(episodes[e]):
        st, r, d, f, lives = episodes[e][t]
        batch.append(ts(st, r, d, f, lives))
        wants.append(refs[e](st, r, d, (f, lives)))
      else:
        batch.append(None)
        wants.append(None)
    gots = dev.step(batch)
    for e in range(n):
      assert same(wants[e], gots[e]), (t, e)
  assert dev.stacks.shape == (n, 84, 84, 4) and dev.stacks.is_cuda


def test_vectorized_device_frame_path_equals_the_per_stream_path():
  """VectorizedAtariPreprocessor.step_arrays (device-resident frames, array state machine, one launch per tick) against
  BatchedAtariPreprocessor.step (host frames, per-stream objects) on the same desynchronised episodes: identical
  emissions and bit-identical frame stacks after every tick."""
  from dqn_zoo_b200 import processors
  rs = np.random.RandomState(21)
  n = 6
  a = processors.BatchedAtariPreprocessor(num_streams=n, device_observations=True)
  b = processors.VectorizedAtariPreprocessor(num_streams=n, device_observations=True)
  episodes = [random_episode(rs, 11 + 3 * e, (210, 160, 3), life_loss_at=(5 if e % 2 else None)) for e in range(n)]
  blank = np.zeros((210, 160, 3), np.uint8)
  for t in range(max(len(ep) for ep in episodes)):
    batch, act = [], np.zeros(n, bool)
    frames = np.zeros((n, 210, 160, 3), np.uint8)
    st_a = np.ones(n, np.int64); rw = np.zeros(n); dc = np.ones(n); lv = np.zeros(n, np.int64)
    for e in range(n):
      if t < len(episodes[e]):
        st, r, d, f, lives = episodes[e][t]
        batch.append(ts(st, r, d, f, lives))
        act[e] = True; frames[e] = f; st_a[e] = int(st); lv[e] = lives
        rw[e] = np.nan if r is None else r; dc[e] = np.nan if d is None else d
      else:
        batch.append(None); frames[e] = blank
    outs = a.step(batch)
    got = b.step_arrays(torch.as_tensor(frames, device='cuda'), st_a, rw, dc, lv, act)
    torch.cuda.synchronize()
    for e in range(n):
      assert (outs[e] is not None) == bool(got['emit'][e]), (t, e)
      if outs[e] is not None:
        assert int(outs[e].step_type) == int(got['step_type'][e])
        assert (outs[e].reward is None and np.isnan(got['reward'][e])) or outs[e].reward == got['reward'][e]
        assert (outs[e].discount is None and np.isnan(got['discount'][e])) or outs[e].discount == got['discount'][e]
    assert torch.equal(a.stacks, b.stacks), t


def test_saturated_and_black_frames():
  from dqn_zoo_b200 import processors
  for value in (0, 255):
    dev = processors.atari()
    ref = po.AtariPreprocessor()
    f = np.full((210, 160, 3), value, dtype=np.uint8)
    assert same(ref(0, None, None, (f, 3)), dev(ts(0, None, None, f)))


def test_agent_with_device_resident_frames_equals_host_path():
  """Raw RGB frames -> device preprocessing -> acting -> replay insert, once with host observations (the reference's
  data flow) and once with the frame stacks staying in HBM (device_observations=True): same actions, same replay."""
  from dqn_zoo_b200 import agent as ag
  from dqn_zoo_b200 import learner as dl
  from dqn_zoo_b200 import processors
  from dqn_zoo_b200 import replay as dr
  rs = np.random.RandomState(21)
  episodes = [random_episode(rs, n, (210, 160, 3), life_loss_at=loss) for n, loss in [(23, 9), (14, None), (31, 17)]]

  def run(device_obs):
    rep = dr.PrioritizedTransitionReplay(16, dr.Transition(None, None, None, None, None), 0.5, lambda t: 0.5, 1e-3, True,
                                         np.random.RandomState(3))
    agent = ag.Rainbow(preprocessor=processors.atari(device_observations=device_obs),
                       sample_network_input=np.zeros((84, 84, 4), np.uint8), network=dl.NetworkSpec('rainbow', 6),
                       support=np.linspace(-10, 10, 51), optimizer=None,
                       transition_accumulator=dr.NStepTransitionAccumulator(3), replay=rep, batch_size=4,
                       min_replay_capacity_fraction=2.0, learn_period=4, target_network_update_period=16, rng_key=[0, 7],
                       use_cuda_graph=False)
    actions = []
    for ep in episodes:
      agent.reset()
      for st, r, d, f, lives in ep:
        actions.append(agent.step(ts(st, r, d, f, lives)))
    torch.cuda.synchronize()
    return actions, rep.get_state()

  a_host, s_host = run(False)
  a_dev, s_dev = run(True)
  assert a_host == a_dev
  assert len(s_host['storage']) == len(s_dev['storage']) > 5
  for (i0, t0), (i1, t1) in zip(s_host['storage'], s_dev['storage']):
    assert i0 == i1
    for x, y in zip(t0, t1):
      np.testing.assert_array_equal(np.asarray(x), np.asarray(y))


def test_luma_of_all_16777216_colours_is_bit_exact():
  """Every RGB colour once, identity-sized resample (Pillow's weights are then exactly (1, 0)): the kernel's
  fixed-point screen + float64 fallback must reproduce the oracle's rgb2y (the golden vector's rounding order)."""
  import ctypes as C
  from dqn_zoo_b200 import _lib, processors
  dev = torch.device('cuda')
  H, W = 65536, 256
  rows = torch.arange(H, device=dev)
  frame = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
  frame[..., 0] = (rows >> 8).to(torch.uint8)[:, None]
  frame[..., 1] = (rows & 255).to(torch.uint8)[:, None]
  frame[..., 2] = torch.arange(W, device=dev).to(torch.uint8)[None, :]
  ax_h, ax_v = processors._Axis(W, W, dev), processors._Axis(H, H, dev)
  band = int(_lib.lib.dz_atari_preprocess_band_rows())
  bv = ax_v.bounds_host
  max_rows = max(int(bv[min(y0 + band, H) - 1].sum() - bv[y0, 0]) for y0 in range(0, H, band))
  out = torch.zeros((H, W, 1), dtype=torch.uint8, device=dev)
  a = torch.tensor([frame.data_ptr()], dtype=torch.int64, device=dev)
  b = torch.zeros(1, dtype=torch.int64, device=dev)
  s = torch.tensor([out.data_ptr()], dtype=torch.int64, device=dev)
  counts = torch.zeros(1, dtype=torch.int32, device=dev)
  luma = (C.c_double * 3)(*processors.LUMA)
  _lib.call('dz_atari_preprocess', a.data_ptr(), b.data_ptr(), 1, C.byref(ax_h.c), C.byref(ax_v.c), s.data_ptr(),
            counts.data_ptr(), 1, C.cast(luma, C.c_void_p), max_rows, torch.cuda.current_stream().cuda_stream)
  torch.cuda.synchronize()
  got = out[..., 0].cpu().numpy()
  host = frame.cpu().numpy()
  for lo in range(0, H, 8192):
    want = po.rgb2y(host[lo:lo + 8192])
    assert np.array_equal(got[lo:lo + 8192], want), lo
