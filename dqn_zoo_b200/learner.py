"""Host wrapper of the CUDA learner (`dz_learner_*` in include/dqn_zoo_b200.h).

Replaces the `jax.jit(update)` closure + `_learn()` glue of every dqn_zoo agent
(`dqn/agent.py:109-119,179-189`, `rainbow/agent.py:111-123,181-198`, ...).  The network and
optimizer are declarative (`kind`, hyper-parameters) instead of `hk.Transformed` /
`optax.GradientTransformation` objects, which cannot cross into CUDA; that is the one place the
agent constructor surface differs from the reference.

PyTorch tensors are used purely as device memory; all math runs in csrc/*.cu.
"""

from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Mapping, NamedTuple, Optional

import numpy as np
import torch

from dqn_zoo_b200 import _lib


class OptimizerSpec(NamedTuple):
  """optax stand-in: 'adam' (+ optional clip_by_global_norm) or centred 'rmsprop'."""
  name: str
  learning_rate: float
  eps: float
  decay: float = 0.95
  b1: float = 0.9
  b2: float = 0.999
  max_global_grad_norm: float = 0.0


def default_optimizer(kind: str) -> OptimizerSpec:
  """Per-agent optimizers from the reference's run_atari flags."""
  if kind in ('dqn', 'double_q'):
    return OptimizerSpec('rmsprop', 0.00025, 0.01 / 32 ** 2)             # dqn/run_atari.py:205-210
  if kind == 'prioritized':
    return OptimizerSpec('rmsprop', 0.00025 / 4, 0.01 / 32 ** 2 / 16)    # prioritized/run_atari.py:92-99
  if kind == 'c51':
    return OptimizerSpec('adam', 0.00025, 0.01 / 32, max_global_grad_norm=10.0)
  if kind == 'qrdqn':
    return OptimizerSpec('adam', 0.00005, 0.01 / 32, max_global_grad_norm=10.0)
  if kind == 'rainbow':
    return OptimizerSpec('adam', 0.0000625, 0.005 / 32, max_global_grad_norm=10.0)  # rainbow/run_atari.py:229-235
  if kind == 'iqn':
    return OptimizerSpec('adam', 0.00005, 0.01 / 32)
  raise ValueError(kind)


class NetworkSpec(NamedTuple):
  """Declarative stand-in for `networks.<kind>_atari_network(...)` (networks.py:224-363)."""
  kind: str
  num_actions: int
  num_atoms: int = 51
  vmax: float = 10.0
  num_quantiles: int = 201
  latent_dim: int = 64
  noisy_weight_init: float = 0.1
  tau_samples_s_tm1: int = 64
  tau_samples_policy: int = 64
  tau_samples_s_t: int = 64
  obs_shape: tuple = (84, 84, 4)


# canonical name -> haiku-style module path (the nested "sequential/..." prefixes are
# [UNVERIFIED-3P]; leaf names conv2_d / linear / mu / sigma / w / b are pinned by networks_test.py:41-53,153-163)
def haiku_name(canonical: str, kind: str):
  parts = canonical.split('/')
  leaf = parts[-1]
  conv = {'conv1': 'conv2_d', 'conv2': 'conv2_d_1', 'conv3': 'conv2_d_2'}
  if parts[0] in conv:
    return 'sequential/sequential/' + conv[parts[0]], leaf
  if kind == 'rainbow':
    idx = {'adv1': '', 'adv2': '_1', 'val1': '_2', 'val2': '_3'}[parts[0]]
    return 'noisy_linear%s/%s' % (idx, parts[1]), leaf
  if kind == 'iqn':
    return {'embed': 'batch_apply/linear', 'fc1': 'batch_apply_1/sequential/linear',
            'head': 'batch_apply_1/sequential/linear_1'}[parts[0]], leaf
  return {'fc1': 'sequential/sequential_1/linear', 'head': 'sequential/sequential_1/linear_1'}[parts[0]], leaf


def noise_vector_sizes(net: NetworkSpec):
  """(name, length) of the 8 factorised-noise vectors of ONE `network.apply`, in
  `hk.next_rng_key()` order (networks.py:169-170, :235-248)."""
  h = net.obs_shape[0]
  for k, s in ((8, 4), (4, 2), (3, 1)):
    h = (h - k) // s + 1
  w = net.obs_shape[1]
  for k, s in ((8, 4), (4, 2), (3, 1)):
    w = (w - k) // s + 1
  d = h * w * 64
  a, k = net.num_actions, net.num_atoms
  return [('adv1/in', d), ('adv1/out', 512), ('adv2/in', 512), ('adv2/out', a * k),
          ('val1/in', d), ('val1/out', 512), ('val2/in', 512), ('val2/out', k)]


def pack_noise(net: NetworkSpec, applies) -> np.ndarray:
  """Flattens a list of per-apply {name: vector} dicts into the device layout (each vector padded
  to a multiple of 4 floats so the kernels can use 16-byte loads)."""
  out = []
  for one in applies:
    for name, n in noise_vector_sizes(net):
      v = np.asarray(one[name], dtype=np.float32).reshape(-1)
      assert v.size == n, (name, v.size, n)
      out.append(np.concatenate([v, np.zeros((-n) % 4, dtype=np.float32)]))
  return np.concatenate(out)


def _cstream():
  return torch.cuda.current_stream().cuda_stream


class Learner:
  """Device-resident parameters, optimizer state and workspace + the fused update."""

  def __init__(self, net: NetworkSpec, batch_size: int = 32, optimizer: Optional[OptimizerSpec] = None,
               grad_error_bound: float = 1.0 / 32, huber_param: float = 1.0, device=None):
    if not torch.cuda.is_available():
      raise RuntimeError('dqn_zoo_b200.learner needs a CUDA device (there is no CPU fallback)')
    self.net = net
    self.kind = net.kind
    self.batch_size = batch_size
    self.opt = optimizer or default_optimizer(net.kind)
    self.device = torch.device(device or ('cuda:%d' % torch.cuda.current_device()))
    cfg = _lib.LearnerConfig()
    cfg.kind = _lib.AGENT_KINDS[net.kind]
    cfg.num_actions, cfg.num_atoms, cfg.num_quantiles, cfg.latent_dim = net.num_actions, net.num_atoms, net.num_quantiles, net.latent_dim
    cfg.tau_samples_s_tm1, cfg.tau_samples_policy, cfg.tau_samples_s_t = net.tau_samples_s_tm1, net.tau_samples_policy, net.tau_samples_s_t
    cfg.batch = batch_size
    cfg.obs_h, cfg.obs_w, cfg.obs_c = net.obs_shape
    cfg.vmax, cfg.grad_error_bound, cfg.huber_param = net.vmax, grad_error_bound, huber_param
    cfg.optimizer = _lib.OPTIMIZERS[self.opt.name]
    cfg.learning_rate, cfg.opt_eps, cfg.rms_decay = self.opt.learning_rate, self.opt.eps, self.opt.decay
    cfg.adam_b1, cfg.adam_b2, cfg.max_global_grad_norm = self.opt.b1, self.opt.b2, self.opt.max_global_grad_norm
    self.cfg = cfg
    plan = _lib.LearnerPlan()
    _lib.call('dz_learner_plan_query', C.byref(cfg), C.byref(plan))
    self.plan = plan
    self.obs_bytes = int(np.prod(net.obs_shape))
    dev = self.device
    P = plan.param_count
    self.online = torch.zeros(P, dtype=torch.float32, device=dev)
    self.target = torch.zeros(P, dtype=torch.float32, device=dev)
    self.grads = torch.zeros(P, dtype=torch.float32, device=dev)
    self.opt_state = torch.zeros(plan.opt_state_floats, dtype=torch.float32, device=dev)
    self.workspace = torch.zeros(plan.workspace_bytes, dtype=torch.uint8, device=dev)
    self.counters = torch.zeros(4, dtype=torch.int64, device=dev)
    self.taus = torch.zeros(max(plan.tau_floats, 1), dtype=torch.float32, device=dev)
    self.noise = torch.zeros(max(plan.noise_floats, 1), dtype=torch.float32, device=dev)
    self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
    self.per_example = torch.zeros(batch_size, dtype=torch.float32, device=dev)
    self.priorities = torch.zeros(batch_size, dtype=torch.float32, device=dev)
    self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)
    self.max_seen_priority = torch.ones(1, dtype=torch.float32, device=dev)   # rainbow/agent.py:79
    self.q_out = torch.zeros(64, dtype=torch.float32, device=dev)
    # tensor table
    self.tensors = {}
    name = C.create_string_buffer(64)
    shape = (C.c_int64 * 4)()
    ndim, off = C.c_int32(), C.c_int64()
    for i in range(plan.num_tensors):
      _lib.call('dz_learner_tensor_info', C.byref(cfg), i, name, shape, C.byref(ndim), C.byref(off))
      self.tensors[name.value.decode()] = (off.value, tuple(shape[k] for k in range(ndim.value)))
    bufs = _lib.LearnerBuffers(self.online.data_ptr(), self.target.data_ptr(), self.grads.data_ptr(),
                               self.opt_state.data_ptr(), self.workspace.data_ptr(), self.counters.data_ptr())
    handle = C.c_void_p()
    _lib.call('dz_learner_create', C.byref(cfg), C.byref(bufs), C.byref(handle))
    self._h = handle
    self._graph = None
    self._learn_io = None

  def __del__(self):
    h, self._h = getattr(self, '_h', None), None
    if h:
      _lib.lib.dz_learner_destroy(h)

  # -- parameters ----------------------------------------------------------------------------------
  def view(self, blob: torch.Tensor, name: str) -> torch.Tensor:
    off, shape = self.tensors[name]
    return blob[off:off + int(np.prod(shape))].view(shape)

  def init_params(self, seed: int) -> None:
    """Legacy U(+-1/sqrt(fan_in)) init for weights AND biases (networks.py:58-79); noisy sigma =
    sigma0/sqrt(in) (networks.py:156-166).  numpy RandomState stream — not the JAX PRNG."""
    rs = np.random.RandomState(seed)
    params = {}
    for name, (_, shape) in self.tensors.items():
      layer = name.rsplit('/', 1)[0]
      n_in = int(np.prod(self.tensors[layer + '/w'][1][:-1]))
      if '/sigma/' in name:
        params[name] = np.full(shape, self.net.noisy_weight_init / math.sqrt(n_in), dtype=np.float32)
      else:
        bound = math.sqrt(1.0 / n_in)
        params[name] = rs.uniform(-bound, bound, size=shape).astype(np.float32)
    self.set_params(params, also_target=True)

  def set_params(self, params: Mapping[str, np.ndarray], also_target: bool = False, blob: str = 'online') -> None:
    dst = getattr(self, blob)
    for name, value in params.items():
      self.view(dst, name).copy_(torch.as_tensor(np.asarray(value, dtype=np.float32)))
    if also_target:
      self.sync_target()

  def get_params(self, blob: str = 'online') -> Dict[str, np.ndarray]:
    src = getattr(self, blob)
    return {name: self.view(src, name).cpu().numpy() for name in self.tensors}

  def haiku_params(self, blob: str = 'online'):
    """Nested {module: {leaf: array}} like `hk.Params` (the reference's `online_params`)."""
    out = {}
    for name, value in self.get_params(blob).items():
      mod, leaf = haiku_name(name, self.kind)
      out.setdefault(mod, {})[leaf] = value
    return out

  def get_opt_state(self):
    """optax-shaped: adam -> {'count','mu','nu'}; rmsprop -> {'mu','nu'} (dicts by canonical name)."""
    P = self.plan.param_count
    mu, nu = self.opt_state[:P], self.opt_state[P:]
    st = {'mu': {n: self.view(mu, n).cpu().numpy() for n in self.tensors},
          'nu': {n: self.view(nu, n).cpu().numpy() for n in self.tensors}}
    st['count'] = int(self.counters[0].item())
    return st

  def set_opt_state(self, st) -> None:
    P = self.plan.param_count
    mu, nu = self.opt_state[:P], self.opt_state[P:]
    for n in self.tensors:
      self.view(mu, n).copy_(torch.as_tensor(np.asarray(st['mu'][n], dtype=np.float32)))
      self.view(nu, n).copy_(torch.as_tensor(np.asarray(st['nu'][n], dtype=np.float32)))
    self.counters[0] = int(st.get('count', 0))

  def sync_target(self) -> None:
    """`self._target_params = self._online_params` (dqn/agent.py:155-156)."""
    _lib.call('dz_learner_sync_target', self._h, _cstream())

  # -- updates -------------------------------------------------------------------------------------
  def _row_table(self, dense: torch.Tensor) -> torch.Tensor:
    n = dense.shape[0]
    stride = dense.stride(0) * dense.element_size()
    return dense.data_ptr() + torch.arange(n, dtype=torch.int64, device=self.device) * stride

  def update(self, s_tm1, a_tm1, r_t, discount_t, s_t, weights=None, taus=None, noise=None, apply_update=True):
    """`jit(update)` on an explicit batch of device tensors (uint8 [B,H,W,C], int, float, float,
    uint8).  r/discount/weights are rounded to float32 as at the jit boundary.  Returns nothing;
    results are in `.loss`, `.per_example`, `.priorities`, `.grad_norm`, `.grads`."""
    dev = self.device
    B = self.batch_size
    s_tm1 = torch.as_tensor(s_tm1, device=dev).contiguous().view(B, -1)
    s_t = torch.as_tensor(s_t, device=dev).contiguous().view(B, -1)
    assert s_tm1.dtype == torch.uint8 and s_tm1.shape[1] == self.obs_bytes
    keep = [s_tm1, s_t, self._row_table(s_tm1), self._row_table(s_t),
            torch.as_tensor(a_tm1, device=dev).to(torch.int32).contiguous(),
            torch.as_tensor(r_t, device=dev).to(torch.float32).contiguous(),
            torch.as_tensor(discount_t, device=dev).to(torch.float32).contiguous()]
    w = None if weights is None else torch.as_tensor(weights, device=dev).to(torch.float32).contiguous()
    if taus is not None:
      flat_t = torch.as_tensor(taus, device=dev).to(torch.float32).reshape(-1)
      self.taus[:flat_t.numel()].copy_(flat_t)
    if noise is not None:
      flat = torch.as_tensor(noise, device=dev).to(torch.float32).reshape(-1)
      self.noise[:flat.numel()].copy_(flat)
    batch = _lib.Batch(keep[2].data_ptr(), keep[3].data_ptr(), keep[4].data_ptr(), keep[5].data_ptr(),
                       keep[6].data_ptr(), 0 if w is None else w.data_ptr(),
                       self.taus.data_ptr() if self.kind == 'iqn' else 0,
                       self.noise.data_ptr() if self.kind == 'rainbow' else 0)
    out = _lib.UpdateOutputs(self.loss.data_ptr(), self.per_example.data_ptr(), self.priorities.data_ptr(),
                             self.grad_norm.data_ptr())
    _lib.call('dz_learner_update', self._h, C.byref(batch), C.byref(out), 1 if apply_update else 0, _cstream())
    self._keep = (keep, w)

  def generate_randomness(self, seed: int, beside_sampler: bool = False) -> None:
    """Fills `.taus` / `.noise` for the next update from the device generator (Philox).  `beside_sampler`: enqueue on
    the learner's side stream (ordered before the next learn()/update()/q_values() only)."""
    _lib.call('dz_learner_generate_randomness_async' if beside_sampler else 'dz_learner_generate_randomness', self._h, seed,
              self.taus.data_ptr(), self.noise.data_ptr(), _cstream())

  def q_values(self, obs_u8: torch.Tensor, taus=None, noise=None) -> torch.Tensor:
    """Online-network Q-values for one observation (the network half of select_action)."""
    obs = torch.as_tensor(obs_u8, device=self.device).contiguous().view(-1)
    t = None if taus is None else torch.as_tensor(taus, device=self.device).to(torch.float32).contiguous()
    n = None if noise is None else torch.as_tensor(noise, device=self.device).to(torch.float32).contiguous()
    _lib.call('dz_learner_q_values', self._h, obs.data_ptr(), 0 if t is None else t.data_ptr(),
              0 if n is None else n.data_ptr(), self.q_out.data_ptr(), _cstream())
    self._keep_q = (obs, t, n)
    return self.q_out[:self.net.num_actions]

  def act_batch(self, obs_u8: torch.Tensor, epsilon: float = 0.0, explore=None, taus=None, noise=None):
    """Batched select_action for E <= batch_size environment streams in ONE enqueue: `obs_u8` is [E, H, W, C] uint8 (device
    or host), `explore` a float32 [2, E] tensor of uniforms in [0, 1) (None: greedy).  Returns (actions int32 [E],
    q_values float32 [E, num_actions]) as device tensors — the caller does one D2H of the actions per tick."""
    obs = torch.as_tensor(obs_u8, device=self.device).contiguous()
    E = int(obs.shape[0])
    if not hasattr(self, '_act_q') or self._act_q.shape[0] < E:
      self._act_q = torch.zeros((self.batch_size, self.net.num_actions), dtype=torch.float32, device=self.device)
      self._act_a = torch.zeros(self.batch_size, dtype=torch.int32, device=self.device)
    t = None if taus is None else torch.as_tensor(taus, device=self.device).to(torch.float32).contiguous()
    n = None if noise is None else torch.as_tensor(noise, device=self.device).to(torch.float32).contiguous()
    x = None if explore is None else torch.as_tensor(explore, device=self.device).to(torch.float32).contiguous()
    _lib.call('dz_learner_act_batch', self._h, obs.data_ptr(), E, 0 if t is None else t.data_ptr(),
              0 if n is None else n.data_ptr(), 0 if x is None else x.data_ptr(), float(epsilon), self._act_q.data_ptr(),
              self._act_a.data_ptr(), _cstream())
    self._keep_act = (obs, t, n, x)
    return self._act_a[:E], self._act_q[:E]

  # -- fused sample -> update -> priority write-back -------------------------------------------------
  def make_learn_io(self, stage: torch.Tensor, prioritized: bool, priority_exponent: float):
    """Binds the per-step staging buffer (float64 view: [pos(int64) B | u_tree B | u_mix B | scalars 4])
    and persistent sample outputs into a dz_learn_io."""
    B = self.batch_size
    dev = self.device
    self.s_ids = torch.zeros(3 * B, dtype=torch.int64, device=dev)
    self.s_f64 = torch.zeros(2 * B, dtype=torch.float64, device=dev)
    io = _lib.LearnIO()
    base = stage.data_ptr()
    io.sample_in = _lib.SampleInputs(base, base + 8 * B, base + 16 * B, base + 24 * B)
    sp, fp = self.s_ids.data_ptr(), self.s_f64.data_ptr()
    io.sample_out = _lib.SampleOutputs(sp, sp + 8 * B, sp + 16 * B, fp, fp + 8 * B)
    io.d_taus = self.taus.data_ptr() if self.kind == 'iqn' else 0
    io.d_noise = self.noise.data_ptr() if self.kind == 'rainbow' else 0
    io.update_out = _lib.UpdateOutputs(self.loss.data_ptr(), self.per_example.data_ptr(), self.priorities.data_ptr(),
                                       self.grad_norm.data_ptr())
    io.d_max_seen_priority = self.max_seen_priority.data_ptr()
    io.priority_exponent = float(priority_exponent)
    return io

  def learn(self, replay_view, prioritized: bool, io) -> None:
    """One `_learn()` enqueue (rainbow/agent.py:181-198)."""
    _lib.call('dz_learner_learn', self._h, C.byref(replay_view), 1 if prioritized else 0, C.byref(io), _cstream())

  @property
  def sampled_ids(self):
    return self.s_ids[:self.batch_size]

  @property
  def sampled_indices(self):
    return self.s_ids[self.batch_size:2 * self.batch_size]

  @property
  def sampled_weights(self):
    return self.s_f64[self.batch_size:]
