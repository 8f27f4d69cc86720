"""Debug: the true timeline of one CUDA-graph learner step.  Every kernel stamps %globaltimer right after its
griddepcontrol.wait (dz_debug_timeline); names come from one eager profiled step (launch geometry -> name).
  python tools/step_timeline.py --agent rainbow [--steps 20]
Prints, per kernel in start order: start offset inside the step (us) and the gap to the next start, averaged over steps."""

import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--agent', default='rainbow')
  ap.add_argument('--steps', type=int, default=20)
  ap.add_argument('--capacity', type=int, default=131072)
  a = ap.parse_args()
  from dqn_zoo_b200 import _lib
  args = argparse.Namespace(agent=a.agent, capacity=a.capacity, batch=32, seed=1, no_graph=False)
  torch.cuda.set_device(0)
  ag, rep = bench.build_agent(args, 0, torch.device('cuda', 0))
  for _ in range(10):
    ag.learn()
  torch.cuda.synchronize()
  # names: one eager profiled step
  ag._use_graph = False
  ag.learn()
  _lib.call('dz_profile_begin')
  ag.learn()
  buf = C.create_string_buffer(1 << 16)
  _lib.call('dz_profile_end', buf, len(buf))
  prof = json.loads(buf.value.decode())
  names = {}
  for k, v in prof.items():
    names.setdefault((v[2], v[3], v[4]), []).append(k)
  ag._use_graph = True
  for _ in range(5):
    ag.learn()
  torch.cuda.synchronize()
  tl = torch.zeros(2 + 2 * 4000, dtype=torch.int64, device='cuda')
  _lib.call('dz_debug_timeline', tl.data_ptr())
  for _ in range(a.steps):
    ag.learn()
  torch.cuda.synchronize()
  _lib.call('dz_debug_timeline', 0)
  t = tl.cpu().numpy()
  n = int(t[0] & 0xffffffff)
  ts = t[2:2 + 2 * n:2].astype(np.int64)
  sig = t[3:3 + 2 * n:2].astype(np.uint64)
  order = np.argsort(ts, kind='stable')
  ts, sig = ts[order], sig[order]
  per = n // a.steps
  print('%d stamps, %d per step' % (n, per))
  if per * a.steps != n:
    print('stamp count is not a multiple of the step count; printing the raw sequence of the last step')
  ts = ts[-per * (a.steps - 2):].reshape(a.steps - 2, per)     # drop the first two steps
  sig = sig[-per * (a.steps - 2):].reshape(a.steps - 2, per)
  start = (ts - ts[:, :1]).mean(axis=0) / 1e3
  step_us = float(np.diff(ts[:, 0]).mean() / 1e3)
  print('step period %.1f us' % step_us)
  for i in range(per):
    s = int(sig[-1, i])
    key = (s >> 32, (s >> 16) & 0xffff, s & 0xffff)
    nxt = start[i + 1] if i + 1 < per else step_us
    print('%7.1f us  +%6.1f  %-40s grid %dx%d block %d' % (start[i], nxt - start[i], '/'.join(names.get(key, ['?'])), key[0], key[1], key[2]))


if __name__ == '__main__':
  main()
