"""Batched acting throughput: E environment streams per tick through BatchedEpsilonGreedyActor (one network evaluation and
ONE device-to-host copy of E actions per tick) versus the single-observation path (one D2H sync per decision).
  python tools/bench_acting.py --agent dqn --streams 32 [--ticks 300]
Observations are device-resident frame stacks (what processors.BatchedAtariPreprocessor hands out)."""

import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--agent', default='dqn')
  ap.add_argument('--streams', type=int, default=32)
  ap.add_argument('--ticks', type=int, default=300)
  a = ap.parse_args()
  from dqn_zoo_b200 import agent as agent_lib
  from dqn_zoo_b200 import learner as dl
  from oracle import learner_oracle as lo
  L = dl.Learner(dl.NetworkSpec(a.agent, 6), batch_size=max(32, a.streams))
  L.set_params(lo.init_params(lo.NetSpec(a.agent, 6), 2), also_target=True)
  E = a.streams
  obs = torch.randint(0, 256, (E, 84, 84, 4), dtype=torch.uint8, device='cuda')
  actor = agent_lib.BatchedEpsilonGreedyActor(L, E, exploration_epsilon=0.01, rng_key=[0, 3])
  for _ in range(20):
    actor.step(obs)
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(a.ticks):
    actor.step(obs)
  torch.cuda.synchronize()
  batched = E * a.ticks / (time.perf_counter() - t0)
  # single-observation path: q_values + D2H per decision
  noise = taus = None
  if a.agent == 'rainbow':
    noise = L.noise
  if a.agent == 'iqn':
    taus = L.taus[:64]
  for _ in range(20):
    L.q_values(obs[0], taus=taus, noise=noise).cpu()
  t0 = time.perf_counter()
  n = max(100, a.ticks)
  for i in range(n):
    L.q_values(obs[i % E], taus=taus, noise=noise).cpu()
  single = n / (time.perf_counter() - t0)
  print(json.dumps({'metric': 'acting decisions per second (network evaluation + action choice, device-resident observations)',
                    'agent': a.agent, 'streams': E, 'batched_decisions_per_s': batched, 'ms_per_tick': 1e3 * E / batched,
                    'single_observation_decisions_per_s': single, 'speedup': batched / single,
                    'd2h_bytes_per_tick': 4 * E, 'h2d_bytes_per_tick': 8 * E}))


if __name__ == '__main__':
  main()
