"""Runs a few eager learner steps of one agent between cudaProfilerStart/Stop (for ncu
--profile-from-start off).  Not a benchmark: numbers printed under a profiler are never reported."""

import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--agent', default='rainbow')
  ap.add_argument('--capacity', type=int, default=131072)
  ap.add_argument('--steps', type=int, default=2)
  ap.add_argument('--warm', type=int, default=5)
  a = ap.parse_args()
  args = argparse.Namespace(agent=a.agent, capacity=a.capacity, batch=32, seed=1, no_graph=True)
  torch.cuda.set_device(0)
  ag, rep = bench.build_agent(args, 0, torch.device('cuda', 0))
  for _ in range(a.warm):
    ag.learn()
  torch.cuda.synchronize()
  torch.cuda.profiler.start()
  for _ in range(a.steps):
    ag.learn()
  torch.cuda.synchronize()
  torch.cuda.profiler.stop()


if __name__ == '__main__':
  main()
