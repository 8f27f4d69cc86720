"""CPU: the oracle restatement vs (1) golden vectors made from the reference's replay.py,
(2) the reference's known-answer tables, (3) the reference itself when it is mounted."""

import numpy as np
import pytest

from oracle import ref_import, replay_oracle, scenarios
import replay_contract as rc


@pytest.mark.parametrize('name', list(scenarios.ALL))
def test_oracle_reproduces_reference_golden(name):
  rc.check_scenario(replay_oracle, name, 'oracle')


@pytest.mark.parametrize('fn', rc.CONTRACT, ids=lambda f: f.__name__)
def test_oracle_contract(fn):
  fn(replay_oracle)


@pytest.mark.skipif(not ref_import.available(), reason='reference not mounted (GPU box)')
@pytest.mark.parametrize('fn', rc.CONTRACT, ids=lambda f: f.__name__)
def test_contract_holds_for_the_reference_itself(fn):
  """The contract file must describe the reference: run it on the reference's own classes."""
  fn(ref_import.load_reference_replay())


@pytest.mark.skipif(not ref_import.available(), reason='reference not mounted (GPU box)')
def test_oracle_matches_reference_on_long_random_per_run():
  ref = ref_import.load_reference_replay()
  res = {}
  for name, lib in (('ref', ref), ('oracle', replay_oracle)):
    res[name] = scenarios.prioritized_replay_script(lib, capacity=257, alpha=0.5, usp=1e-3, normalize=True,
                                                    batch=32, rounds=400, seed=31)
  for k in res['ref']:
    np.testing.assert_array_equal(res['ref'][k], res['oracle'][k], err_msg=k)


def test_synthetic_rows_are_deterministic_and_in_range():
  obs, a, r, d = replay_oracle.synthetic_rows(1, np.arange(100), 64, 6)
  obs2, a2, r2, d2 = replay_oracle.synthetic_rows(1, np.arange(100), 64, 6)
  np.testing.assert_array_equal(obs, obs2)
  assert obs.shape == (100, 2, 64) and obs.dtype == np.uint8
  assert a.min() >= 0 and a.max() < 6
  assert set(np.unique(r)).issubset({-1.0, 0.0, 1.0}) and set(np.unique(d)).issubset({0.0, 0.99})
  obs3, *_ = replay_oracle.synthetic_rows(2, np.arange(100), 64, 6)
  assert (obs3 != obs).mean() > 0.9
