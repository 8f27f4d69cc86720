"""Writes tests/golden/learner_hand_vectors.json: vectors computed BY HAND (plain python arithmetic spelled out below)
from the published rlax 0.1.2 / optax 0.1.2 / dqn_zoo formulas.  Nothing here imports the
oracle: tests/test_oracle_learner.py uses these numbers as an independent pin of oracle/learner_oracle.py."""

import json
import math
import os

g = {'_about': 'hand-computed from the published formulas; see the derivation strings and make_learner_hand_vectors.py'}

g['categorical_l2_project'] = [
    {'z_q': [-1.0, 0.0, 1.0], 'r_t': 0.5, 'discount_t': 0.5, 'z_p': [0.0, 0.5, 1.0], 'probs': [0.2, 0.3, 0.5],
     'expected': [0.0, 0.35, 0.65],
     'derivation': 'support gap 1. atom 0.0 sits on z=0: 0.2. atom 0.5 is halfway between z=0 and z=1: 0.15 each. '
                   'atom 1.0 sits on z=1: 0.5. Sum: [0, 0.2+0.15, 0.15+0.5].'},
    {'z_q': [-1.0, 0.0, 1.0], 'r_t': 1.0, 'discount_t': 0.99, 'z_p': [0.01, 1.0, 1.99], 'probs': [0.25, 0.25, 0.5],
     'expected': [0.0, 0.2475, 0.7525],
     'derivation': 'targets clipped to [-1,1]: [0.01, 1, 1]. atom 0.01: 0.99*0.25 on z=0, 0.01*0.25 on z=1; the clipped '
                   'atoms put 0.25+0.5 on z=1.'},
    {'z_q': [-1.0, 0.0, 1.0], 'r_t': -1.0, 'discount_t': 0.0, 'z_p': [-1.0, -1.0, -1.0], 'probs': [0.1, 0.6, 0.3],
     'expected': [1.0, 0.0, 0.0],
     'derivation': 'terminal transition: every atom collapses onto r=-1 = z_0 (delta = 0 takes the d_pos branch).'},
]
g['categorical_cross_entropy'] = {
    'logits_tm1': [0.0, math.log(2.0), math.log(3.0)], 'target': [0.0, 0.35, 0.65],
    'expected': -(0.35 * math.log(1.0 / 3.0) + 0.65 * math.log(0.5)),
    'derivation': 'softmax([0, ln2, ln3]) = [1/6, 1/3, 1/2]; loss = -(0.35 ln(1/3) + 0.65 ln(1/2))'}
g['quantile_regression_loss'] = [
    {'dist_src': [0.0, 1.0], 'tau': [0.25, 0.75], 'target': [0.5, 2.5], 'kappa': 1.0,
     'expected': (0.25 * 0.125 + 0.25 * 2.0) / 2 + (0.25 * 0.125 + 0.75 * 1.0) / 2,
     'derivation': 'delta[i][j] = target[j]-src[i]: i=0: [0.5, 2.5], i=1: [-0.5, 1.5]; huber_1 = 0.125, 2.0, 0.125, 1.0; '
                   'weights |tau_i - 1[delta<0]|: [0.25,0.25], [0.25,0.75]; mean over j, sum over i = 0.265625 + 0.390625'},
    {'dist_src': [0.0, 1.0], 'tau': [0.25, 0.75], 'target': [0.5, 2.5], 'kappa': 0.0,
     'expected': (0.25 * 0.5 + 0.25 * 2.5) / 2 + (0.25 * 0.5 + 0.75 * 1.5) / 2,
     'derivation': 'kappa=0 -> |delta|: i=0: (0.125+0.625)/2 = 0.375; i=1: (0.125+1.125)/2 = 0.625; total 1.0'},
]
td = 0.5 + 0.9 * 20.0 - 2.0
g['double_q_learning'] = {'q_tm1': [1.0, 2.0, 3.0], 'a_tm1': 1, 'r_t': 0.5, 'discount_t': 0.9,
                          'q_t_value': [10.0, 20.0, 30.0], 'q_t_selector': [0.3, 0.9, 0.1],
                          'expected_td': td, 'expected_l2': 0.5 * td * td,
                          'derivation': 'selector argmax = action 1 -> bootstrap 20; td = 0.5 + 0.9*20 - 2 = 16.5; l2 = 136.125'}
m, v = 0.1 * 0.5, 0.001 * 0.25
p1 = 1.0 - 0.1 * (m / 0.1) / (math.sqrt(v / 0.001) + 1e-3)
m2, v2 = 0.9 * m + 0.1 * (-1.0), 0.999 * v + 0.001 * 1.0
p2 = p1 - 0.1 * (m2 / (1 - 0.9 ** 2)) / (math.sqrt(v2 / (1 - 0.999 ** 2)) + 1e-3)
mu, nu = 0.05 * 0.5, 0.05 * 0.25
r1 = 1.0 - 0.1 * 0.5 / math.sqrt(nu - mu * mu + 1e-4)
g['optimizer'] = [
    {'name': 'adam', 'lr': 0.1, 'eps': 1e-3, 'p': [1.0], 'grads': [[0.5]], 'expected_p': [p1],
     'derivation': 'm=0.05, v=0.00025; m_hat=0.5, v_hat=0.25; p = 1 - 0.1*0.5/(sqrt(0.25)+0.001): eps OUTSIDE the sqrt'},
    {'name': 'adam', 'lr': 0.1, 'eps': 1e-3, 'p': [1.0], 'grads': [[0.5], [-1.0]], 'expected_p': [p2],
     'derivation': 'second step, g=-1: m=-0.055, v=0.00124975; bias corrections 1-0.9^2, 1-0.999^2'},
    {'name': 'rmsprop', 'lr': 0.1, 'eps': 1e-4, 'decay': 0.95, 'p': [1.0], 'grads': [[0.5]], 'expected_p': [r1],
     'derivation': 'centred rmsprop, zero init: mu=0.025, nu=0.0125; p = 1 - 0.1*0.5/sqrt(nu - mu^2 + eps): eps INSIDE the sqrt'},
    {'name': 'adam', 'lr': 0.1, 'eps': 1e-3, 'max_norm': 1.0, 'p': [0.0, 0.0], 'grads': [[3.0, 4.0]],
     'expected_p': [-0.1 * 0.6 / (0.6 + 1e-3), -0.1 * 0.8 / (0.8 + 1e-3)], 'expected_norm': 5.0,
     'derivation': 'clip_by_global_norm(1): |g|=5 -> [0.6,0.8]; adam step 1: m_hat=g, v_hat=g^2 -> g/(|g|+eps)'},
]
g['noisy_linear'] = {'x': [1.0, 2.0], 'mu_w': [[1.0, 0.0, -1.0], [0.5, 2.0, 1.0]], 'mu_b': [0.1, 0.2, 0.3],
                     'sigma_w': [[0.1, 0.2, 0.3], [0.4, 0.5, 0.6]], 'sigma_b': [1.0, 1.0, 1.0],
                     'eps_in': [2.0, -1.0], 'eps_out': [1.0, 0.5, -2.0],
                     'expected': [2.1 + (0.2 - 0.8 + 1.0) * 1.0, 4.2 + (0.4 - 1.0 + 1.0) * 0.5, 1.3 + (0.6 - 1.2 + 1.0) * (-2.0)],
                     'derivation': 'mu: xW+b = [2.1, 4.2, 1.3]; eps_in*x = [2,-2]; sigma: [0.4, 0.4, 0.4]; y = mu + sigma*eps_out '
                                   '(networks.py:160-178: the sigma layer always has a bias)'}
g['dueling'] = {'adv': [[1.0, 2.0], [3.0, 6.0]], 'val': [10.0, 20.0],
                'expected': [[9.0, 18.0], [11.0, 22.0]],
                'derivation': 'logits[a][k] = val[k] + adv[a][k] - mean_a adv[a][k] (networks.py:251); column means [2, 4]'}

if __name__ == '__main__':
  out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'learner_hand_vectors.json')
  with open(out, 'w') as f:
    json.dump(g, f, indent=1)
  print('wrote', out)
