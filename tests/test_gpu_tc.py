"""GPU: the tcgen05 (3xTF32) GEMM against float64 numpy, through the C ABI self-test hook.

Expected accuracy: each product carries ~2^-21 relative error (hi*hi + hi*lo + lo*hi of an 11+11 bit
split), far inside the 1e-5 parity bar; a plain 1xTF32 product would be ~5e-4 and fail this test."""

import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def run_tc(A_src, a_red_is_b, B_src, b_red_is_b, MI, NJ, R, splits=1, tile_n=64, scale_a=None, scale_b=None,
           ones_row=-1, transpose_out=False):
  from dqn_zoo_b200 import _lib
  dev = 'cuda'
  dA = torch.as_tensor(A_src, device=dev).contiguous()
  dB = torch.as_tensor(B_src, device=dev).contiguous()
  out = torch.full((splits, MI, NJ), float('nan'), dtype=torch.float32, device=dev)
  sc_i, sc_j = (1, MI) if transpose_out else (NJ, 1)
  if transpose_out:
    out = torch.full((splits, NJ, MI), float('nan'), dtype=torch.float32, device=dev)
  sa = None if scale_a is None else torch.as_tensor(scale_a, device=dev).contiguous()
  sb = None if scale_b is None else torch.as_tensor(scale_b, device=dev).contiguous()
  _lib.call('dz_test_tc_gemm', dA.data_ptr(), dA.shape[0], dA.shape[1], dA.shape[1], a_red_is_b,
            dB.data_ptr(), dB.shape[0], dB.shape[1], dB.shape[1], b_red_is_b,
            0 if sa is None else sa.data_ptr(), 0 if sb is None else sb.data_ptr(), ones_row, out.data_ptr(), MI, NJ, R,
            sc_i, sc_j, splits, MI * NJ, tile_n, torch.cuda.current_stream().cuda_stream)
  torch.cuda.synchronize()
  res = out.sum(0).cpu().numpy().astype(np.float64)
  return res.T if transpose_out else res


def rel(got, want):
  return np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30)


@pytest.mark.parametrize('tile_n', [32, 64, 128])
@pytest.mark.parametrize('MI,NJ,R,splits', [(128, 64, 64, 1), (256, 128, 256, 1), (200, 50, 100, 1), (130, 70, 328, 3),
                                            (64, 512, 3136, 14)])
def test_all_operand_majors(MI, NJ, R, splits, tile_n):
  rs = np.random.RandomState(MI + NJ + R)
  Am = rs.standard_normal((MI, R)).astype(np.float32)     # logical A(i, r)
  Bm = rs.standard_normal((NJ, R)).astype(np.float32)     # logical B(j, r)
  want = Am.astype(np.float64) @ Bm.astype(np.float64).T
  # K-major x K-major: sources are [rows][r]
  got = run_tc(Am, 1, Bm, 1, MI, NJ, R, splits, tile_n)
  assert rel(got, want) < 3e-6, ('KK', rel(got, want))
  if True:
    # K-major x MN-major (forward NN: weights [R][NJ])
    got = run_tc(Am, 1, np.ascontiguousarray(Bm.T), 0, MI, NJ, R, splits, tile_n)
    assert rel(got, want) < 3e-6, ('K-MN', rel(got, want))
    # MN-major x MN-major (weight gradient: sources [r][rows])
    got = run_tc(np.ascontiguousarray(Am.T), 0, np.ascontiguousarray(Bm.T), 0, MI, NJ, R, splits, tile_n)
    assert rel(got, want) < 3e-6, ('MN-MN', rel(got, want))
    # MN-major x K-major, transposed store (FC forward with weights as the MMA "A" operand)
    got = run_tc(np.ascontiguousarray(Am.T), 0, Bm, 1, MI, NJ, R, splits, tile_n, transpose_out=True)
    assert rel(got, want) < 3e-6, ('MN-K', rel(got, want))


def test_scales_and_ones_row():
  rs = np.random.RandomState(5)
  MI, NJ, R = 260, 64, 96
  Am = rs.standard_normal((MI, R)).astype(np.float32)
  Bm = rs.standard_normal((NJ, R)).astype(np.float32)
  s = rs.uniform(0.5, 1.5, R).astype(np.float32)
  want = (Am.astype(np.float64) * s[None, :].astype(np.float64)) @ Bm.astype(np.float64).T
  assert rel(run_tc(Am, 1, Bm, 1, MI, NJ, R, 1, 64, scale_a=s), want) < 3e-6
  assert rel(run_tc(Am, 1, Bm, 1, MI, NJ, R, 1, 64, scale_b=s), want) < 3e-6
  # bias-gradient trick: MN-major A with an extra row (index MI) reading as ones -> column sums of B^T
  At = np.ascontiguousarray(Am.T)                              # [r][rows]
  got = run_tc(At, 0, np.ascontiguousarray(Bm.T), 0, MI + 1, NJ, R, 1, 64, ones_row=MI)
  want2 = np.concatenate([Am.astype(np.float64) @ Bm.astype(np.float64).T, Bm.astype(np.float64).sum(1)[None, :]])
  assert rel(got, want2) < 3e-6


def test_accuracy_is_fp32_grade_not_tf32_grade():
  """The tensor core adds into its fp32 accumulator with round-towards-zero, which costs ~2.4e-9 of
  relative bias per reduction element of one uninterrupted accumulation run (measured: R=1024 ->
  2.5e-6).  The learner therefore keeps runs <= 256 elements (split-K, partials added in fp32 RN),
  where the result is as good as an fp32 FMA loop; a plain 1xTF32 product would sit at ~5e-4."""
  rs = np.random.RandomState(1)
  A = rs.standard_normal((128, 1024)).astype(np.float32)
  B = rs.standard_normal((64, 1024)).astype(np.float32)
  want = A.astype(np.float64) @ B.astype(np.float64).T
  fp32 = (torch.tensor(A) @ torch.tensor(B).T).numpy().astype(np.float64)
  one_run = rel(run_tc(A, 1, B, 1, 128, 64, 1024), want)
  split8 = rel(run_tc(A, 1, B, 1, 128, 64, 1024, splits=8), want)
  assert one_run < 4e-6, one_run
  assert split8 < 2 * max(rel(fp32, want), 2e-7), (split8, rel(fp32, want))


# ---- packed-operand kernel (dz_tcp.cuh): the IQN 3136 -> 512 layer's shapes -----------------------------------

def run_pgemm(A_src, a_contig, B_src, b_contig, a_rows, b_rows, red, splits=1, ones_row=False, bias=None, relu=False,
              transpose_out=False):
  from dqn_zoo_b200 import _lib
  dev = 'cuda'
  dA = torch.as_tensor(A_src, device=dev).contiguous()
  dB = torch.as_tensor(B_src, device=dev).contiguous()
  MI = a_rows + (1 if ones_row else 0)
  work = torch.empty(int(_lib.lib.dz_test_tc_pgemm_work(a_rows, b_rows, red)), dtype=torch.float32, device=dev)
  shape = (splits, b_rows, MI) if transpose_out else (splits, MI, b_rows)
  out = torch.full(shape, float('nan'), dtype=torch.float32, device=dev)
  sc_i, sc_j = (1, MI) if transpose_out else (b_rows, 1)
  db = None if bias is None else torch.as_tensor(bias, device=dev).contiguous()
  _lib.call('dz_test_tc_pgemm', dA.data_ptr(), a_rows, dA.shape[1], a_contig, dB.data_ptr(), b_rows, dB.shape[1], b_contig,
            red, a_rows if ones_row else -1, work.data_ptr(), out.data_ptr(), sc_i, sc_j, splits, MI * b_rows,
            0 if db is None else db.data_ptr(), int(relu), torch.cuda.current_stream().cuda_stream)
  torch.cuda.synchronize()
  res = out.sum(0).cpu().numpy().astype(np.float64)
  return res.T if transpose_out else res


@pytest.mark.parametrize('a_rows,b_rows,red,splits', [(128, 256, 16, 1), (128, 256, 128, 1), (128, 256, 272, 1),
                                                      (300, 512, 1000, 1), (2048, 512, 3136, 3), (200, 70, 100, 2),
                                                      (384, 3136, 512, 1)])
def test_packed_gemm_all_source_orientations(a_rows, b_rows, red, splits):
  rs = np.random.RandomState(a_rows + b_rows + red)
  Am = rs.standard_normal((a_rows, red)).astype(np.float32)
  Bm = rs.standard_normal((b_rows, red)).astype(np.float32)
  want = Am.astype(np.float64) @ Bm.astype(np.float64).T
  scale = np.sqrt(red)
  for a_contig in (1, 0):
    for b_contig in (1, 0):
      got = run_pgemm(Am if a_contig else np.ascontiguousarray(Am.T), a_contig, Bm if b_contig else np.ascontiguousarray(Bm.T),
                      b_contig, a_rows, b_rows, red, splits)
      assert np.isfinite(got).all()
      assert rel(got, want) < 2e-6, (a_contig, b_contig, rel(got, want))
      assert np.abs(got - want).max() / scale < 2e-5
  got = run_pgemm(Am, 1, Bm, 1, a_rows, b_rows, red, splits, transpose_out=True)
  assert rel(got, want) < 2e-6


def test_packed_gemm_bias_relu_and_ones_row():
  rs = np.random.RandomState(5)
  a_rows, b_rows, red = 333, 512, 2048
  Am = rs.standard_normal((a_rows, red)).astype(np.float32)
  Bm = rs.standard_normal((b_rows, red)).astype(np.float32)
  bias = rs.standard_normal(b_rows).astype(np.float32)
  want = np.maximum(Am.astype(np.float64) @ Bm.astype(np.float64).T + bias, 0.0)
  got = run_pgemm(Am, 1, Bm, 1, a_rows, b_rows, red, 1, bias=bias, relu=True)
  assert np.abs(got - want).max() < 2e-4 and rel(got, want) < 2e-6
  # weight-gradient form: sources [red][rows], an appended row of ones yields the column sums (bias gradient)
  got = run_pgemm(np.ascontiguousarray(Am.T), 0, np.ascontiguousarray(Bm.T), 0, a_rows, b_rows, red, 5, ones_row=True)
  want = np.concatenate([Am.astype(np.float64) @ Bm.astype(np.float64).T, Bm.astype(np.float64).sum(1)[None]], 0)
  assert got.shape == want.shape and rel(got, want) < 2e-6


def test_packed_gemm_sign_consistent_sum_has_no_truncation_bias():
  """All-positive operands: a round-towards-zero accumulator run over the whole reduction would lose ~2e-8 per
  accumulation (3136/8*3 of them); draining every 128 elements keeps the result at fp32 grade."""
  rs = np.random.RandomState(9)
  Am = rs.uniform(0.5, 1.5, size=(256, 3136)).astype(np.float32)
  Bm = rs.uniform(0.5, 1.5, size=(256, 3136)).astype(np.float32)
  want = Am.astype(np.float64) @ Bm.astype(np.float64).T
  got = run_pgemm(Am, 1, Bm, 1, 256, 256, 3136, 1)
  bias = ((got - want) / want).mean()
  assert abs(bias) < 1.5e-6, bias
  assert np.abs((got - want) / want).max() < 3e-6
