"""GPU parity tests for cosine scoring through the C ABI.  Bar: max-abs <= 1e-5."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import cosine as OC
from speech_transcript_embeddings_b200 import ops, scoring, synth

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _dev(x, device):
    return torch.from_numpy(np.ascontiguousarray(x)).to(device)


def test_golden_pairwise_and_matrix(cuda_device):
    g = load_golden("cosine.npz")
    a, b = g["a"], g["b"]
    for key, (sa, sb) in {"pair_unit": (1.0, 1.0), "pair_scaled": (3.0, 0.25), "pair_near_unit": (1.00005, 1.0)}.items():
        got = ops.cosine_pairwise(_dev(a * np.float32(sa), cuda_device), _dev(b * np.float32(sb), cuda_device)).cpu().numpy()
        assert got.dtype == np.float32 and got.shape == (64,)
        assert np.abs(got - g[key]).max() <= TOL, key
    S = ops.cosine_nxm(_dev(a, cuda_device), _dev(b, cuda_device)).cpu().numpy()
    assert np.abs(S - g["matrix_f64"]).max() <= TOL


def test_conditional_normalisation_matches_reference(cuda_device):
    a, b = synth.embedding_pairs(33, 100, seed=4)
    a2 = a.copy()
    a2[7] *= 1.0002                                  # one row off by 2e-4 -> the whole operand is re-normalised
    a3 = a * np.float32(1.00005)                     # all rows within 1e-4 -> left alone
    for x in (a, a2, a3):
        got = ops.cosine_pairwise(_dev(x, cuda_device), _dev(b, cuda_device)).cpu().numpy()
        assert np.abs(got - OC.pairwise_reference(x, b)).max() <= 2e-6


@pytest.mark.parametrize("N,M,D", [(4096, 4096, 768), (512, 4096, 1024), (37, 129, 100), (1, 1, 1), (65, 63, 17)])
def test_matrix_vs_float64(cuda_device, N, M, D):
    a, _ = synth.embedding_pairs(N, D, seed=1)
    _, b = synth.embedding_pairs(M, D, seed=2)
    a = a * np.float32(1.7)                          # not pre-normalised
    S = ops.cosine_nxm(_dev(a, cuda_device), _dev(b, cuda_device)).cpu().numpy()
    ref = OC.matrix_f64(a, b)
    err = np.abs(S - ref).max()
    print(f"cosine {N}x{M}x{D}: max-abs {err:.2e}")
    assert err <= TOL


def test_diagonal_equals_pairwise(cuda_device):
    a, b = synth.embedding_pairs(4096, 768, seed=0)
    ad, bd = _dev(a, cuda_device), _dev(b, cuda_device)
    S = scoring.cosine_matrix(ad, bd)
    p = ops.cosine_pairwise(ad, bd, always_normalize=True)
    # tensor-core (split-TF32, fp32 accumulate in TMEM) matrix vs CUDA-core pairwise: both within 1e-5 of float64
    assert (S.diagonal() - p).abs().max().item() <= 5e-6
    assert np.abs(p.cpu().numpy() - OC.pairwise_reference(a, b)).max() <= TOL


def test_zero_rows_do_not_produce_nans(cuda_device):
    a = np.zeros((4, 32), np.float32)
    b = np.ones((4, 32), np.float32)
    got = ops.cosine_pairwise(_dev(a, cuda_device), _dev(b, cuda_device)).cpu().numpy()
    assert np.array_equal(got, np.zeros(4, np.float32))     # x / max(|x|, 1e-12) = 0, like F.normalize
