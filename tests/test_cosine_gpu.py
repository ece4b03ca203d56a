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


def test_pos_neg_scoring_consumers(cuda_device):
    """sigmoid(cos / t), the 2-way InfoNCE and its corrupt penalty (R/training/trainer_unfreeze.py:716-741, 924-939),
    checked against the float64 oracle and against torch's own float32 ops on the CPU (the reference's arithmetic)."""
    import torch.nn.functional as F
    rng = np.random.default_rng(5)
    for B, D in ((8, 768), (257, 1024), (1, 7)):
        aud, pos = synth.embedding_pairs(B, D, seed=B)
        neg = (pos + 0.8 * rng.standard_normal((B, D))).astype(np.float32) * np.float32(2.5)      # not normalised
        fac = (1.0 - 0.3 / (1.0 + np.exp(-rng.standard_normal(B)))).astype(np.float32)
        for gamma, factor in ((0.35, None), (0.0, fac)):
            got = ops.score_pos_neg(_dev(aud, cuda_device), _dev(pos, cuda_device), _dev(neg, cuda_device), 0.1, gamma,
                                    _dev(factor, cuda_device) if factor is not None else None)
            ref = OC.pos_neg_reference(aud, pos, neg, 0.1, gamma, factor)
            for k in ("s_pos", "s_neg"):
                assert np.abs(got[k].cpu().numpy() - ref[k]).max() <= TOL
            for k in ("hr_pos", "hr_neg"):
                assert np.abs(got[k].cpu().numpy() - ref[k]).max() <= 2e-5          # sigmoid(s / 0.1): slope <= 2.5
            assert np.abs(got["per_sample"].cpu().numpy() - ref["per_sample"]).max() <= 1e-4   # d/ds <= 10
            assert abs(float(got["loss"]) - ref["loss"]) <= 1e-4
            # the reference's own torch ops (CPU, float32)
            ta, tp, tn = (F.normalize(torch.from_numpy(x), p=2, dim=1) for x in (aud, pos, neg))
            s_pos, s_neg = (ta * tp).sum(1), (ta * tn).sum(1)
            per = F.cross_entropy(torch.stack([s_pos, s_neg], 1) / 0.1, torch.zeros(B, dtype=torch.long), reduction="none")
            if factor is not None:
                per = per * torch.from_numpy(factor)
            loss = per.mean() + (gamma * F.relu(s_neg).mean() if gamma > 0 else 0.0)
            assert abs(float(got["loss"]) - float(loss)) <= 1e-4
            assert (got["hr_pos"].cpu() - torch.sigmoid(s_pos / 0.1)).abs().max().item() <= 2e-5


@pytest.mark.parametrize("N,M,D,k", [(4096, 4096, 768, 8), (1000, 1003, 100, 5), (130, 3, 64, 8), (7, 1, 32, 1), (300, 129, 1024, 3)])
def test_topk_equals_sorted_rows_of_the_matrix(cuda_device, N, M, D, k):
    """Retrieval without the N x M matrix (stx_cosine_topk): bit-identical to sorting stx_cosine_nxm's rows by (score
    descending, column ascending) -- same kernel, same arithmetic -- and within 1e-5 of the float64 oracle's top-k scores."""
    from speech_transcript_embeddings_b200 import ops
    rng = np.random.default_rng(N + M)
    a = rng.standard_normal((N, D)).astype(np.float32)
    b = (a[rng.integers(0, N, size=M)] + 0.7 * rng.standard_normal((M, D))).astype(np.float32)
    ta, tb = torch.from_numpy(a).to(cuda_device), torch.from_numpy(b).to(cuda_device)
    val, idx = ops.cosine_topk(ta, tb, k)
    S = ops.cosine_nxm(ta, tb).cpu().numpy()
    order = np.lexsort((np.broadcast_to(np.arange(M), S.shape), -S), axis=1)[:, :k]     # score descending, column ascending
    kk = min(k, M)
    assert np.array_equal(idx.cpu().numpy()[:, :kk], order[:, :kk])
    assert np.array_equal(val.cpu().numpy()[:, :kk], np.take_along_axis(S, order[:, :kk], axis=1))
    if M < k:
        assert bool((idx[:, M:] == -1).all()) and bool(torch.isinf(val[:, M:]).all())
    ref = np.sort(OC.matrix_f64(a, b), axis=1)[:, ::-1][:, :kk]
    assert np.abs(val.cpu().numpy()[:, :kk] - ref).max() <= 1e-5


@pytest.mark.parametrize("D", [768, 1024])
def test_matrix_scores_of_well_matched_pairs_hold_the_bar(cuda_device, D):
    """Scores near 1 are where the tensor core's round-toward-zero accumulation bites: every product has the same sign, so
    each truncation loses in the same direction (about -1.8e-8 * D * score with one accumulator: -1.9e-5 at D = 1024).  With
    the A_hi B_hi sum kept apart from the cross terms the worst case stays under the 1e-5 bar (measured 6.4e-6).
    The seeded cfg5 pairs never exercised this (their cosines are ~0.07)."""
    from speech_transcript_embeddings_b200 import ops
    rng = np.random.default_rng(D)
    a = rng.standard_normal((512, D)).astype(np.float32)
    worst = 0.0
    for noise in (0.0, 0.3, 0.7):
        b = (a + noise * rng.standard_normal((512, D))).astype(np.float32)
        S = ops.cosine_nxm(torch.from_numpy(a).to(cuda_device), torch.from_numpy(b).to(cuda_device)).cpu().numpy()
        ref = OC.matrix_f64(a, b)
        worst = max(worst, float(np.abs(S - ref).max()))
        pw = ops.cosine_pairwise(torch.from_numpy(a).to(cuda_device), torch.from_numpy(b).to(cuda_device),
                                 always_normalize=True).cpu().numpy()
        assert np.abs(np.diag(S) - pw).max() <= 1e-5          # the diagonal equals the pairwise scores
    print(f"D = {D}: worst |S - S_f64| over cos in {{1, 0.96, 0.82}}: {worst:.2e}")
    assert worst <= 1e-5
