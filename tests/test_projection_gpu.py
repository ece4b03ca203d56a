"""GPU: the encoder's input stage (LayerNorm(160) + Linear(160 -> 1024)) on tcgen05 against torch's own modules
(the reference model's Wav2Vec2BertFeatureProjection, TF/models/wav2vec2_bert/modeling_wav2vec2_bert.py:118-130)."""
import numpy as np
import pytest
import torch

from speech_transcript_embeddings_b200 import ops, synth
from speech_transcript_embeddings_b200.feature_extraction import B200SeamlessM4TFeatureExtractor

pytestmark = pytest.mark.gpu


def _modules(in_dim, out_dim, seed):
    g = torch.Generator().manual_seed(seed)
    ln = torch.nn.LayerNorm(in_dim, eps=1e-5)
    lin = torch.nn.Linear(in_dim, out_dim)
    with torch.no_grad():
        ln.weight.copy_(1.0 + 0.1 * torch.randn(in_dim, generator=g))
        ln.bias.copy_(0.1 * torch.randn(in_dim, generator=g))
        lin.weight.copy_(0.05 * torch.randn(out_dim, in_dim, generator=g))
        lin.bias.copy_(0.1 * torch.randn(out_dim, generator=g))
    return ln, lin


@pytest.mark.parametrize("rows,in_dim,out_dim", [(1499, 160, 1024), (300, 160, 1024), (37, 100, 65), (1, 160, 1024)])
def test_feature_projection_matches_torch(cuda_device, rows, in_dim, out_dim):
    ln, lin = _modules(in_dim, out_dim, rows)
    x = torch.randn(rows, in_dim, generator=torch.Generator().manual_seed(1)) * 1.3 + 0.2
    with torch.no_grad():
        norm64 = ln.double()(x.double())
        ref64 = lin.double()(norm64)
    ln, lin = ln.float(), lin.float()
    dev = cuda_device
    hidden, norm = ops.feature_projection(x.to(dev), ln.weight.detach().to(dev), ln.bias.detach().to(dev),
                                          lin.weight.detach().to(dev), lin.bias.detach().to(dev), eps=1e-5)
    assert hidden.shape == (rows, out_dim) and norm.shape == (rows, in_dim)
    err_n = (norm.cpu().double() - norm64).abs().max().item()
    err_h = (hidden.cpu().double() - ref64).abs().max().item()
    print(f"projection {rows}x{in_dim}->{out_dim}: norm {err_n:.2e}, hidden {err_h:.2e}")
    assert err_n <= 2e-6 and err_h <= 2e-5           # float32-grade: torch's own float32 modules are at ~1e-6 / ~5e-6


def test_projection_of_extractor_output(cuda_device):
    """End of the front end: input_features [B, T', 160] straight into the projection, batch dims preserved."""
    fe = B200SeamlessM4TFeatureExtractor(device=cuda_device)
    clips = [synth.clip("G", 16000, 1), synth.clip("AM", 12000, 2)]
    feats = fe(clips, sampling_rate=16000, return_tensors="pt")["input_features"]
    ln, lin = _modules(160, 1024, 0)
    dev = cuda_device
    hidden, norm = ops.feature_projection(feats, ln.weight.detach().to(dev), ln.bias.detach().to(dev),
                                          lin.weight.detach().to(dev), None)
    assert hidden.shape == feats.shape[:-1] + (1024,)
    with torch.no_grad():
        ref = torch.nn.functional.linear(ln.double()(feats.cpu().double()), lin.weight.double())
    assert (hidden.cpu().double() - ref).abs().max().item() <= 2e-5
