"""Back-to-back kernels that both allocate tensor memory, with no host synchronisation in between: the recipe-K kernel (256
TMEM columns on every SM: the stash of its exchange halves) and the cosine kernel (all 512 columns, one CTA per SM).  Every
SM has to hand its tensor memory from one grid to the next; a leak or a missed deallocation would block tcgen05.alloc forever,
so tests/test_tmem_handover_gpu.py runs this under a timeout.

    python tests/scripts/tmem_handover.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from speech_transcript_embeddings_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
B, n = 64, 480000
pcm = 0.1 * torch.randn(B * n, generator=torch.Generator(device=dev).manual_seed(0), device=dev)
off = torch.arange(B, device=dev, dtype=torch.int64) * n
ln = torch.full((B,), n, dtype=torch.int32, device=dev)
T_pad = 2 * ((ops.k_num_frames(n) + 1) // 2)
a = torch.randn(4096, 768, device=dev)
b = torch.randn(4096, 768, device=dev)
w = 0.05 * torch.randn(1024, 160, device=dev)
g, z = torch.ones(160, device=dev), torch.zeros(160, device=dev)
ref_f, ref_s = None, None
for it in range(6):
    f, _ = ops.fbank_k(pcm, off, ln, n, T_pad, uniform=bool(it & 1))        # scheduled and plain item lists alternate
    s = ops.cosine_nxm(a, b)                                                 # 148 CTAs x 512 TMEM columns
    h, _ = ops.feature_projection(f, g, z, w, None, return_norm=False)       # the same tensor-core kernel on the features
    v, i = ops.cosine_topk(a, b, 8)
    f2, _ = ops.fbank_k(pcm, off, ln, n, T_pad, uniform=True)
    if ref_f is None:
        ref_f, ref_s = f.clone(), s.clone()
    assert torch.equal(f, ref_f) and torch.equal(f2, ref_f) and torch.equal(s, ref_s)
torch.cuda.synchronize()
assert bool(torch.isfinite(h).all())
print("tmem hand-over ok")
