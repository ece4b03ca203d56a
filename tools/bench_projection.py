import sys, torch
sys.path.insert(0, '.')
from speech_transcript_embeddings_b200 import ops, _lib
dev = torch.device('cuda', 0)
rows = 64 * 1499
x = torch.randn(rows, 160, device=dev)
lnw, lnb = torch.ones(160, device=dev), torch.zeros(160, device=dev)
w, b = 0.05 * torch.randn(1024, 160, device=dev), torch.zeros(1024, device=dev)
for _ in range(3): h, n = ops.feature_projection(x, lnw, lnb, w, b)
torch.cuda.synchronize()
_lib.profile(True)
for _ in range(5): h, n = ops.feature_projection(x, lnw, lnb, w, b)
torch.cuda.synchronize()
r = _lib.profile_collect(); _lib.profile(False)
import statistics
print({k: round(statistics.mean([m for nm, m in r if nm == k]), 4) for k in set(nm for nm, _ in r)})
ln = torch.nn.LayerNorm(160).to(dev); lin = torch.nn.Linear(160, 1024).to(dev)
torch.backends.cuda.matmul.allow_tf32 = False
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    for _ in range(3): y = lin(ln(x))
    e0.record()
    for _ in range(5): y = lin(ln(x))
    e1.record(); torch.cuda.synchronize()
print("torch fp32 LN+Linear ms:", e0.elapsed_time(e1) / 5)
