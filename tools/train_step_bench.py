"""Training-step timing of DPOT-S through the drop-in API (train_temporal.py:201-230 with T_ar = 1, noise off):
forward (autograd path) + SimpleLpLoss + backward + clip_grad_norm_ + Adam.step, CUDA events, synthetic data."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dpot_oracle as O      # synthetic weights only
from dpot_b200.models.dpot import DPOTNet
from dpot_b200.utils.optimizer import Adam
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = O.zoo_cfg("S")
m = DPOTNet(**cfg)
m.load_state_dict({k: torch.from_numpy(v) for k, v in O.make_params(cfg, seed=0).items()})
m = m.cuda().train()
opt = Adam(m.parameters(), lr=1e-4, betas=(0.9, 0.9), weight_decay=1e-6)
x = torch.randn(B, 128, 128, 10, 4, device="cuda")
y = torch.randn(B, 128, 128, 1, 4, device="cuda")
def step():
    im, _ = m(x)
    diff = (im - y).reshape(B, -1, 4)
    loss = (diff.norm(dim=1) / y.reshape(B, -1, 4).norm(dim=1)).mean(dim=1).sum()      # SimpleLpLoss, utils/criterion.py:38-59
    opt.zero_grad()
    loss.backward()
    torch.nn.utils.clip_grad_norm_(m.parameters(), 10000.0)
    opt.step()
    return loss
for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    l = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"DPOT-S train step B={B}: {ms:.2f} ms/step = {B / ms * 1e3:.0f} field-steps/s (fwd+bwd+clip+Adam), loss {float(l):.4f}")
