"""Three DPOT-S forwards at the BASELINE shape (B=32, 128x128x10x4) -- target for the ncu launch list."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dpot_oracle as O      # synthetic weights only
from dpot_b200.models.dpot import DPOTNet
cfg = O.zoo_cfg(sys.argv[1] if len(sys.argv) > 1 else "S")
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
m = DPOTNet(**cfg)
m.load_state_dict({k: torch.from_numpy(v) for k, v in O.make_params(cfg, seed=0).items()})
m = m.cuda().eval()
x = torch.randn(B, cfg["img_size"], cfg["img_size"], cfg["in_timesteps"], cfg["in_channels"], device="cuda")
with torch.no_grad():
    for i in range(3):
        if i == 2:                      # ncu --profile-from-start off: only the third forward is captured
            torch.cuda.synchronize(); torch.cuda.profiler.start()
        y, c = m(x)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("ok", float(y.abs().mean()))
