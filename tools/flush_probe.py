"""Whole-model rel-L2 of DPOT-S (golden fwd_c2_s128) as a function of the in-TMEM accumulation chain length."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dpot_oracle as O
from dpot_b200 import _lib
from dpot_b200.models.dpot import DPOTNet
from dpot_b200.rollout import rollout
lib = _lib.load()
z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "fwd_c2_s128.npz"))
cfg = json.loads(str(z["cfg"]))
params = O.make_params(cfg, seed=0)
x = O.make_input(cfg, int(z["B"]), seed=0, kind=str(z["kind"]))
m = DPOTNet(**cfg)
m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in params.items()})
m = m.cuda().eval()
xt = torch.from_numpy(x).cuda()
for fl in [1, 2, 4, 6, 8, 32]:
    lib.dpot_tc_set_flush(fl)
    with torch.no_grad():
        y, cls = m(xt)
        pred = rollout(m, xt, int(z["nsteps"]))
    torch.cuda.synchronize()
    print(f"flush={fl:3d} (chain {12*fl} MMAs): y {O.rel_l2(y.cpu().numpy(), z['y']):.2e}  rollout {O.rel_l2(pred.cpu().numpy(), z['pred']):.2e}", flush=True)
