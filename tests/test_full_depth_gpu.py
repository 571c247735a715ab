"""Full-depth parity of the published sizes against the UNMODIFIED reference run on the same GPU: BASELINE.json configs
3-5 name DPOT-M, DPOT-L (256^2, patch 16, modes 64) and DPOT-H; the golden fixtures and the numpy oracle cover them at
reduced depth only (a float64 numpy forward of a 1 B-parameter model does not fit a test budget).

Ground truth = the reference DPOTNet constructed under float64 defaults (the only way to run it in fp64) with
the same weights, evaluated in fp64 on the GPU.  Two numbers are held: ours against that truth (the 1e-5 bar of
north_star), and -- as context, printed -- the reference's own fp32 eager result (cuBLAS / cuDNN / cuFFT, allow_tf32 =
False) against the same truth: fp32 arithmetic 24-27 blocks deep is itself a few 1e-6 away from exact, so a comparison of
two fp32 paths with each other would measure the sum of both errors.  One thing is shared with the fp32 paths on purpose:
the time-embedding table cos(linspace(0,1,T) x gamma) (models/dpot.py:230-231) is evaluated in fp32 -- gamma reaches 1024,
so the fp32 rounding of the cosine's ARGUMENT moves the table by up to 6e-5; the reference and this library compute that
table with the same fp32 torch ops, bit-identically, so it is an input of the contraction under test, not an error of it.

The reference travels to the GPU box as git-ignored baseline/_ref (staged by __graft_entry__.build() in the build
container); without it the tests skip."""
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-5

SHAPES = {"M": dict(img_size=128, patch_size=8), "L": dict(img_size=256, patch_size=16, modes=64), "H": dict(img_size=128, patch_size=8)}


def _reference():
    from baseline.install_ref import import_reference
    ref = import_reference()
    if ref is None:
        pytest.skip("baseline/_ref (the unmodified reference) is not staged on this box")
    return ref[0]


def _rel(a, b):
    return float((a.double() - b).norm() / b.norm())


def _truth(RefNet, cfg, sd, x):
    ref64 = RefNet(**cfg)
    ref64.load_state_dict(sd)
    ref64 = ref64.to(x.device).eval()
    tagg = ref64.time_agg_layer
    if tagg.type == "exp_mlp":
        t32 = torch.linspace(0, 1, cfg["in_timesteps"], dtype=torch.float32).unsqueeze(-1).to(x.device)
        table = torch.cos(t32 @ tagg.gamma.float()).double()
        tagg.forward = lambda z: torch.einsum("tij,...ti->...j", tagg.w, z * table)
    return ref64(x.double())


@pytest.mark.parametrize("name,B", [("M", 2), ("L", 1), ("H", 1)])
def test_full_depth_forward_matches_reference(name, B):
    from dpot_b200 import zoo
    from dpot_b200.models.dpot import DPOTNet
    RefNet = _reference()
    cfg = zoo.zoo_cfg(name, **SHAPES[name])
    ours = zoo.synthetic_weights_(DPOTNet(**cfg), seed=0)
    sd = ours.state_dict()
    g = torch.Generator().manual_seed(7)
    x = torch.randn((B, cfg["img_size"], cfg["img_size"], 10, 4), generator=g).cuda()
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            torch.set_default_dtype(torch.float64)
            try:  # construction AND forward under fp64 defaults: the spectral buffers (models/dpot.py:64-67) take the default
                y64, c64 = _truth(RefNet, cfg, sd, x)
            finally:
                torch.set_default_dtype(torch.float32)
            ref = RefNet(**cfg)
            ref.load_state_dict(sd)
            ref = ref.cuda().eval()
            y_ref, c_ref = ref(x)
            del ref
            ours = ours.cuda().eval()
            y, c = ours(x)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    assert y64.dtype == torch.float64
    e_y, e_c = _rel(y, y64), _rel(c, c64)
    r_y, r_c = _rel(y_ref, y64), _rel(c_ref, c64)
    print(f"\nDPOT-{name} full depth ({cfg['depth']} blocks) vs fp64 reference: ours y {e_y:.2e} cls {e_c:.2e} | "
          f"reference fp32 eager y {r_y:.2e} cls {r_c:.2e} | ours vs fp32 eager y {_rel(y, y_ref.double()):.2e}")
    assert e_y < TOL and e_c < TOL, (e_y, e_c)


def test_half_mode_against_reference_bf16_autocast():
    """BASELINE.json config 3 names a bf16 DPOT-M step.  The reference has no reduced-precision code of its own; what a
    user would run is its DPOTNet under torch.autocast(bfloat16).  This library's 16-bit mode is set_precision("half"):
    fp16 operands (11-bit significand), one tensor-core MMA per product, fp32 accumulation / epilogues / residual stream.
    Both against the fp64 reference at full depth: the half mode must be at least as close to it as the autocast run
    (measured: ~3e-4 vs ~1e-2), and the 1e-5 bar applies to the default fp32-faithful mode only."""
    import dpot_b200
    from dpot_b200 import zoo
    from dpot_b200.models.dpot import DPOTNet
    RefNet = _reference()
    cfg = zoo.zoo_cfg("M", **SHAPES["M"])
    ours = zoo.synthetic_weights_(DPOTNet(**cfg), seed=0)
    sd = ours.state_dict()
    x = torch.randn((2, 128, 128, 10, 4), generator=torch.Generator().manual_seed(7)).cuda()
    with torch.no_grad():
        torch.set_default_dtype(torch.float64)
        try:
            y64, _ = _truth(RefNet, cfg, sd, x)
        finally:
            torch.set_default_dtype(torch.float32)
        ref = RefNet(**cfg)
        ref.load_state_dict(sd)
        ref = ref.cuda().eval()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y_bf16, _ = ref(x)
        del ref
        ours = ours.cuda().eval()
        prev = dpot_b200.set_precision("half")
        try:
            y_half, _ = ours(x)
        finally:
            dpot_b200.set_precision(prev)
        y_fp32, _ = ours(x)
    e_half, e_bf16, e_fp32 = _rel(y_half, y64), _rel(y_bf16.float(), y64), _rel(y_fp32, y64)
    print(f"\nDPOT-M full depth vs fp64 reference: half mode {e_half:.2e} | reference under bf16 autocast {e_bf16:.2e} | "
          f"fp32-faithful mode {e_fp32:.2e}")
    assert e_half < e_bf16 and e_half < 2e-3 and e_fp32 < TOL
