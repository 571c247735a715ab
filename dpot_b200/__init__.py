"""dpot_b200: B200-native (sm_100a) implementation of DPOT's autoregressive Fourier-operator
hot path behind the reference's own Python API.  See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1.0"

from . import _lib  # noqa: F401


def lib():
    """The loaded libdpot_b200.so (raises if it has not been built)."""
    return _lib.load()


def set_precision(mode: str = "fp32") -> str:
    """Operand precision of every dense contraction (forward, backward, weight gradients) of the f16-split tcgen05 engine:
    "fp32" (default): fp32-faithful, three MMAs per product; "half": 16-bit mixed precision -- fp16 operands (the hi
    plane of the split storage), one MMA per product, fp32 accumulation, fp32 master weights / optimizer / epilogues
    (the reference's bf16 autocast configs, e.g. configs/pretrain_medium.yaml; fp16's 11-bit significand is finer than
    bf16's 8 at the same tensor-core rate).  Returns the previous mode."""
    from . import _lib
    if mode not in ("fp32", "half"):
        raise ValueError("precision must be 'fp32' or 'half'")
    prev = _lib.load().dpot_tc16_set_precision(1 if mode == "half" else 0)
    return "half" if prev else "fp32"


def set_chain(k_chain: int = 2048) -> int:
    """Longest tensor-core accumulation chain of one launch of the f16-split engine (dpot_tc16_set_chain): deeper forward
    contractions (the channel MLP's fc2 of DPOT-M / L / H) run as chained launches over K-ranges.  0 = never chain
    (one launch whatever the depth: 1.2e-9 * K relative error).  Returns the previous limit."""
    return int(_lib.load().dpot_tc16_set_chain(int(k_chain)))
