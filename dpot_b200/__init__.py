"""dpot_b200: B200-native (sm_100a) implementation of DPOT's autoregressive Fourier-operator
hot path behind the reference's own Python API.  See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1.0"

from . import _lib  # noqa: F401


def lib():
    """The loaded libdpot_b200.so (raises if it has not been built)."""
    return _lib.load()
