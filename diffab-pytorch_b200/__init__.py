"""diffab-pytorch_b200: B200-native denoising hot path of DiffAb behind the reference's API.

Host-side mirror of ``diffab_pytorch`` (same module names: ``diffab_pytorch``, ``diffusion``,
``so3``) whose per-timestep work runs in hand-written sm_100a CUDA kernels reached through the flat
C ABI of ``csrc/libdiffab_b200.so`` (``include/diffab_b200.h``).  There is no CPU fallback: the
library is loaded on first use and every op raises if it is missing or if it is handed CPU tensors.
"""
__version__ = "0.1.0"

_LAZY = {
    "DiffAb": "diffab_pytorch",
    "Denoiser": "diffab_pytorch",
    "InvariantPointAttentionLayer": "diffab_pytorch",
    "InvariantPointAttentionModule": "diffab_pytorch",
    "OrientationLoss": "diffab_pytorch",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib

        return getattr(importlib.import_module(f"{__name__}.{_LAZY[name]}"), name)
    raise AttributeError(name)
