"""Make the unmodified reference importable in the build container (test infrastructure).

The reference package imports two third-party modules that are absent here and that contribute
no arithmetic to the hot path:

* ``pytorch_lightning``  - ``DiffAb`` only needs ``nn.Module`` behaviour plus ``log_dict``.
* ``protstruc.general``  - two integer constants, ``ATOM.CA`` (``diffab_pytorch.py:820``,
  equal to the hard-coded ``CA_IDX = 1`` at ``:110,249``) and ``AA.UNK`` (``:115,273``; must index
  ``Embedding(21, .)`` so UNK = 20 is assumed; unpinned by any reference test).

``load_reference()`` pre-seeds ``sys.modules`` with stand-ins for those names and then imports
``diffab_pytorch`` untouched from the first of: ``$DIFFAB_REFERENCE_ROOT``, ``<repo>/baseline/_ref``
(the offline ``pip install --no-deps --target`` of the reference; git-ignored, it travels
to the GPU box with the snapshot) and ``/root/reference`` (build container only).  It returns ``None``
when none of them exists; only the committed goldens are used then.
"""
import enum
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    for cand in (os.environ.get("DIFFAB_REFERENCE_ROOT"), os.path.join(_REPO, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "diffab_pytorch", "diffab_pytorch.py")):
            return cand
    return os.environ.get("DIFFAB_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _find_root()


def _install_standins():
    import torch
    import torch.nn as nn

    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            def log_dict(self, *args, **kwargs):
                return None

            def log(self, *args, **kwargs):
                return None

        class LightningDataModule:
            def __init__(self, *args, **kwargs):
                pass

        pl.LightningModule = LightningModule
        pl.LightningDataModule = LightningDataModule
        pl.seed_everything = torch.manual_seed
        callbacks = types.ModuleType("pytorch_lightning.callbacks")
        callbacks.LearningRateMonitor = object
        pl.callbacks = callbacks
        sys.modules["pytorch_lightning"] = pl
        sys.modules["pytorch_lightning.callbacks"] = callbacks

    if "protstruc" not in sys.modules:
        ps = types.ModuleType("protstruc")
        general = types.ModuleType("protstruc.general")

        class ATOM(enum.IntEnum):
            N = 0
            CA = 1
            C = 2
            O = 3

        class AA(enum.IntEnum):
            UNK = 20

        general.ATOM = ATOM
        general.AA = AA
        ps.general = general
        ps.AntibodyStructureBatch = None
        ps.StructureBatch = None
        sys.modules["protstruc"] = ps
        sys.modules["protstruc.general"] = general


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "diffab_pytorch"))


def load_reference():
    """Import the unmodified reference; returns the ``diffab_pytorch`` package or ``None``."""
    if not reference_available():
        return None
    _install_standins()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    pkg = importlib.import_module("diffab_pytorch")
    importlib.import_module("diffab_pytorch.so3")
    importlib.import_module("diffab_pytorch.diffusion")
    importlib.import_module("diffab_pytorch.diffab_pytorch")
    return pkg
