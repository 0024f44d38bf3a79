"""mvster_b200 - sm_100a (B200) implementation of MVSTER's per-frame forward hot path behind
the reference's ``models`` API: ``from mvster_b200 import MVS4net, MVS4net_loss, Blend_loss``.
See DESIGN.md for scope and INTEGRATION.md for the drop-in recipe."""
from .network import MVS4net  # noqa: F401
from .losses import MVS4net_loss, Blend_loss  # noqa: F401

from . import formats  # noqa: F401  (PFM / camera-file / pair-list I/O compatible with the reference's datasets package)
from . import fusion  # noqa: F401   (geometric-consistency filter of test_mvs4.py on the GPU)
from . import prefetch  # noqa: F401 (asynchronous sample pipeline in front of the forward)

__all__ = ["MVS4net", "MVS4net_loss", "Blend_loss", "formats", "fusion", "prefetch"]
