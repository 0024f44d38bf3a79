"""Drop-in replacement for the reference's ``models`` package (models/__init__.py:2):
put ``<repo>/dropin`` ahead of the reference checkout on ``sys.path`` and
``from models import *`` in train_mvs4.py / test_mvs4.py resolves to the B200 implementation."""
from mvster_b200 import MVS4net, MVS4net_loss, Blend_loss  # noqa: F401
