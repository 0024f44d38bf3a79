"""Points mvster_b200's ctypes layer at the CPU emulation library (tests only; see simt_emu.h / build_emu.py).
``install()`` patches process-wide and is meant for spawned worker processes; pytest tests in the main process use the
``emu`` fixture of tests/test_emu_kernels.py, which restores everything afterwards."""
import contextlib
import ctypes as C
import sys
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import build_emu  # noqa: E402

HOST_ONLY = {"mvster_conv_tc3_plan", "mvster_conv_tc3_packed_bytes", "mvster_conv_tc3_supported", "mvster_deconv_tc3_packed_bytes",
             "mvster_deconv_tc3_supported", "mvster_conv3d_tc2_supported"}


class EmuWithHostLogic:
    """The emulation library for everything it exports; pure host logic of the tensor-core files (layer plans, packed sizes -
    no kernel launches) is answered by the real library, which loads without a GPU."""

    def __init__(self, emu, real):
        self._emu, self._real = emu, real

    def __getattr__(self, name):
        return getattr(self._real if name in HOST_ONLY else self._emu, name)


def load():
    from mvster_b200 import _lib
    lib = C.CDLL(str(build_emu.build()))
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = None if name in HOST_ONLY else getattr(lib, name, None)
        if fn is not None:
            fn.restype, fn.argtypes = res, args
    real = _lib._lib if isinstance(_lib._lib, C.CDLL) else None
    if real is None:
        saved, _lib._lib = _lib._lib, None
        real = _lib.load()
        _lib._lib = saved
    return EmuWithHostLogic(lib, real)


def cpu_chk(t, name, shape=None, dtype=torch.float32):
    assert isinstance(t, torch.Tensor) and t.dtype == dtype and t.is_contiguous() and not t.is_cuda, name
    assert shape is None or tuple(t.shape) == tuple(shape), (name, tuple(t.shape), tuple(shape))
    return t


def install():
    from mvster_b200 import _lib, capi
    _lib._lib = load()
    capi._chk = cpu_chk
    capi._stream = lambda: C.c_void_p(0)
    torch.cuda.device = lambda d: contextlib.nullcontext()
