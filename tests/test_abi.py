"""CPU: the C-ABI library loads and exports every symbol include/toucan_cuda.h declares;
without a GPU every compute call fails loudly (no CPU fallback)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:tc|tr)_[a-z0-9_]+)\s*\(", txt)))


def test_cuda_abi_exports(built):
    from toucan_b200 import capi

    lib = capi.lib()
    names = _declared("toucan_cuda.h")
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in toucan_cuda.h but not exported"
    assert set(capi.EXPORTS) <= set(names)
    assert b"sm_100a" in lib.tc_version()


def test_type_helpers(built):
    from toucan_b200 import capi

    lib = capi.lib()
    assert [lib.tc_type_size(t) for t in range(4)] == [1, 2, 2, 4]
    assert lib.tc_image_bytes(7680, 4320, capi.F32) == 530841600
    # OIIO basetype_merge cases used on the path
    U8, U16, F16, F32 = 0, 1, 2, 3
    assert lib.tc_type_merge(U8, U8) == U8
    assert lib.tc_type_merge(U8, F32) == F32
    assert lib.tc_type_merge(U8, U16) == U16
    assert lib.tc_type_merge(F16, U8) == F16
    assert lib.tc_type_merge(U16, F16) == F32


def test_no_cpu_fallback(built):
    import torch

    from toucan_b200 import capi

    if torch.cuda.is_available():
        return
    lib = capi.lib()
    n = ctypes.c_int(-1)
    st = lib.tc_device_count(ctypes.byref(n))
    assert st != 0 and n.value == 0
    import numpy as np

    host = np.zeros((4, 4, 4), np.float32)
    im = capi.tc_image(host.ctypes.data, 4, 4, capi.F32)
    st = lib.tc_fill(ctypes.byref(im), capi.f4([1, 0, 0, 1]), None)
    assert st != 0, "compute call must fail without a CUDA device"
    assert host.sum() == 0


def test_argument_validation(built):
    from toucan_b200 import capi

    lib = capi.lib()
    im = capi.tc_image(None, 4, 4, capi.F32)
    assert lib.tc_fill(ctypes.byref(im), capi.f4([0, 0, 0, 0]), None) == -1
    assert b"null data" in lib.tc_last_error()
    im = capi.tc_image(16, 0, 4, capi.F32)
    assert lib.tc_fill(ctypes.byref(im), capi.f4([0, 0, 0, 0]), None) == -1
    im = capi.tc_image(16, 4, 4, 9)
    assert lib.tc_fill(ctypes.byref(im), capi.f4([0, 0, 0, 0]), None) == -1
