"""CPU: the C-ABI library builds for sm_100a, loads without a GPU and exports every symbol the header declares;
the Python binding's struct layout and slot numbers match the C side.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as ge
    ge.build()
    from color_modem_b200 import _native
    return ctypes.CDLL(_native.LIB_PATH)


def test_header_symbols_exported(lib):
    header = open(os.path.join(ROOT, 'include', 'color_modem_b200.h')).read()
    declared = set(re.findall(r'^\s*(?:int|void|int64_t|const char \*)\s*\*?(cm_[a-z0-9_]+)\s*\(', header, re.M))
    assert len(declared) >= 15
    from color_modem_b200 import _native
    assert declared == set(_native.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_struct_layout_and_version(lib):
    from color_modem_b200 import _native
    assert lib.cm_abi_version() == _native.ABI_VERSION
    assert lib.cm_sizeof_desc() == ctypes.sizeof(_native.Desc)


def test_slot_numbers_in_sync():
    from color_modem_b200 import _slots
    text = open(os.path.join(ROOT, 'color_modem_b200', 'csrc', 'cm_slots.h')).read()
    defines = dict((k, int(v)) for k, v in re.findall(r'^#define\s+([A-Z][A-Z0-9_]+)\s+(\d+)\b', text, re.M))
    names = [n for n in dir(_slots) if re.match(r'^[A-Z]{2}_', n)]
    assert len(names) > 40
    for n in names:
        assert defines[n] == getattr(_slots, n), n


def test_no_gpu_means_loud_failure():
    """There is no CPU fallback: without a CUDA device a modem cannot run."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    import numpy as np
    from color_modem_b200 import _native
    from color_modem_b200.line import LineConfig
    from color_modem_b200.color.ntsc import NtscModem
    m = NtscModem(LineConfig((720, 480)))
    with pytest.raises(_native.NativeUnavailable):
        m.encode_frames_host(np.zeros((1, 480, 720, 3), dtype=np.uint8))
    with pytest.raises(_native.NativeUnavailable):
        m.modulate(0, 0, np.zeros(720), np.zeros(720), np.zeros(720))


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'color_modem_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), f


def test_build_stamp_follows_source_content_not_file_times(lib):
    """build() decides staleness by the digest of the sources (file times do not survive the snapshot to the GPU box)."""
    import __graft_entry__ as ge
    srcs = [os.path.join(ge.CSRC, f) for f in sorted(os.listdir(ge.CSRC)) if f.endswith(('.cu', '.cuh', '.h'))]
    srcs.append(os.path.join(ROOT, 'include', 'color_modem_b200.h'))
    assert os.path.exists(ge.STAMP)
    assert open(ge.STAMP).read().strip() == ge._digest(srcs)
    newest = max(os.path.getmtime(s) for s in srcs)
    os.utime(srcs[0], (newest + 10, newest + 10))          # a newer file time alone must not trigger a rebuild
    try:
        assert ge._lib_current(srcs)
    finally:
        os.utime(srcs[0], (newest, newest))
