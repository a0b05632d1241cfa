"""CPU: the C-ABI library loads and exports every symbol declared in include/sradsgan_b200.h; compute
entry points fail loudly (SR_ERR_*) without a B200 instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from sradsgan_b200 import _lib
    return _lib.load()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "sradsgan_b200.h")).read()
    hdr = re.sub(r"#ifdef SR_WITH_PROBES.*?#endif", "", hdr, flags=re.S)     # hardware probes: diagnostics builds only
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)                           # prose mentions calls such as sr_conv_pool_rows()
    declared = set(re.findall(r"\b(sr_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("sr_conv_desc")
    assert len(declared) >= 10
    for name in sorted(declared):
        assert hasattr(lib, name), "missing export %s" % name
    from sradsgan_b200 import _lib
    assert set(_lib.EXPORTS) == declared


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.sr_version() >= 100
    rc = lib.sr_device_check()
    assert rc != 0
    assert len(lib.sr_last_error()) > 0
    rc = lib.sr_colsum(None, 0, 0, 1, None, None, 0, None)
    assert rc != 0
