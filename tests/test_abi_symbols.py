"""The C-ABI libraries load on a box without a GPU and export every function their headers
declare (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from oracle_lib import ROOT

HEADERS = {"smalt_b200.h": "libsmalt_b200.so", "smalt_b200_map.h": "libsmalt_b200_map.so"}


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(smbm?_[a-z0-9_]+)\s*\(", txt)))


@pytest.mark.parametrize("header", sorted(HEADERS))
def test_exports(header):
    path = os.path.join(ROOT, "smalt_b200", HEADERS[header])
    if not os.path.exists(path):
        pytest.skip("%s not built here" % HEADERS[header])
    C.CDLL(os.path.join(ROOT, "smalt_b200", "libsmalt_b200.so"))
    lib = C.CDLL(path)
    names = _declared(header)
    assert len(names) >= 5
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_device_fails_loudly():
    """without a CUDA device the product path refuses to run: there is no CPU fallback"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import smalt_b200
    with pytest.raises(smalt_b200.SmbError):
        smalt_b200.Context(0)
