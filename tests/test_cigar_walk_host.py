"""The device's CIGAR walk without a GPU: csrc/cigar.cuh `cg_walk` (what cigar_task_count_kernel, gather_results and
the dense kernels call) is `__host__ __device__`; nvcc compiles it for the host here and the same source is checked
against the reference's writeDiffStrCIGAR / diffStrGetLevenshteinDistance (oracle/_ref) and against the oracle -
count pass and fill pass, all flag combinations, clips, the reference's error cases."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle_lib import ROOT, Oracle, RefLib, have_ref
from test_oracle_cigar_vs_ref import _cases

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.fixture(scope="module")
def walk(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("no nvcc")
    so = str(tmp_path_factory.mktemp("cgw") / "libcgw.so")
    src = os.path.join(ROOT, "tests", "c", "cigar_walk_host.cu")
    r = subprocess.run([NVCC, "-std=c++17", "-O1", "-shared", "-Xcompiler", "-fPIC", "-gencode",
                        "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "smalt_b200", "csrc"), src, "-o", so],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    lib = C.CDLL(so)
    lib.cgw_host.restype = C.c_int

    def f(d, cs, ce, soft, xm):
        d = bytes(d)
        out = C.create_string_buffer(6 * len(d) + 64)
        nm = C.c_int(0)
        n = lib.cgw_host(d, C.c_uint(cs), C.c_uint(ce), (2 if soft else 0) | (4 if xm else 0), out, len(out), C.byref(nm))
        assert n >= 0, n
        return out.raw[:n], nm.value
    return f


def test_device_walk_on_the_host_equals_oracle_and_reference(walk):
    orc = Oracle()
    ref = RefLib() if have_ref() else None
    rng = np.random.default_rng(91)
    n = 0
    for d in _cases(rng):
        for soft in (True, False):
            for xm in (False, True):
                cs, ce = (0, 0) if n % 3 == 0 else (int(rng.integers(0, 40)), int(rng.integers(0, 12000)))
                text, nm = walk(d, cs, ce, soft, xm)
                assert (text, nm) == orc.cigar(d, cs, ce, soft, xm), (d, cs, ce, soft, xm)
                if ref is not None:
                    e, rtext, rnm = ref.cigar(d, cs, ce, soft, xm)
                    assert e == 0 and (text, nm) == (rtext, rnm), (d, cs, ce, soft, xm)
                n += 1
    assert n > 700


def test_device_walk_error_cases(walk):
    # empty string: -1 (ERRCODE_FAILURE), no closing S byte: -59 (-ERRCODE_DIFFSTR); no text in either case
    for d, want in ((b"\0", -1), (bytes([(0 << 6) | 5, 0]), -59), (bytes([(3 << 6) | 2, (1 << 6) | 1, 0]), -59)):
        text, nm = walk(d, 3, 4, True, False)
        assert text == b"" and nm == want
