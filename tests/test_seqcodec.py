"""Encoding reads while they are loaded and decoding SAM sequence fields (smalt_b200/hostc/shim_sequence.c, SSSE3
paths and table loops) against the reference's own sequence.c functions.  Needs the reference tree (the shim compiles
sequence.c in place), so it runs where /root/reference exists."""
import os
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SMALT_REF", "/root/reference")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "src", "sequence.c")), reason="reference tree not present")
def test_simd_codec_matches_reference(tmp_path):
    exe = str(tmp_path / "seqcodec_check")
    hostc = os.path.join(ROOT, "smalt_b200", "hostc")
    cmd = ["gcc", "-O2", "-w", "-DHAVE_CONFIG_H", "-I", os.path.join(hostc, "refcfg"), "-I", os.path.join(REF, "src"),
           "-I", hostc, os.path.join(ROOT, "tests", "c", "seqcodec_check.c")]
    cmd += [os.path.join(REF, "src", f) for f in ("elib.c", "filio.c")] + ["-lz", "-lm", "-o", exe]
    subprocess.run(cmd, check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok"), r.stdout
    assert "simd=1" in r.stdout      # the codec has the structure the SSSE3 paths assume
