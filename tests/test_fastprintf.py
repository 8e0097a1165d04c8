"""The fprintf replacement of the output formatting code (smalt_b200/hostc/fastprintf.c) against libc:
same bytes for the reference's SAM / CIGAR formats, the fallback for everything else, direct appends."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fastprintf_matches_libc(tmp_path):
    exe = str(tmp_path / "fastprintf_check")
    subprocess.run(["gcc", "-O2", "-Wall", "-I", os.path.join(ROOT, "smalt_b200", "hostc"),
                    os.path.join(ROOT, "tests", "c", "fastprintf_check.c"),
                    os.path.join(ROOT, "smalt_b200", "hostc", "fastprintf.c"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == "ok"
