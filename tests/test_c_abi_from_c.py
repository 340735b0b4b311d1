"""The C ABI called from C: tests/c/abi_smoke.c is compiled with gcc against include/zoomvit.h, linked to the in-tree
libzoomvit.so and run (host entry points only - no GPU needed)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_host_entry_points_from_a_c_program(tmp_path):
    from zoomearth_b200 import _lib
    _lib.lib()                                    # raises if the library has not been built
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / "abi_smoke")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
           "-o", exe, "-L", libdir, "-l:libzoomvit.so", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().startswith("ok"), r.stdout + r.stderr
