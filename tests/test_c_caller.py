"""A plain C program written against include/cvortex/libcvtx.h only (examples/dropin.c), compiled
with gcc and linked to libcvortex.so: the drop-in claim exercised from the language the reference's
callers use.  CPU box: host loops both times.  GPU box: CUDA path vs host loops."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    from cvortex_b200 import _native
    exe = str(tmp_path / "dropin")
    libdir = os.path.dirname(_native.LIB_PATH)
    subprocess.run(["/usr/bin/gcc", "-std=gnu99", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "dropin.c"), "-L" + libdir, "-lcvortex", "-lm",
                    "-Wl,-rpath," + libdir, "-o", exe], check=True)
    return exe


def test_c_caller_links_and_runs_on_the_host_path(product, tmp_path):
    if product.num_accelerators() > 0:
        pytest.skip("GPU present: covered by the gpu-marked test")
    res = subprocess.run([_build(tmp_path), "600"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "relative L2 difference 0.00e+00" in res.stdout          # same host loops both times
    assert "give the same particles" in res.stdout                  # and the redistribution ran (host stage)


@pytest.mark.gpu
def test_c_caller_on_the_gpu(gpu, tmp_path):
    res = subprocess.run([_build(tmp_path), "20000"], capture_output=True, text=True, timeout=300)
    print(res.stdout)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "sm_100a kernels" in res.stdout and "NVIDIA" in res.stdout
    assert "give the same particles" in res.stdout                  # CUDA node build == host stage, bit for bit
