import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")
    # the shared libraries are build artefacts (git-ignored); make sure they exist before collecting
    need = [os.path.join(ROOT, "zsc_b200", "libzsc_b200.so"), os.path.join(ROOT, "tests", "libzsc_cpuharness.so"), os.path.join(ROOT, "tests", "libzsc_cpuharness_n.so"),
            os.path.join(ROOT, "tools", "libzscgen.so"), os.path.join(ROOT, "oracle", "libzsc_oracle.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.run(["make", "-j8", "zsc_b200/libzsc_b200.so", "testlibs"], cwd=ROOT, check=True)
        subprocess.run(["make", "-C", "oracle", "all"], cwd=ROOT, check=True)


def has_gpu():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20)
        return out.returncode == 0 and "GPU" in out.stdout
    except Exception:
        return False


@pytest.fixture(scope="session")
def engine():
    from zsc_b200 import Engine
    e = Engine(raw_bytes=160 << 20, comp_bytes=200 << 20, deflate_batch_max=160 << 20, max_streams=8192, max_chunks=16384)
    yield e
    e.close()


@pytest.fixture(scope="session")
def checker():
    """The oracle: the reference's own code when oracle/_ref is present, else the CPU restatement."""
    import refimpl
    return refimpl.ref() if refimpl.have_ref() else refimpl.oracle()
