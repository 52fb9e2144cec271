"""Per-kernel parity on the B200, every call through the C ABI (ctypes -> libsfb200.so)."""
import pytest

import gpu_checks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(gpu_checks.ALL))
def test_kernel_parity(name):
    gpu_checks.ALL[name]()


def test_native_library_is_loaded():
    """The CUDA path is the one that runs: libsfb200.so must be mapped into this process."""
    from self_forcing_b200.ops import CudaOps
    CudaOps()
    maps = open("/proc/self/maps").read()
    assert "libsfb200.so" in maps
