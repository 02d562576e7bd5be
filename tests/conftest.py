import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# CWSL_MODE_STFT normally leaves groups of < 64 channels to the FAST kernel; the parity tests want the channelizer
# kernel itself on their small channel sets (the knob is read once, at the first demodulation of the process)
os.environ.setdefault("CWSL_STFT_MIN_CHANNELS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port():
    from oracle.oracle import Port
    return Port()


@pytest.fixture(scope="session")
def ref():
    """The reference's own headers compiled into oracle/_ref (prebuilt here; travels to the GPU box)."""
    from oracle.oracle import Ref
    try:
        return Ref()
    except (FileNotFoundError, OSError, Exception) as e:  # noqa: BLE001
        pytest.skip(f"oracle/_ref unavailable: {e}")


@pytest.fixture(scope="session")
def cw():
    import cwsl_digi_b200 as cw
    if not os.path.exists(cw.lib_path()):
        cw.build_library()
    return cw


def has_gpu():
    try:
        import cwsl_digi_b200 as cw
        return cw.device_count() > 0
    except Exception:  # noqa: BLE001
        return False


@pytest.fixture(scope="session")
def gpu(cw):
    if cw.device_count() <= 0:
        pytest.fail("no CUDA device: the product has no CPU path, -m gpu tests need the B200 box")
    return cw
