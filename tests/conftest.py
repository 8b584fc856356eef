import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_unavailable():
    """Reason string when `gpu` tests cannot run here (no device / library not built), else None."""
    try:
        from fpsb200 import _lib
        if _lib.lib().fpsb_device_count() <= 0:
            return "no CUDA device visible"
    except Exception as e:
        # library missing: on a machine WITH a GPU the tests must fail loudly (no silent skip of the
        # product path); without one there is nothing they could run on
        try:
            import torch
            if torch.cuda.is_available():
                return None
        except Exception:
            pass
        return f"libfpsb200.so not loadable and no CUDA device: {e}"
    return None


def pytest_collection_modifyitems(config, items):
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    why = _gpu_unavailable()
    if why is None:
        return
    skip = pytest.mark.skip(reason=f"gpu test: {why}")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O
