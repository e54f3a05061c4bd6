"""pytest configuration: the `gpu` marker, oracle / library fixtures."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_available() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle_port():
    from oracle import refbind
    if not os.path.exists(refbind.PORT_SO):
        refbind.build()
    return refbind.Oracle("port")


@pytest.fixture(scope="session")
def oracle_ref():
    from oracle import refbind
    if not os.path.exists(refbind.REF_SO):
        pytest.skip("oracle/_ref/libsakura_ref.so not built (needs /root/reference)")
    return refbind.Oracle("reference")


@pytest.fixture(scope="session")
def oracle(oracle_port):
    """The strongest checker available: the reference's own object code, else the C port."""
    from oracle import refbind
    if os.path.exists(refbind.REF_SO):
        return refbind.Oracle("reference")
    return oracle_port


@pytest.fixture(scope="session")
def libpath():
    from sakura_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def lib(libpath):
    """The product library on a GPU. Fails loudly (no CPU fallback) if CUDA is missing."""
    if not _cuda_available():
        pytest.fail("gpu-marked test collected on a box without CUDA")
    from sakura_b200.capi import SakuraLib
    return SakuraLib(libpath)


@pytest.fixture(scope="session")
def hostlib(libpath):
    """The product library loaded without touching CUDA (argument validation, contexts)."""
    from sakura_b200.capi import SakuraLib
    return SakuraLib(libpath)
