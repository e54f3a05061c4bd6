"""The driver's round-end smoke run must stay green: run it as part of the GPU suite."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_graft_entry_smoke():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import __graft_entry__
    __graft_entry__.smoke()
