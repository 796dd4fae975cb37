"""pytest configuration: markers + import path.

``-m "not gpu"``: oracle vs golden vectors, host logic, C-ABI symbol checks,
world_size-2 gloo runs -- no CUDA device needed.
``-m gpu``: parity tests proper, through the C-ABI on a real B200.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
