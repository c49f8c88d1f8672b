"""rayito_b200: B200-native render core behind the Rayito C++ API.

The product is two native libraries built in-tree (see build.py):
  * rayito_b200/csrc/librayito_b200.so -- hand-written sm_100a CUDA kernels + C ABI
  * rayito_b200/host/librayito_host.so -- C++ mirror of the tutorial's API (namespace Rayito)
This Python package only holds the build script and ctypes bindings used by the
tests, bench.py and tools.
"""
from . import build  # noqa: F401
