"""
h264-scroll-encoder_b200 - B200 (sm_100a) implementation of the h264-scroll-encoder composer's
per-frame P-slice hot path, behind the reference's C API plus the C ABI of include/composer_gpu.h.

The product is the shared library libh264scroll_b200.so (csrc/ + host/).  This Python package is
only the loader: ctypes bindings (binding.py), numpy mirrors of the ABI structs (abi.py), the
synthetic workload generators of SURVEY.md Appendix B (workloads.py) and the in-tree build (build.py).
The directory name carries a hyphen; import it through __graft_entry__.load_package(), which
registers it as `h264_scroll_encoder_b200`.
"""
from . import abi  # noqa: F401
