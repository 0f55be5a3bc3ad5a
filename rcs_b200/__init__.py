"""rcs_b200 — batched rigid-body kinematics / dynamics of HRI-EU/Rcs graphs on B200.

The product is the C library ``rcs_b200/librcsb.so`` (host C loader + hand-written sm_100a
CUDA kernels behind the C ABI of ``include/rcsb.h``). This package is a thin ctypes binding
used by the tests and ``bench.py``; it adds no computation of its own and raises if the
library is missing — there is no Python or CPU fallback for the compute path.
"""
from .api import (  # noqa: F401
    BLEND_APPROXIMATE,
    BLEND_BINARY,
    BLEND_LINEAR,
    Graph,
    Group,
    Model,
    RcsbError,
    TASK_ABC,
    TASK_COGX,
    TASK_COGY,
    TASK_COGZ,
    TASK_JOINT,
    TASK_POLARX,
    TASK_POLARY,
    TASK_POLARZ,
    TASK_XYZ,
    TASK_XYZABC,
    bind_host_to_device,
    build_library,
    kernel_launch_count,
    lib,
    library_path,
    pinned_empty,
    summary_stats,
)

__all__ = [
    "BLEND_APPROXIMATE",
    "BLEND_BINARY",
    "BLEND_LINEAR",
    "Graph",
    "Group",
    "Model",
    "RcsbError",
    "TASK_ABC",
    "TASK_COGX",
    "TASK_COGY",
    "TASK_COGZ",
    "TASK_JOINT",
    "TASK_POLARX",
    "TASK_POLARY",
    "TASK_POLARZ",
    "TASK_XYZ",
    "TASK_XYZABC",
    "bind_host_to_device",
    "build_library",
    "kernel_launch_count",
    "lib",
    "library_path",
    "pinned_empty",
    "summary_stats",
]
