"""ctypes binding of ``librcsb.so`` (C ABI: include/rcsb.h, include/rcsb_graph.h).

Mirrors the reference's host interface for the hot path: a ``Graph`` is an ``RcsGraph``
loaded from the reference's XML files (src/RcsCore/Rcs_graph.c:398), a ``Model`` is its
flattened device copy, and the methods of ``Model`` are the batched counterparts of
``RcsGraph_setState`` / ``RcsGraph_bodyPointJacobian`` / ``RcsGraph_rotationJacobian`` /
``RcsGraph_computeKineticTerms`` / ``IkSolverRMR::solveRightInverse``.

Array arguments may be torch CUDA tensors (device pointers, asynchronous on the current
torch stream) or numpy arrays (host pointers: the library copies in / out itself).
PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes
import json
import os
import subprocess
from typing import Dict, Iterable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TASK_XYZ = 1
TASK_ABC = 2
TASK_XYZABC = 3
_HOST_PTRS = 1
_IK_LEFT_INVERSE = 2


class RcsbError(RuntimeError):
    """Raised when a librcsb call returns a non-zero status."""


def library_path() -> str:
    # RCSB_LIBRARY: an alternative build of the same library (kernel tuning experiments)
    return os.environ.get("RCSB_LIBRARY") or os.path.join(_HERE, "librcsb.so")


def build_library(verbose: bool = False) -> str:
    """Compile librcsb.so in-tree (nvcc, sm_100a). Used by __graft_entry__.build()."""
    res = subprocess.run(
        ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"],
        stdout=subprocess.PIPE,
        stderr=subprocess.STDOUT,
        text=True,
    )
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building librcsb.so failed")
    return library_path()


class _RcsbRelJac(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("effector", ctypes.c_int), ("refBdy", ctypes.c_int),
                ("refFrame", ctypes.c_int)]


class _RcsbTask(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("effector", ctypes.c_int)]


class _RcsbIkOptions(ctypes.Structure):
    _fields_ = [("activation", ctypes.c_void_p), ("scaleDx", ctypes.c_int), ("blending", ctypes.c_int),
                ("lambdaRows", ctypes.c_void_p), ("dqTs", ctypes.c_void_p), ("dqNs", ctypes.c_void_p)]


BLEND_BINARY, BLEND_LINEAR, BLEND_APPROXIMATE = 0, 1, 2
TASK_COGX, TASK_COGY, TASK_COGZ, TASK_JOINT = 4, 5, 6, 7
TASK_POLARX, TASK_POLARY, TASK_POLARZ = 8, 9, 10


class _RcsbTaskSpec(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("effector", ctypes.c_int), ("refBdy", ctypes.c_int),
                ("refFrame", ctypes.c_int), ("refGain", ctypes.c_double)]

    @classmethod
    def of(cls, t):
        """(kind, effector[, refBdy[, refFrame[, refGain]]]); refGain defaults to the reference's 1.0"""
        t = tuple(t)
        ints = [int(v) for v in t[:4]] + [-1] * (4 - min(len(t), 4))
        ints[0], ints[1] = int(t[0]), int(t[1])
        return cls(ints[0], ints[1], ints[2], ints[3], float(t[4]) if len(t) > 4 else 1.0)


def lib() -> ctypes.CDLL:
    """Loads librcsb.so once. Fails loudly if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RcsbError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (rcs_b200 has no fallback implementation)"
        )
    L = ctypes.CDLL(path)
    c_void, c_int, c_long, c_dbl_p = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_void_p

    def sig(name, restype, *argtypes):
        fn = getattr(L, name)
        fn.restype = restype
        fn.argtypes = list(argtypes)

    sig("RcsGraph_create", c_void, ctypes.c_char_p)
    sig("RcsGraph_fromURDFFile", c_void, ctypes.c_char_p)
    sig("RcsGraph_createFromBuffer", c_void, ctypes.c_char_p, ctypes.c_uint)
    sig("RcsGraph_destroy", None, c_void)
    sig("RcsGraph_clone", c_void, c_void)
    sig("RcsGraph_check", ctypes.c_bool, c_void, ctypes.POINTER(c_int), ctypes.POINTER(c_int))
    sig("RcsGraph_layoutJSON", c_void, c_void)
    sig("RcsGraph_writeXML", c_int, c_void, ctypes.c_char_p)
    sig("RcsGraph_numBodies", ctypes.c_uint, c_void)
    sig("RcsGraph_getBodyByName", c_void, c_void, ctypes.c_char_p)
    sig("RcsGraph_bodyIndex", c_int, c_void, c_void)
    sig("RcsGraph_mass", ctypes.c_double, c_void)
    sig("Rcs_addResourcePath", ctypes.c_bool, ctypes.c_char_p)
    sig("Rcs_clearResourcePath", None)
    sig("Rcs_setFatalMode", None, c_int)
    sig("rcsb_last_error", ctypes.c_char_p)
    sig("rcsb_device_count", c_int)
    sig("rcsb_host_alloc", c_void, ctypes.c_size_t)
    sig("rcsb_host_free", None, c_void)
    sig("rcsb_kernel_launch_count", ctypes.c_ulonglong)
    sig("rcsb_model_create", c_int, c_void, c_int, ctypes.POINTER(c_void))
    sig("rcsb_model_destroy", None, c_void)
    sig("rcsb_model_dof", c_int, c_void)
    sig("rcsb_model_nj", c_int, c_void)
    sig("rcsb_model_nbodies", c_int, c_void)
    sig("rcsb_fk", c_int, c_void, c_long, c_int, c_dbl_p, c_dbl_p, c_int, c_void, c_dbl_p,
        c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, ctypes.c_uint, c_void)
    sig("rcsb_fk_jacobians", c_int, c_void, c_long, c_int, c_dbl_p, c_int, c_void, c_dbl_p,
        c_int, c_void, c_void, c_dbl_p, c_dbl_p, ctypes.c_uint, c_void)
    sig("rcsb_fk_jacobians_mass", c_int, c_void, c_long, c_int, c_dbl_p, c_int, c_void, c_dbl_p,
        c_int, c_void, c_void, c_dbl_p, c_dbl_p, c_dbl_p, ctypes.c_uint, c_void)
    sig("rcsb_kinetic_terms", c_int, c_void, c_long, c_int, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p,
        c_dbl_p, c_dbl_p, ctypes.c_uint, c_void)
    sig("rcsb_task_jacobians", c_int, c_void, c_long, c_int, c_dbl_p, c_int, c_void, c_dbl_p,
        ctypes.c_uint, c_void)
    sig("rcsb_direct_dynamics", c_int, c_void, c_long, c_int, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p,
        c_dbl_p, c_dbl_p, ctypes.c_uint, c_void)
    sig("rcsb_cog", c_int, c_void, c_long, c_int, c_dbl_p, c_dbl_p, c_dbl_p, ctypes.c_uint, c_void)
    sig("rcsb_kinetic_terms_cog", c_int, c_void, c_long, c_int, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p,
        c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, ctypes.c_uint, c_void)
    sig("rcsb_ik_rmr", c_int, c_void, c_int, c_void, c_long, c_dbl_p, c_dbl_p, ctypes.c_double,
        ctypes.c_double, c_int, c_dbl_p, c_dbl_p, c_dbl_p, ctypes.c_uint, c_void)
    sig("rcsb_ik_rmr_ex", c_int, c_void, c_int, c_void, c_long, c_dbl_p, c_dbl_p, ctypes.c_double,
        ctypes.c_double, c_int, c_dbl_p, c_dbl_p, c_dbl_p, c_void, ctypes.c_uint, c_void)
    sig("rcsb_ik_prio_rmr", c_int, c_void, c_int, c_void, c_void, c_void, c_long, c_dbl_p, c_dbl_p,
        ctypes.c_double, ctypes.c_double, c_int, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, ctypes.c_uint, c_void)
    sig("rcsb_ik_tasks_rmr", c_int, c_void, c_int, c_void, c_long, c_dbl_p, c_dbl_p, ctypes.c_double,
        ctypes.c_double, c_int, c_dbl_p, c_dbl_p, c_dbl_p, ctypes.c_uint, c_void)
    sig("rcsb_ik_constraint_rmr", c_int, c_void, c_int, c_void, c_long, c_dbl_p, c_dbl_p, ctypes.c_double,
        ctypes.c_double, c_int, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, ctypes.c_uint, c_void)
    sig("rcsb_summary_stats", c_int, c_dbl_p, c_long, c_long, c_dbl_p, c_void)
    sig("rcsb_group_create", c_int, c_void, c_int, c_void, ctypes.POINTER(c_void))
    sig("rcsb_group_destroy", None, c_void)
    sig("rcsb_group_size", c_int, c_void)
    sig("rcsb_group_slice", None, c_void, c_long, c_int, ctypes.POINTER(c_long), ctypes.POINTER(c_long))
    sig("rcsb_group_fk_jacobians_mass", c_int, c_void, c_long, c_int, c_dbl_p, c_int, c_void, c_dbl_p, c_int,
        c_void, c_void, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p)
    sig("rcsb_group_kinetic_terms", c_int, c_void, c_long, c_int, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p,
        c_dbl_p, c_dbl_p)
    sig("rcsb_group_ik_rmr", c_int, c_void, c_int, c_void, c_long, c_dbl_p, c_dbl_p, ctypes.c_double,
        ctypes.c_double, c_int, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p)
    sig("rcsb_group_wholebody", c_int, c_void, c_long, c_int, c_dbl_p, c_int, c_void, c_dbl_p, c_int, c_void,
        c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p)
    sig("rcsb_wholebody", c_int, c_void, c_long, c_int, c_dbl_p, c_int, c_void, c_dbl_p, c_int, c_void,
        c_dbl_p, c_dbl_p, c_dbl_p, ctypes.c_uint, c_void)
    sig("rcsb_bind_host_to_device", c_int, c_int, ctypes.c_char_p, ctypes.c_size_t)
    sig("RcsGraph_setState", ctypes.c_bool, c_void, c_void, c_void)
    sig("MatNd_create", c_void, ctypes.c_uint, ctypes.c_uint)
    sig("MatNd_destroy", None, c_void)
    # Contract violations return errors instead of abort()ing the interpreter.
    L.Rcs_setFatalMode(0)
    _LIB = L
    return L


def kernel_launch_count() -> int:
    return int(lib().rcsb_kernel_launch_count())


def bind_host_to_device(device: int) -> str:
    """rcsb_bind_host_to_device: run this process's host threads and place its pinned buffers on
    the NUMA node of GPU `device` (as far as the cpuset allows). Returns what was done."""
    buf = ctypes.create_string_buffer(512)
    _err(lib().rcsb_bind_host_to_device(int(device), buf, len(buf)), "rcsb_bind_host_to_device")
    return buf.value.decode()


def pinned_empty(shape):
    """Page-locked float64 host tensor for RCSB_HOST_PTRS calls (torch owns the memory)."""
    import torch

    return torch.empty(shape, dtype=torch.float64, pin_memory=True)


def summary_stats(x, out=None):
    """[count, sum, sum of squares, min, max] of a (possibly strided) 1-D float64 CUDA tensor,
    computed by one librcsb kernel on the current torch stream (rcsb_summary_stats)."""
    import torch

    if not (x.is_cuda and x.dtype == torch.float64 and x.dim() == 1):
        raise RcsbError("summary_stats: need a 1-D float64 CUDA tensor")
    if out is None:
        out = torch.empty(5, dtype=torch.float64, device=x.device)
    stride = x.stride(0) if x.numel() > 1 else 1
    if stride < 1:
        raise RcsbError("summary_stats: need a positive stride")
    with torch.cuda.device(x.device):
        rc = lib().rcsb_summary_stats(ctypes.c_void_p(x.data_ptr()), x.numel(), stride,
                                      ctypes.c_void_p(out.data_ptr()),
                                      ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
    _err(rc, "rcsb_summary_stats")
    return out


def _err(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().rcsb_last_error()
        raise RcsbError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


class Graph:
    """An RcsGraph loaded by librcsb's own loaders (host only, no GPU needed): RcsGraph_create, or
    RcsGraph_fromURDFFile for a file name ending in .urdf."""

    def __init__(self, source: str, _handle: Optional[int] = None):
        L = lib()
        if _handle is not None:
            self._h = _handle
        else:
            urdf = source.lower().endswith(".urdf")
            self._h = (L.RcsGraph_fromURDFFile if urdf else L.RcsGraph_create)(source.encode())
        if not self._h:
            msg = L.rcsb_last_error()
            raise RcsbError(f"RcsGraph_create({source!r}) failed: {msg.decode() if msg else ''}")
        self._layout = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _LIB is not None:
            _LIB.RcsGraph_destroy(h)
            self._h = None

    @property
    def handle(self) -> int:
        return self._h

    def layout(self) -> Dict:
        if self._layout is None:
            L = lib()
            p = L.RcsGraph_layoutJSON(self._h)
            self._layout = json.loads(ctypes.string_at(p).decode())
            ctypes.CDLL(None).free(ctypes.c_void_p(p))
        return self._layout

    @property
    def dof(self) -> int:
        return self.layout()["dof"]

    @property
    def nJ(self) -> int:
        return self.layout()["nJ"]

    @property
    def n_bodies(self) -> int:
        return self.layout()["nBodies"]

    def body_names(self):
        return [b["name"] for b in self.layout()["bodies"]]

    def body_index(self, name: str) -> int:
        L = lib()
        b = L.RcsGraph_getBodyByName(self._h, name.encode())
        if not b:
            raise KeyError(name)
        return L.RcsGraph_bodyIndex(self._h, b)

    def clone(self) -> "Graph":
        return Graph("", _handle=lib().RcsGraph_clone(self._h))

    def check(self):
        ne, nw = ctypes.c_int(0), ctypes.c_int(0)
        ok = lib().RcsGraph_check(self._h, ctypes.byref(ne), ctypes.byref(nw))
        return bool(ok), ne.value, nw.value

    def mass(self) -> float:
        return float(lib().RcsGraph_mass(self._h))

    def write_xml(self, path: str) -> None:
        if lib().RcsGraph_writeXML(self._h, path.encode()) != 0:
            raise RcsbError(f"RcsGraph_writeXML({path!r}) failed")


def _is_torch(x) -> bool:
    if x is None or not type(x).__module__.startswith("torch"):
        return False
    import torch

    return isinstance(x, torch.Tensor)


class _Args:
    """Collects pointers for one call and decides host vs device mode."""

    def __init__(self, model: "Model"):
        self.model = model
        self.mode = None  # "device" | "host"
        self.keep = []

    def ptr(self, x, name: str):
        if x is None:
            return None
        if _is_torch(x):
            import torch

            if x.dtype != torch.float64 or not x.is_contiguous():
                raise RcsbError(f"{name}: need a contiguous float64 tensor")
            if x.is_cuda:
                if x.device.index != self.model.device:
                    raise RcsbError(f"{name}: tensor is on {x.device}, model on cuda:{self.model.device}")
                self._set("device", name)
            else:
                self._set("host", name)
            self.keep.append(x)
            return ctypes.c_void_p(x.data_ptr())
        a = x
        if not isinstance(a, np.ndarray) or a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
            raise RcsbError(f"{name}: need a C-contiguous float64 numpy array")
        self._set("host", name)
        self.keep.append(a)
        return ctypes.c_void_p(a.ctypes.data)

    def _set(self, mode: str, name: str):
        if self.mode is None:
            self.mode = mode
        elif self.mode != mode:
            raise RcsbError(f"{name}: host and device arrays mixed in one call")

    def flags(self) -> int:
        return _HOST_PTRS if self.mode == "host" else 0

    def stream(self):
        if self.mode == "device":
            import torch

            return ctypes.c_void_p(torch.cuda.current_stream(self.model.device).cuda_stream)
        return None


def _int_array(values: Optional[Iterable[int]]):
    if values is None:
        return None, 0
    v = list(values)
    arr = (ctypes.c_int * max(len(v), 1))(*v)
    return arr, len(v)


class Model:
    """Flattened, device-resident copy of a Graph (immutable)."""

    def __init__(self, graph: Graph, device: int = 0):
        L = lib()
        h = ctypes.c_void_p()
        _err(L.rcsb_model_create(graph.handle, device, ctypes.byref(h)), "rcsb_model_create")
        self._h = h
        self.device = device
        self.dof = L.rcsb_model_dof(h)
        self.nJ = L.rcsb_model_nj(h)
        self.n_bodies = L.rcsb_model_nbodies(h)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _LIB is not None:
            _LIB.rcsb_model_destroy(h)
            self._h = None

    # -- helpers -------------------------------------------------------------------------
    def _alloc(self, like, shape):
        if _is_torch(like):
            import torch

            return torch.empty(shape, dtype=torch.float64, device=like.device)
        return np.empty(shape, dtype=np.float64)

    @staticmethod
    def _check(what, arrays, shapes):
        """Caller-supplied arrays must have exactly the shape the C call will read / write: the C
        ABI trusts the sizes, a short buffer would be an out-of-bounds access on the device."""
        for k, v in arrays.items():
            if v is not None and k in shapes and tuple(v.shape) != tuple(shapes[k]):
                raise RcsbError(f"{what}: {k} has shape {tuple(v.shape)}, expected {tuple(shapes[k])}")

    @staticmethod
    def _batch(q):
        if q.ndim != 2:
            raise RcsbError("q must be N x qdim")
        return int(q.shape[0]), int(q.shape[1])

    # -- forward kinematics --------------------------------------------------------------
    def fk(self, q, qd=None, bodies: Optional[Sequence[int]] = None, want=("A_BI",), out=None):
        """RcsGraph_setState over a batch. want ⊆ {A_BI, A_JI, xdot, omega, q}."""
        N, qdim = self._batch(q)
        nsel = self.n_bodies if bodies is None else len(bodies)
        shapes = {
            "A_BI": (N, nsel, 12),
            "A_JI": (N, self.dof, 12),
            "xdot": (N, nsel, 3),
            "omega": (N, nsel, 3),
            "q": (N, self.dof),
        }
        out = dict(out or {})
        for k in want:
            if k not in out:
                out[k] = self._alloc(q, shapes[k])
        self._check("fk", dict(out, qd=qd), dict(shapes, qd=(N, qdim)))
        a = _Args(self)
        sel, ns = _int_array(bodies)
        rc = lib().rcsb_fk(
            self._h, N, qdim, a.ptr(q, "q"), a.ptr(qd, "qd"), ns, sel,
            a.ptr(out.get("A_BI"), "A_BI"), a.ptr(out.get("A_JI"), "A_JI"),
            a.ptr(out.get("xdot"), "xdot"), a.ptr(out.get("omega"), "omega"),
            a.ptr(out.get("q"), "q_out"), a.flags(), a.stream(),
        )
        _err(rc, "rcsb_fk")
        return out

    def fk_jacobians(self, q, effectors: Sequence[int], points=None,
                     bodies: Optional[Sequence[int]] = None, want=("A_BI", "JT", "JR"), out=None):
        """FK + RcsGraph_bodyPointJacobian / RcsGraph_rotationJacobian per effector; with "M" in
        `want` also the mass matrix from the same launch (rcsb_fk_jacobians_mass)."""
        N, qdim = self._batch(q)
        nsel = self.n_bodies if bodies is None else len(bodies)
        neff = len(effectors)
        shapes = {"A_BI": (N, nsel, 12), "JT": (N, neff, 3, self.nJ), "JR": (N, neff, 3, self.nJ),
                  "M": (N, self.nJ, self.nJ)}
        out = dict(out or {})
        for k in want:
            if k not in out:
                out[k] = self._alloc(q, shapes[k])
        self._check("fk_jacobians", out, shapes)
        a = _Args(self)
        sel, ns = _int_array(bodies)
        eff, ne = _int_array(effectors)
        pts = None
        if points is not None:
            pa = np.ascontiguousarray(np.asarray(points, dtype=np.float64).reshape(neff, 3))
            a.keep.append(pa)
            pts = ctypes.c_void_p(pa.ctypes.data)
        rc = lib().rcsb_fk_jacobians_mass(
            self._h, N, qdim, a.ptr(q, "q"), ns, sel, a.ptr(out.get("A_BI"), "A_BI"), ne, eff, pts,
            a.ptr(out.get("JT"), "JT"), a.ptr(out.get("JR"), "JR"), a.ptr(out.get("M"), "M"),
            a.flags(), a.stream(),
        )
        _err(rc, "rcsb_fk_jacobians_mass")
        return out

    def task_jacobians(self, q, spec, out=None):
        """RcsGraph_3dPosJacobian / RcsGraph_3dOmegaJacobian over a batch.

        spec: sequence of ("POS" | "OMEGA", effector, refBdy, refFrame), body indices, -1 = NULL.
        Returns J [N x len(spec) x 3 x nJ]."""
        N, qdim = self._batch(q)
        if out is None:
            out = self._alloc(q, (N, len(spec), 3, self.nJ))
        arr = (_RcsbRelJac * max(len(spec), 1))(
            *[_RcsbRelJac(1 if k == "POS" else 2, int(e), int(rb), int(rf)) for k, e, rb, rf in spec])
        a = _Args(self)
        rc = lib().rcsb_task_jacobians(self._h, N, qdim, a.ptr(q, "q"), len(spec), arr,
                                       a.ptr(out, "J"), a.flags(), a.stream())
        _err(rc, "rcsb_task_jacobians")
        return out

    def wholebody(self, q, spec, bodies: Optional[Sequence[int]] = None,
                  want=("A_BI", "Jt", "M", "Jcog"), out=None):
        """rcsb_wholebody: frames, task Jacobians (spec as in task_jacobians), M and the COG
        Jacobian of the same states from one call."""
        N, qdim = self._batch(q)
        nsel = self.n_bodies if bodies is None else len(bodies)
        shapes = {"A_BI": (N, nsel, 12), "Jt": (N, len(spec), 3, self.nJ), "M": (N, self.nJ, self.nJ),
                  "Jcog": (N, 3, self.nJ)}
        out = dict(out or {})
        for k in want:
            if k not in out:
                out[k] = self._alloc(q, shapes[k])
        for k, v in out.items():
            if tuple(v.shape) != shapes[k]:
                raise RcsbError(f"wholebody: {k} has shape {tuple(v.shape)}, expected {shapes[k]}")
        arr = (_RcsbRelJac * max(len(spec), 1))(
            *[_RcsbRelJac(1 if k == "POS" else 2, int(e), int(rb), int(rf)) for k, e, rb, rf in spec])
        sel, ns = _int_array(bodies)
        a = _Args(self)
        has_j = out.get("Jt") is not None
        rc = lib().rcsb_wholebody(self._h, N, qdim, a.ptr(q, "q"), ns, sel, a.ptr(out.get("A_BI"), "A_BI"),
                                  len(spec) if has_j else 0, arr if has_j else None, a.ptr(out.get("Jt"), "Jt"),
                                  a.ptr(out.get("M"), "M"), a.ptr(out.get("Jcog"), "Jcog"), a.flags(),
                                  a.stream())
        _err(rc, "rcsb_wholebody")
        return out

    # -- dynamics ------------------------------------------------------------------------
    def kinetic_terms(self, q, qd=None, want=("M", "h", "Fg", "E"), out=None):
        """RcsGraph_computeKineticTerms over a batch; with "cog" / "Jcog" in `want` also
        RcsGraph_COG / RcsGraph_COGJacobian from the same pass (rcsb_kinetic_terms_cog)."""
        N, qdim = self._batch(q)
        shapes = {"M": (N, self.nJ, self.nJ), "h": (N, self.nJ), "Fg": (N, self.nJ), "E": (N,),
                  "cog": (N, 3), "Jcog": (N, 3, self.nJ)}
        out = dict(out or {})
        for k in want:
            if k not in out:
                out[k] = self._alloc(q, shapes[k])
        self._check("kinetic_terms", dict(out, qd=qd), dict(shapes, qd=(N, qdim)))
        a = _Args(self)
        if "cog" in out or "Jcog" in out:
            rc = lib().rcsb_kinetic_terms_cog(
                self._h, N, qdim, a.ptr(q, "q"), a.ptr(qd, "qd"), a.ptr(out.get("M"), "M"),
                a.ptr(out.get("h"), "h"), a.ptr(out.get("Fg"), "Fg"), a.ptr(out.get("E"), "E"),
                a.ptr(out.get("cog"), "cog"), a.ptr(out.get("Jcog"), "Jcog"), a.flags(), a.stream(),
            )
            _err(rc, "rcsb_kinetic_terms_cog")
            return out
        rc = lib().rcsb_kinetic_terms(
            self._h, N, qdim, a.ptr(q, "q"), a.ptr(qd, "qd"), a.ptr(out.get("M"), "M"),
            a.ptr(out.get("h"), "h"), a.ptr(out.get("Fg"), "Fg"), a.ptr(out.get("E"), "E"),
            a.flags(), a.stream(),
        )
        _err(rc, "rcsb_kinetic_terms")
        return out

    def direct_dynamics(self, q, qd, F_ext=None, F_jnt=None, want=("qdd", "E"), out=None):
        """Rcs_directDynamics over a batch: M qdd = F_gravity + F_ext - F_jnt + h."""
        N, qdim = self._batch(q)
        shapes = {"qdd": (N, self.nJ), "E": (N,)}
        out = dict(out or {})
        for k in set(want) | {"qdd"}:
            if k not in out:
                out[k] = self._alloc(q, shapes[k])
        self._check("direct_dynamics", dict(out, qd=qd, F_ext=F_ext, F_jnt=F_jnt),
                    dict(shapes, qd=(N, qdim), F_ext=(N, self.nJ), F_jnt=(N, self.nJ)))
        a = _Args(self)
        rc = lib().rcsb_direct_dynamics(
            self._h, N, qdim, a.ptr(q, "q"), a.ptr(qd, "qd"), a.ptr(F_ext, "F_ext"),
            a.ptr(F_jnt, "F_jnt"), a.ptr(out["qdd"], "qdd"), a.ptr(out.get("E"), "E"), a.flags(),
            a.stream())
        _err(rc, "rcsb_direct_dynamics")
        return out

    def cog(self, q, want=("cog", "Jcog"), out=None):
        """RcsGraph_COG / RcsGraph_COGJacobian over a batch."""
        N, qdim = self._batch(q)
        shapes = {"cog": (N, 3), "Jcog": (N, 3, self.nJ)}
        out = dict(out or {})
        for k in want:
            if k not in out:
                out[k] = self._alloc(q, shapes[k])
        self._check("cog", out, shapes)
        a = _Args(self)
        rc = lib().rcsb_cog(self._h, N, qdim, a.ptr(q, "q"), a.ptr(out.get("cog"), "cog"),
                            a.ptr(out.get("Jcog"), "Jcog"), a.flags(), a.stream())
        _err(rc, "rcsb_cog")
        return out

    # -- inverse kinematics --------------------------------------------------------------
    def ik_rmr_ex(self, tasks: Sequence, q, x_des, activation=None, scale_dx=False, lambda_rows=None,
                  blending=BLEND_BINARY, lam=1.0e-8, alpha=0.05, iters=1,
                  want=("det", "dq_ts", "dq_ns"), out=None, left_inverse=False):
        """rcsb_ik_rmr_ex: the solver with task activations, per-row regularisation, task-space
        blending (left inverse) and the task-space / null-space parts of the step. x_des keeps the
        full layout over ALL tasks; "dx" holds the active rows only."""
        N, qdim = self._batch(q)
        if qdim != self.dof:
            raise RcsbError("ik_rmr_ex: q must be N x dof")
        dims = [6 if int(k) == TASK_XYZABC else 3 for k, _ in tasks]
        act = np.ones(len(tasks)) if activation is None else np.ascontiguousarray(activation, dtype=np.float64)
        if act.shape != (len(tasks),):
            raise RcsbError("ik_rmr_ex: one activation per task")
        nx_all = sum(dims)
        nx_act = sum(d for d, a_ in zip(dims, act) if a_ > 0.0)
        if tuple(x_des.shape) != (N, nx_all):
            raise RcsbError(f"ik_rmr_ex: x_des must be {N} x {nx_all}")
        lr = None
        if lambda_rows is not None:
            lr = np.ascontiguousarray(lambda_rows, dtype=np.float64)
            if lr.shape != (nx_act,):
                raise RcsbError(f"ik_rmr_ex: lambda_rows needs {nx_act} entries (active rows)")
        shapes = {"dx": (N, nx_act), "dq": (N, self.dof), "det": (N,), "dq_ts": (N, self.dof),
                  "dq_ns": (N, self.dof)}
        out = dict(out or {})
        for k in want:
            if k not in out:
                out[k] = self._alloc(q, shapes[k])
        tarr = (_RcsbTask * max(len(tasks), 1))(*[_RcsbTask(int(k), int(e)) for k, e in tasks])
        a = _Args(self)
        opt = _RcsbIkOptions(act.ctypes.data, int(bool(scale_dx)), int(blending),
                             None if lr is None else lr.ctypes.data,
                             a.ptr(out.get("dq_ts"), "dq_ts"), a.ptr(out.get("dq_ns"), "dq_ns"))
        rc = lib().rcsb_ik_rmr_ex(
            self._h, len(tasks), tarr, N, a.ptr(q, "q"), a.ptr(x_des, "x_des"), float(lam),
            float(alpha), int(iters), a.ptr(out.get("dx"), "dx"), a.ptr(out.get("dq"), "dq"),
            a.ptr(out.get("det"), "det"), ctypes.byref(opt),
            a.flags() | (_IK_LEFT_INVERSE if left_inverse else 0), a.stream())
        _err(rc, "rcsb_ik_rmr_ex")
        return out

    def ik_tasks_rmr(self, tasks: Sequence, q, x_des, lam=1.0e-8, alpha=0.05, iters=1, want=("det",), out=None,
                     constraint=False):
        """rcsb_ik_tasks_rmr: RMR steps over any stack of task kinds with reference bodies / frames.
        tasks: sequence of (kind, effector[, refBdy[, refFrame]]), -1 = none. constraint=True:
        rcsb_ik_constraint_rmr (IkSolverConstraintRMR's joint-limit active set); "n_violations" may then
        be asked for."""
        N, qdim = self._batch(q)
        who = "ik_constraint_rmr" if constraint else "ik_tasks_rmr"
        if qdim != self.dof:
            raise RcsbError(f"{who}: q must be N x dof")
        spec = [tuple(t) + (-1,) * (4 - len(t)) for t in tasks]
        dims = {TASK_XYZ: 3, TASK_ABC: 3, TASK_XYZABC: 6, TASK_COGX: 1, TASK_COGY: 1, TASK_COGZ: 1, TASK_JOINT: 1,
                TASK_POLARX: 2, TASK_POLARY: 2, TASK_POLARZ: 2}
        nx = sum(dims.get(int(t[0]), 0) for t in spec)
        if tuple(x_des.shape) != (N, nx):
            raise RcsbError(f"{who}: x_des must be {N} x {nx}")
        shapes = {"dx": (N, nx), "dq": (N, self.dof), "det": (N,), "n_violations": (N,)}
        out = dict(out or {})
        for k in want:
            if k not in out:
                out[k] = self._alloc(q, shapes[k])
        self._check(who, out, shapes)
        tarr = (_RcsbTaskSpec * max(len(spec), 1))(*[_RcsbTaskSpec.of(t) for t in spec])
        a = _Args(self)
        common = (self._h, len(spec), tarr, N, a.ptr(q, "q"), a.ptr(x_des, "x_des"), float(lam), float(alpha),
                  int(iters), a.ptr(out.get("dx"), "dx"), a.ptr(out.get("dq"), "dq"), a.ptr(out.get("det"), "det"))
        if constraint:
            rc = lib().rcsb_ik_constraint_rmr(*common, a.ptr(out.get("n_violations"), "n_violations"), a.flags(),
                                              a.stream())
        else:
            if "n_violations" in out:
                raise RcsbError("ik_tasks_rmr: n_violations needs constraint=True")
            rc = lib().rcsb_ik_tasks_rmr(*common, a.flags(), a.stream())
        _err(rc, "rcsb_" + who)
        return out

    def ik_prio_rmr(self, tasks: Sequence, priority: Sequence[int], q, x_des, activation=None, lam=1.0e-8,
                    alpha=0.05, iters=1, want=("dq1", "dq2", "dq_ns", "ok"), out=None):
        """rcsb_ik_prio_rmr: IkSolverPrioRMR::solve steps with two priority levels + null space."""
        N, qdim = self._batch(q)
        if qdim != self.dof:
            raise RcsbError("ik_prio_rmr: q must be N x dof")
        if len(priority) != len(tasks):
            raise RcsbError("ik_prio_rmr: one priority level per task")
        nx_all = sum(6 if int(k) == TASK_XYZABC else 3 for k, _ in tasks)
        if tuple(x_des.shape) != (N, nx_all):
            raise RcsbError(f"ik_prio_rmr: x_des must be {N} x {nx_all}")
        shapes = {"dq1": (N, self.dof), "dq2": (N, self.dof), "dq_ns": (N, self.dof), "ok": (N,)}
        out = dict(out or {})
        for k in want:
            if k not in out:
                out[k] = self._alloc(q, shapes[k])
        tarr = (_RcsbTask * max(len(tasks), 1))(*[_RcsbTask(int(k), int(e)) for k, e in tasks])
        prio = (ctypes.c_int * len(priority))(*[int(p) for p in priority])
        act = None
        if activation is not None:
            act_np = np.ascontiguousarray(activation, dtype=np.float64)
            if act_np.shape != (len(tasks),):
                raise RcsbError("ik_prio_rmr: one activation per task")
            act = ctypes.c_void_p(act_np.ctypes.data)
        a = _Args(self)
        rc = lib().rcsb_ik_prio_rmr(
            self._h, len(tasks), tarr, prio, act, N, a.ptr(q, "q"), a.ptr(x_des, "x_des"), float(lam),
            float(alpha), int(iters), a.ptr(out.get("dq1"), "dq1"), a.ptr(out.get("dq2"), "dq2"),
            a.ptr(out.get("dq_ns"), "dq_ns"), a.ptr(out.get("ok"), "ok"), a.flags(), a.stream())
        _err(rc, "rcsb_ik_prio_rmr")
        return out

    def ik_rmr(self, tasks: Sequence, q, x_des, lam=1.0e-8, alpha=0.05, iters=1,
               want=("det",), out=None, left_inverse=False):
        """`iters` IkSolverRMR::solveRightInverse steps per state (solveLeftInverse with
        `left_inverse`); q is updated in place.

        tasks: sequence of (TASK_XYZ | TASK_ABC | TASK_XYZABC, effector body index).
        """
        N, qdim = self._batch(q)
        if qdim != self.dof:
            raise RcsbError("ik_rmr: q must be N x dof")
        nx = sum(6 if int(k) == TASK_XYZABC else 3 for k, _ in tasks)
        shapes = {"dx": (N, nx), "dq": (N, self.dof), "det": (N,)}
        out = dict(out or {})
        for k in want:
            if k not in out:
                out[k] = self._alloc(q, shapes[k])
        self._check("ik_rmr", dict(out, x_des=x_des), dict(shapes, x_des=(N, nx)))
        tarr = (_RcsbTask * max(len(tasks), 1))(*[_RcsbTask(int(k), int(e)) for k, e in tasks])
        a = _Args(self)
        rc = lib().rcsb_ik_rmr(
            self._h, len(tasks), tarr, N, a.ptr(q, "q"), a.ptr(x_des, "x_des"), float(lam),
            float(alpha), int(iters), a.ptr(out.get("dx"), "dx"), a.ptr(out.get("dq"), "dq"),
            a.ptr(out.get("det"), "det"), a.flags() | (_IK_LEFT_INVERSE if left_inverse else 0), a.stream(),
        )
        _err(rc, "rcsb_ik_rmr")
        return out


def _host_ptr(x, name):
    """Pointer of a host array of a group call: C-contiguous float64 numpy array or CPU tensor."""
    if x is None:
        return None
    if _is_torch(x):
        import torch

        if x.is_cuda or x.dtype != torch.float64 or not x.is_contiguous():
            raise RcsbError(f"{name}: need a contiguous float64 host tensor")
        return ctypes.c_void_p(x.data_ptr())
    if not isinstance(x, np.ndarray) or x.dtype != np.float64 or not x.flags["C_CONTIGUOUS"]:
        raise RcsbError(f"{name}: need a C-contiguous float64 numpy array")
    return ctypes.c_void_p(x.ctypes.data)


class Group:
    """rcsb_group_*: the GPUs of one node behind one handle (single process). Host arrays of the
    whole batch in, results and the gathered per-rank statistics [n_devices x 5] out."""

    def __init__(self, graph: Graph, devices: Sequence[int]):
        devs = list(devices)
        if len(devs) > 1:
            # librcsb picks up the NCCL copy that is already in the process: let it be PyTorch's (a
            # different libnccl.so.2 loaded first would break a later `import torch`)
            try:
                import torch  # noqa: F401
            except ImportError:
                pass
        arr = (ctypes.c_int * len(devs))(*devs)
        h = ctypes.c_void_p()
        _err(lib().rcsb_group_create(graph.handle, len(devs), arr, ctypes.byref(h)), "rcsb_group_create")
        self._h = h
        self.devices = devs
        self.dof, self.nJ, self.n_bodies = graph.dof, graph.nJ, graph.n_bodies

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _LIB is not None:
            _LIB.rcsb_group_destroy(h)
            self._h = None

    def slice_of(self, rank: int, n: int):
        first, count = ctypes.c_long(), ctypes.c_long()
        lib().rcsb_group_slice(self._h, n, rank, ctypes.byref(first), ctypes.byref(count))
        return first.value, count.value

    def _stats(self):
        return np.empty((len(self.devices), 5))

    def fk_jacobians(self, q, effectors: Sequence[int], out: Dict):
        """out: dict of host arrays among A_BI, JT, JR, M (whole batch)."""
        N, qdim = int(q.shape[0]), int(q.shape[1])
        eff, ne = _int_array(effectors)
        st = self._stats()
        rc = lib().rcsb_group_fk_jacobians_mass(
            self._h, N, qdim, _host_ptr(q, "q"), 0, None, _host_ptr(out.get("A_BI"), "A_BI"), ne, eff, None,
            _host_ptr(out.get("JT"), "JT"), _host_ptr(out.get("JR"), "JR"), _host_ptr(out.get("M"), "M"),
            _host_ptr(st, "stats"))
        _err(rc, "rcsb_group_fk_jacobians_mass")
        return st

    def kinetic_terms(self, q, qd, out: Dict):
        N, qdim = int(q.shape[0]), int(q.shape[1])
        st = self._stats()
        rc = lib().rcsb_group_kinetic_terms(
            self._h, N, qdim, _host_ptr(q, "q"), _host_ptr(qd, "qd"), _host_ptr(out.get("M"), "M"),
            _host_ptr(out.get("h"), "h"), _host_ptr(out.get("Fg"), "Fg"), _host_ptr(out.get("E"), "E"),
            _host_ptr(st, "stats"))
        _err(rc, "rcsb_group_kinetic_terms")
        return st

    def ik_rmr(self, tasks: Sequence, q, x_des, out: Dict, lam=1.0e-8, alpha=0.05, iters=1):
        N = int(q.shape[0])
        tarr = (_RcsbTask * max(len(tasks), 1))(*[_RcsbTask(int(k), int(e)) for k, e in tasks])
        st = self._stats()
        rc = lib().rcsb_group_ik_rmr(
            self._h, len(tasks), tarr, N, _host_ptr(q, "q"), _host_ptr(x_des, "x_des"), float(lam), float(alpha),
            int(iters), _host_ptr(out.get("dx"), "dx"), _host_ptr(out.get("dq"), "dq"),
            _host_ptr(out.get("det"), "det"), _host_ptr(st, "stats"))
        _err(rc, "rcsb_group_ik_rmr")
        return st

    def wholebody(self, q, spec, out: Dict):
        N, qdim = int(q.shape[0]), int(q.shape[1])
        arr = (_RcsbRelJac * max(len(spec), 1))(
            *[_RcsbRelJac(1 if k == "POS" else 2, int(e), int(rb), int(rf)) for k, e, rb, rf in spec])
        st = self._stats()
        has_j = out.get("Jt") is not None
        rc = lib().rcsb_group_wholebody(
            self._h, N, qdim, _host_ptr(q, "q"), 0, None, _host_ptr(out.get("A_BI"), "A_BI"),
            len(spec) if has_j else 0, arr if has_j else None, _host_ptr(out.get("Jt"), "Jt"),
            _host_ptr(out.get("M"), "M"), _host_ptr(out.get("Jcog"), "Jcog"), _host_ptr(st, "stats"))
        _err(rc, "rcsb_group_wholebody")
        return st
