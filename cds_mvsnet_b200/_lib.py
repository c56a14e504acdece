"""ctypes binding of libcds_b200.so -- the only way the Python host reaches the CUDA kernels.

The prototypes are read from ``include/cds_b200.h`` so the header stays the single source of truth
for the C ABI.  There is NO fallback: if the library is missing or a call fails, a RuntimeError is
raised (the product path never routes through torch ops or the CPU oracle).
"""
from __future__ import annotations

import ctypes
import os
import re

import torch

PKG = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(PKG), "include", "cds_b200.h")
LIB_PATH = os.path.join(PKG, "libcds_b200.so")

CDS_F32, CDS_F16 = 0, 1
ACT_NONE, ACT_LRELU, ACT_TANH = 0, 1, 2

_CTYPES = {
    "int": ctypes.c_int, "float": ctypes.c_float, "long long": ctypes.c_longlong,
    "cudaStream_t": ctypes.c_void_p, "void": None,
}


def _ctype_of(decl: str):
    decl = decl.strip()
    if "*" in decl:
        return ctypes.c_void_p
    base = re.sub(r"\bconst\b", "", decl).strip()
    base = re.sub(r"\s+\w+$", "", base).strip() if base not in _CTYPES else base
    if base not in _CTYPES:
        raise ValueError(f"unmapped C type in header: {decl!r}")
    return _CTYPES[base]


def parse_header(path: str = HEADER) -> dict:
    """{name: (restype, [argtypes])} for every prototype declared in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"^\s*(const char\*|int)\s+(cds_\w+)\s*\(([^)]*)\)\s*;", text, flags=re.M):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if ret.startswith("const char") else ctypes.c_int
        argtypes = [] if args in ("", "void") else [_ctype_of(a) for a in args.split(",")]
        protos[name] = (restype, argtypes)
    return protos


class _Lib:
    def __init__(self):
        self._dll = None
        self._devices_checked = set()   # CUDA device indices cds_check_device has accepted

    def load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA library is the product and there is no fallback. "
                "Build it with `python -m cds_mvsnet_b200.build` (or __graft_entry__.build()).")
        dll = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in parse_header().items():
            fn = getattr(dll, name)  # AttributeError here = header/library mismatch
            fn.restype = restype
            fn.argtypes = argtypes
        self._dll = dll
        return dll

    def check_device(self, index: int):
        """cds_check_device on the CUDA runtime's current device (once per device index)."""
        if index in self._devices_checked:
            return
        if not torch.cuda.is_available():
            raise RuntimeError("cds_b200 needs a CUDA device (B200, sm_100a); none is visible")
        rc = self.load().cds_check_device()
        if rc != 0:
            raise RuntimeError(self._dll.cds_last_error_string().decode())
        self._devices_checked.add(index)

    def call(self, name: str, *args):
        dll = self.load()
        rc = getattr(dll, name)(*args)
        if rc != 0:
            raise RuntimeError(f"{name} failed (code {rc}): {dll.cds_last_error_string().decode()}")


LIB = _Lib()


# The library launches on the CUDA runtime's CURRENT device with the stream it is handed.  ``ptr`` therefore notes the device
# of every tensor that becomes a kernel argument and ``call`` (1) refuses a launch whose tensors live on different devices,
# (2) makes that device current for the launch and (3) passes torch's current stream OF THAT DEVICE -- so
# ``model.to("cuda:1")`` works whatever the process-wide current device is.
_ARG_DEVICES = []   # device indices of the tensors handed to ``ptr`` since the last launch


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  CPU tensors are refused: there is no CPU fallback."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("cds_b200: a CPU tensor reached a kernel argument; there is no CPU fallback")
    _ARG_DEVICES.append(t.device.index)
    return t.data_ptr()


def _launch_device():
    """Device index the next launch goes to: the one its tensor arguments live on (the current device if it has none)."""
    if not _ARG_DEVICES:
        return torch.cuda.current_device()
    idx = _ARG_DEVICES[0]
    for d in _ARG_DEVICES:
        if d != idx:
            devs = sorted(set(_ARG_DEVICES))
            _ARG_DEVICES.clear()
            raise RuntimeError(f"cds_b200: one launch was handed tensors on different devices (cuda:{devs})")
    _ARG_DEVICES.clear()
    return idx


LAUNCHES = 0   # C-ABI calls issued by this process (every inference entry enqueues exactly one kernel)


class LaunchProfile:
    """Optional per-launch CUDA-event timing (bench.py): ``with LaunchProfile() as p: ...; p.summary()``."""

    def __init__(self):
        self.records = []   # (name, tag, meta, start_event, end_event)

    def __enter__(self):
        global _PROFILE
        _PROFILE = self
        return self

    def __exit__(self, *exc):
        global _PROFILE
        _PROFILE = None

    def summary(self):
        """{(name, tag): dict(ms_total, launches, meta)} -- call after a device synchronize."""
        out = {}
        for name, tag, meta, a, b in self.records:
            d = out.setdefault((name, tag), dict(ms_total=0.0, launches=0, meta=meta))
            d["ms_total"] += a.elapsed_time(b)
            d["launches"] += 1
        return out


_PROFILE = None
_TAG = [None, None]   # (tag, meta) attached to the launches issued next; set by the engine


def set_tag(tag, meta=None):
    _TAG[0], _TAG[1] = tag, meta


def call(name, *args):
    global LAUNCHES
    idx = _launch_device()
    if idx != torch.cuda.current_device():
        with torch.cuda.device(idx):   # also switches the CUDA runtime's device, which is what the library launches on
            return call(name, *args)
    LIB.check_device(idx)
    st = torch.cuda.current_stream(idx)
    prof = _PROFILE
    if prof is not None:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        LIB.call(name, *args, st.cuda_stream)
        b.record(st)
        prof.records.append((name, _TAG[0], _TAG[1], a, b))
    else:
        LIB.call(name, *args, st.cuda_stream)
    LAUNCHES += 1


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float16:
        return CDS_F16
    if dt == torch.float32:
        return CDS_F32
    raise ValueError(f"storage dtype must be float16 or float32, got {dt}")
