"""ctypes binding of libga_sm100.so (the C ABI declared in include/ga_sm100.h).

There is no CPU fallback: importing this module only loads the shared library; calling any kernel without a
CUDA device (or without the built library) raises.  PyTorch is used for device memory and streams only -- every
entry point receives raw device pointers plus `torch.cuda.current_stream().cuda_stream`.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libga_sm100.so')
HEADER_PATH = os.path.join(os.path.dirname(_HERE), 'include', 'ga_sm100.h')

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_RELU, ACT_MUL = 0, 1, 2, 3
BACKEND_AUTO, BACKEND_SIMT, BACKEND_TCGEN05 = 0, 1, 2


class GaGemm(C.Structure):
    _fields_ = [
        ('A', C.c_void_p), ('a_rs', C.c_longlong), ('a_cs', C.c_longlong), ('a_bs', C.c_longlong),
        ('B', C.c_void_p), ('b_rs', C.c_longlong), ('b_cs', C.c_longlong), ('b_bs', C.c_longlong),
        ('D', C.c_void_p), ('ldd', C.c_longlong), ('d_bs', C.c_longlong), ('d_cs', C.c_longlong),
        ('M', C.c_int), ('N', C.c_int), ('K', C.c_int), ('batch', C.c_int),
        ('in_dtype', C.c_int), ('out_dtype', C.c_int), ('accumulate', C.c_int), ('alpha', C.c_float),
        ('bias', C.c_void_p), ('bias_bs', C.c_longlong),
        ('act', C.c_int),
        ('Z', C.c_void_p),
        ('colscale', C.c_void_p), ('colscale_bs', C.c_longlong),
        ('rowscale', C.c_void_p), ('rows_per_scale', C.c_int),
        ('R', C.c_void_p), ('ldr', C.c_longlong), ('r_bs', C.c_longlong),
        ('Zin', C.c_void_p), ('ldz', C.c_longlong), ('z_bs', C.c_longlong), ('zmode', C.c_int),
        ('backend', C.c_int), ('splits', C.c_int), ('z_shadow', C.c_int),
        ('backend_used', C.POINTER(C.c_int)),
        ('colsum', C.c_void_p),
        ('ln_xhat', C.c_void_p), ('ld_xhat', C.c_longlong), ('ln_rstd', C.c_void_p),
    ]


def build(force: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libga_sm100.so (in-tree).  nvcc cross-compiles without a GPU."""
    csrc = os.path.join(_HERE, 'csrc')
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    subprocess.run(['make', '-j8', '-C', csrc], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


def exported_symbols():
    """Function names declared in include/ga_sm100.h (what the library must export)."""
    text = open(HEADER_PATH).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ga_[a-z0-9_]+)\s*\(', text)))


_lib = None
_timed = None
_UNTIMED = ('ga_last_error', 'ga_launch_count', 'ga_version', 'ga_workspace_bytes')


class CallTimer:
    """Optional CUDA-event instrumentation of EVERY C-ABI call (bench.py's live per-kernel roofline): events are recorded on
    the launching stream around each `ga_*` entry point, keyed by (entry point, integer arguments = the call site's shape).
    Enabled with start_timing(); read with summary() after a synchronize."""

    def __init__(self, lib):
        self._lib, self.records = lib, []

    @staticmethod
    def _sig(name, args):
        if name == 'ga_gemm':
            g = args[0]._obj
            return (g.batch, g.M, g.N, g.K, g.in_dtype, g.out_dtype, g.act, int(bool(g.Z)), int(g.z_shadow), int(bool(g.R)),
                    int(bool(g.Zin)), g.accumulate, int(g.a_cs != 1), int(g.b_cs != 1))
        sig = []
        for a in args:
            if isinstance(a, int):
                sig.append(a)
            elif isinstance(a, (C.c_int, C.c_longlong)):
                sig.append(a.value)
        return tuple(sig)

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.startswith('ga_') or name in _UNTIMED or name.endswith('_parts'):
            return fn

        def timed(*args):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            self.records.append((name, self._sig(name, args), e0, e1))
            return rc
        return timed

    def summary(self):
        """{(entry point, signature): (calls, total ms)}"""
        agg = {}
        for name, sig, e0, e1 in self.records:
            n, t = agg.get((name, sig), (0, 0.0))
            agg[(name, sig)] = (n + 1, t + e0.elapsed_time(e1))
        return agg


def start_timing() -> 'CallTimer':
    global _timed
    _timed = CallTimer(load_raw())
    return _timed


def stop_timing():
    global _timed
    _timed = None


def load_raw() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                               '(there is no CPU or PyTorch fallback for the GA kernels)')
        _lib = C.CDLL(LIB_PATH)
        _lib.ga_launch_count.restype = C.c_longlong
        _lib.ga_workspace_bytes.restype = C.c_longlong
    return _lib


def load():
    """The library handle every op calls through (a CallTimer proxy while start_timing() is active)."""
    return _timed if _timed is not None else load_raw()


class GaError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != 0:
        buf = C.create_string_buffer(512)
        load_raw().ga_last_error(buf, C.c_size_t(len(buf)))
        raise GaError(f'{what} failed (code {rc}): {buf.value.decode()}')


def stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f'unsupported dtype {t.dtype} (fp32 or bf16 only)')


def ptr(t) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise GaError('GA kernels need CUDA tensors: there is no CPU path')
    return C.c_void_p(t.data_ptr())


def launch_count() -> int:
    return int(load().ga_launch_count())


def ll(v) -> C.c_longlong:
    return C.c_longlong(int(v))


def f(v) -> C.c_float:
    return C.c_float(float(v))
