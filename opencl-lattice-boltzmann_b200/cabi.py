"""ctypes binding of include/lbm.h (liblbm_b200.so) — what tests/ and bench.py call.

There is no fallback: if the CUDA library is missing or no GPU is visible the calls
raise ``LbmError`` (the product path never routes through oracle/ or any CPU code).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .decks import NSPEEDS, Params

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "liblbm_b200.so")

# every symbol include/lbm.h declares (checked by tests/test_cabi_symbols.py)
EXPORTS = [
    "lbm_create", "lbm_create_on", "lbm_create_slab", "lbm_partition_rows", "lbm_export_size", "lbm_export",
    "lbm_connect", "lbm_destroy", "lbm_upload", "lbm_halo_push", "lbm_download_cells", "lbm_download_av_vels",
    "lbm_download_av_sums", "lbm_download_final_state", "lbm_combine_av_sums", "lbm_host_alloc", "lbm_host_free", "lbm_run", "lbm_sync",
    "lbm_upload_packed", "lbm_mask_words_per_row", "lbm_pack_obstacles", "lbm_host_alloc_on", "lbm_device_numa_node",
    "lbm_run_timed", "lbm_set_option", "lbm_get_info", "lbm_debug_pad_nonzero", "lbm_debug_fastmath_mismatches", "lbm_debug_tile_timing", "lbm_device_count", "lbm_abi_version", "lbm_last_error",
]


class LbmError(RuntimeError):
    pass


class LbmParams(C.Structure):
    """lbm_params == t_param (d2q9-bgk.c:81-92)."""
    _fields_ = [("density", C.c_float), ("accel", C.c_float), ("omega", C.c_float),
                ("free_cells_inv", C.c_float), ("nx", C.c_int), ("ny", C.c_int),
                ("maxIters", C.c_int), ("reynolds_dim", C.c_int)]


class LbmInfo(C.Structure):
    _fields_ = [("abi_version", C.c_int), ("nslabs", C.c_int), ("rank", C.c_int), ("nranks", C.c_int),
                ("y0", C.c_int), ("rows", C.c_int), ("pitch", C.c_int), ("cells_per_thread", C.c_int),
                ("threads_per_block", C.c_int), ("streaming", C.c_int), ("steps_per_launch", C.c_int),
                ("steps_done", C.c_longlong), ("kernel_launches", C.c_longlong),
                ("partials_per_step", C.c_longlong), ("kernel_name", C.c_char * 64)]


_lib = None


def load_library(path: str | None = None):
    """dlopen the in-tree library and declare the prototypes.  Raises LbmError if absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise LbmError(f"{path} is not built — run `make` (or __graft_entry__.build()) first")
    lib = C.CDLL(path)
    vp, fp, ip, dp = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double)
    PP = C.POINTER(LbmParams)
    lib.lbm_create.argtypes = [C.POINTER(vp), PP, C.c_int]
    lib.lbm_create_on.argtypes = [C.POINTER(vp), PP, C.c_int, ip]
    lib.lbm_create_slab.argtypes = [C.POINTER(vp), PP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.lbm_partition_rows.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip]
    lib.lbm_partition_rows.restype = None
    lib.lbm_export_size.argtypes = []
    lib.lbm_export_size.restype = C.c_size_t
    lib.lbm_export.argtypes = [vp, vp]
    lib.lbm_connect.argtypes = [vp, vp, vp]
    lib.lbm_destroy.argtypes = [vp]
    lib.lbm_destroy.restype = None
    lib.lbm_upload.argtypes = [vp, vp, vp]
    lib.lbm_upload_packed.argtypes = [vp, vp, vp]
    lib.lbm_mask_words_per_row.argtypes = [C.c_int]
    lib.lbm_mask_words_per_row.restype = C.c_size_t
    lib.lbm_pack_obstacles.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.lbm_pack_obstacles.restype = None
    lib.lbm_host_alloc_on.argtypes = [C.c_size_t, C.c_int]
    lib.lbm_host_alloc_on.restype = vp
    lib.lbm_device_numa_node.argtypes = [C.c_int]
    lib.lbm_halo_push.argtypes = [vp]
    lib.lbm_download_cells.argtypes = [vp, vp]
    lib.lbm_download_av_vels.argtypes = [vp, fp, C.c_int]
    lib.lbm_download_av_sums.argtypes = [vp, dp, dp, C.c_int]
    lib.lbm_download_final_state.argtypes = [vp, vp, vp, vp, vp]
    lib.lbm_combine_av_sums.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int, C.c_float, fp]
    lib.lbm_combine_av_sums.restype = None
    lib.lbm_host_alloc.argtypes = [C.c_size_t]
    lib.lbm_host_alloc.restype = vp
    lib.lbm_host_free.argtypes = [vp]
    lib.lbm_host_free.restype = None
    lib.lbm_run.argtypes = [vp, C.c_int]
    lib.lbm_sync.argtypes = [vp]
    lib.lbm_run_timed.argtypes = [vp, C.c_int, fp]
    lib.lbm_set_option.argtypes = [vp, C.c_char_p, C.c_long]
    lib.lbm_get_info.argtypes = [vp, C.POINTER(LbmInfo)]
    lib.lbm_debug_pad_nonzero.argtypes = [vp, C.POINTER(C.c_longlong)]
    lib.lbm_debug_fastmath_mismatches.argtypes = [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    lib.lbm_debug_tile_timing.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.lbm_device_count.argtypes = []
    lib.lbm_abi_version.argtypes = []
    lib.lbm_last_error.argtypes = []
    lib.lbm_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def fastmath_mismatches():
    """lbm_debug_fastmath_mismatches: (rcp, sqrt) mismatch counts of the two-step kernel's fast
    reciprocal / square root against the correctly rounded built-ins over all 2^32 floats."""
    lib = load_library()
    a, b = C.c_ulonglong(0), C.c_ulonglong(0)
    if lib.lbm_debug_fastmath_mismatches(C.byref(a), C.byref(b)) != 0:
        raise LbmError(lib.lbm_last_error().decode())
    return int(a.value), int(b.value)


def partition_rows(ny: int, nparts: int, part: int):
    """lbm_partition_rows in Python (kept in lock-step by tests/test_host_logic.py)."""
    base, rem = divmod(ny, nparts)
    rows = base + (1 if part < rem else 0)
    y0 = part * base + min(part, rem)
    return y0, rows


def to_c_params(p: Params) -> LbmParams:
    return LbmParams(p.density, p.accel, p.omega, p.free_cells_inv, p.nx, p.ny, p.maxIters, p.reynolds_dim)


def combine_av_sums(hi: np.ndarray, lo: np.ndarray, free_cells_inv: float) -> np.ndarray:
    """lbm_combine_av_sums in numpy: hi/lo are [nparts, n]; parts are added in index order
    with an error-free TwoSum so the result does not depend on how rows were split."""
    hi = np.asarray(hi, dtype=np.float64)
    lo = np.asarray(lo, dtype=np.float64)
    H = np.zeros(hi.shape[1], dtype=np.float64)
    L = np.zeros(hi.shape[1], dtype=np.float64)
    for part in range(hi.shape[0]):
        x = hi[part]
        s = H + x
        bb = s - H
        err = (H - (s - bb)) + (x - bb)
        H = s
        L = (L + lo[part]) + err
    return ((H + L) * np.float64(np.float32(free_cells_inv))).astype(np.float32)


class PinnedArray:
    """A numpy array over pinned host memory from lbm_host_alloc_on(bytes, device): pages on the NUMA
    node the GPU is attached to.  Keep the object alive as long as `.array` is used; close() (or the
    finaliser) returns the memory with lbm_host_free."""

    def __init__(self, shape, dtype, device: int):
        self.lib = load_library()
        self.array = None
        self._ptr = None
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = self.lib.lbm_host_alloc_on(max(nbytes, 1), device)
        if not ptr:
            raise LbmError(self.lib.lbm_last_error().decode())
        self._ptr = ptr
        buf = (C.c_char * max(nbytes, 1)).from_address(ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self.numa_node = int(self.lib.lbm_device_numa_node(device))

    def close(self):
        if self._ptr:
            self.array = None
            self.lib.lbm_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_obstacles(obstacles: np.ndarray) -> np.ndarray:
    """lbm_pack_obstacles: [rows, nx] int32 map -> [rows, ceil(nx/32)] uint32 bit mask (bit x&31 of word
    x>>5 set = blocked), the layout lbm_upload_packed takes."""
    lib = load_library()
    obstacles = np.ascontiguousarray(obstacles, dtype=np.int32)
    rows, nx = obstacles.shape
    out = np.empty((rows, int(lib.lbm_mask_words_per_row(nx))), dtype=np.uint32)
    lib.lbm_pack_obstacles(C.c_void_p(obstacles.ctypes.data), nx, rows, C.c_void_p(out.ctypes.data))
    return out


class Simulation:
    """One lbm_ctx.  Mirrors the reference host's use of its device state (d2q9-bgk.c:194-277):
    create -> upload -> run -> sync -> download -> destroy."""

    def __init__(self, params: Params, ngpus: int = 1, devices=None, slab=None, options=None):
        """slab = (device, rank, nranks, y0, rows) creates the one-process-per-GPU form."""
        self.lib = load_library()
        self.params = params
        self._cp = to_c_params(params)
        self._ctx = C.c_void_p()
        if slab is not None:
            device, rank, nranks, y0, rows = slab
            self._ck(self.lib.lbm_create_slab(C.byref(self._ctx), C.byref(self._cp), device, rank, nranks, y0, rows))
            self.rows = rows
        elif devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            self._ck(self.lib.lbm_create_on(C.byref(self._ctx), C.byref(self._cp), len(devices), arr))
            self.rows = params.ny
        else:
            self._ck(self.lib.lbm_create(C.byref(self._ctx), C.byref(self._cp), ngpus))
            self.rows = params.ny
        for k, v in (options or {}).items():
            self.set_option(k, v)

    def _ck(self, status: int):
        if status != 0:
            raise LbmError(self.lib.lbm_last_error().decode())

    # -- options / info ------------------------------------------------------
    def set_option(self, key: str, value: int):
        self._ck(self.lib.lbm_set_option(self._ctx, key.encode(), int(value)))

    def info(self) -> dict:
        info = LbmInfo()
        self._ck(self.lib.lbm_get_info(self._ctx, C.byref(info)))
        d = {name: getattr(info, name) for name, _ in LbmInfo._fields_}
        d["kernel_name"] = info.kernel_name.decode()
        return d

    def tile_timing(self) -> np.ndarray:
        """[64, 16] SM clock stamps of tile 0 (option tile_debug = 1); see csrc/lbm_tile.cuh."""
        out = np.zeros((64, 16), dtype=np.int64)
        self._ck(self.lib.lbm_debug_tile_timing(self._ctx, C.c_void_p(out.ctypes.data), 64, 16))
        return out

    def pad_nonzero(self) -> int:
        n = C.c_longlong(0)
        self._ck(self.lib.lbm_debug_pad_nonzero(self._ctx, C.byref(n)))
        return int(n.value)

    # -- ring plumbing ---------------------------------------------------------
    def export_blob(self) -> bytes:
        buf = C.create_string_buffer(self.lib.lbm_export_size())
        self._ck(self.lib.lbm_export(self._ctx, buf))
        return buf.raw

    def connect(self, blob_down: bytes, blob_up: bytes):
        self._ck(self.lib.lbm_connect(self._ctx, C.c_char_p(blob_down), C.c_char_p(blob_up)))

    def halo_push(self):
        self._ck(self.lib.lbm_halo_push(self._ctx))

    # -- data ------------------------------------------------------------------
    @staticmethod
    def _ptr(a):
        """numpy array or torch tensor (CPU) -> void*"""
        if hasattr(a, "data_ptr"):
            return C.c_void_p(a.data_ptr())
        return C.c_void_p(a.ctypes.data)

    def upload(self, cells, obstacles):
        """cells: [9, rows, nx] float32, obstacles: [rows, nx] int32 (numpy, or pinned torch CPU tensors)."""
        n = self.rows * self.params.nx
        if isinstance(cells, np.ndarray):
            cells = np.ascontiguousarray(cells, dtype=np.float32)
            obstacles = np.ascontiguousarray(obstacles, dtype=np.int32)
            assert cells.size == NSPEEDS * n and obstacles.size == n
        else:
            assert cells.numel() == NSPEEDS * n and obstacles.numel() == n and cells.is_contiguous()
        self._keep = (cells, obstacles)
        self._ck(self.lib.lbm_upload(self._ctx, self._ptr(cells), self._ptr(obstacles)))

    def upload_packed(self, cells, mask_words):
        """cells as for upload(); mask_words: [rows, ceil(nx/32)] uint32 from pack_obstacles()."""
        n = self.rows * self.params.nx
        words = self.rows * int(self.lib.lbm_mask_words_per_row(self.params.nx))
        if isinstance(cells, np.ndarray):
            cells = np.ascontiguousarray(cells, dtype=np.float32)
            assert cells.size == NSPEEDS * n
        else:
            assert cells.numel() == NSPEEDS * n and cells.is_contiguous()
        if isinstance(mask_words, np.ndarray):
            mask_words = np.ascontiguousarray(mask_words, dtype=np.uint32)
            assert mask_words.size == words
        self._keep = (cells, mask_words)
        self._ck(self.lib.lbm_upload_packed(self._ctx, self._ptr(cells), self._ptr(mask_words)))

    def run(self, nsteps: int):
        self._ck(self.lib.lbm_run(self._ctx, nsteps))

    def run_timed(self, nsteps: int) -> float:
        ms = C.c_float(0.0)
        self._ck(self.lib.lbm_run_timed(self._ctx, nsteps, C.byref(ms)))
        return float(ms.value)

    def sync(self):
        self._ck(self.lib.lbm_sync(self._ctx))

    def download_cells(self, out=None):
        if out is None:
            out = np.empty((NSPEEDS, self.rows, self.params.nx), dtype=np.float32)
        self._ck(self.lib.lbm_download_cells(self._ctx, self._ptr(out)))
        return out

    def download_av_vels(self, n: int) -> np.ndarray:
        av = np.empty(max(n, 1), dtype=np.float32)
        self._ck(self.lib.lbm_download_av_vels(self._ctx, av.ctypes.data_as(C.POINTER(C.c_float)), n))
        return av[:n]

    def download_av_sums(self, n: int):
        hi = np.empty(max(n, 1), dtype=np.float64)
        lo = np.empty(max(n, 1), dtype=np.float64)
        dp = C.POINTER(C.c_double)
        self._ck(self.lib.lbm_download_av_sums(self._ctx, hi.ctypes.data_as(dp), lo.ctypes.data_as(dp), n))
        return hi[:n], lo[:n]

    def download_final_state(self):
        """(u_x, u_y, u, pressure), each [rows, nx] float32, computed on the device (d2q9-bgk.c:789-831)."""
        outs = [np.empty((self.rows, self.params.nx), dtype=np.float32) for _ in range(4)]
        self._ck(self.lib.lbm_download_final_state(self._ctx, *[self._ptr(o) for o in outs]))
        return outs

    def close(self):
        if self._ctx:
            self.lib.lbm_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
