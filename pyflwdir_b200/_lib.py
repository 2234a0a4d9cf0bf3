"""ctypes binding of libpfd_b200.so (C ABI declared in include/pfd_b200.h).

There is no CPU fallback: if the shared library has not been built, or no CUDA device is usable, every compute
entry point raises. Build with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C pyflwdir_b200/csrc``.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PFD_B200_LIB") or os.path.join(_HERE, "libpfd_b200.so")  # override: kernel experiments

# pfd_status
OK, ERR_CUDA, ERR_INVALID_ARG, ERR_INVALID_D8, ERR_NO_PITS, ERR_STATE, ERR_UNSUPPORTED, ERR_OOM, ERR_NCCL = range(9)

# pfd_dtype
DTYPES = {
    np.dtype(np.int8): 0, np.dtype(np.uint8): 1, np.dtype(np.int16): 2, np.dtype(np.uint16): 3,
    np.dtype(np.int32): 4, np.dtype(np.uint32): 5, np.dtype(np.int64): 6, np.dtype(np.uint64): 7,
    np.dtype(np.float32): 8, np.dtype(np.float64): 9,
}

# pfd_array
ARR_IDXS_DS, ARR_PITS, ARR_PIT_IS_OUTLET, ARR_SEQ, ARR_RANK, ARR_N_UPSTREAM, ARR_D8, ARR_LEVEL_OFFSETS, ARR_LDD, \
    ARR_SUBBASIN_OUTLETS, ARR_REGION_LABELS, ARR_REGION_SLICES, ARR_NEXTXY, ARR_STREAM_OFFSETS, \
    ARR_STREAM_CELLS = range(15)

# every symbol include/pfd_b200.h declares: name -> (restype, argtypes)
_vp, _i64, _int, _u32 = C.c_void_p, C.c_int64, C.c_int, C.c_uint32
_pi64 = C.POINTER(C.c_int64)
SYMBOLS = {
    "pfd_version": (C.c_char_p, []),
    "pfd_device_count": (_int, []),
    "pfd_status_string": (C.c_char_p, [_int]),
    "pfd_device_pci_bus_id": (_int, [_int, C.c_char_p, _int]),
    "pfd_create": (_int, [_int, C.POINTER(_vp)]),
    "pfd_destroy": (None, [_vp]),
    "pfd_last_error": (C.c_char_p, [_vp]),
    "pfd_host_alloc": (_int, [C.c_size_t, C.POINTER(_vp)]),
    "pfd_host_free": (_int, [_vp]),
    "pfd_dev_alloc": (_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "pfd_dev_free": (_int, [_vp, _vp]),
    "pfd_memcpy": (_int, [_vp, _vp, _vp, C.c_size_t]),
    "pfd_synchronize": (_int, [_vp]),
    "pfd_d8_parse": (_int, [_vp, _vp, _i64, _i64, _int, _vp, _int, _pi64, _pi64, _pi64]),
    "pfd_ldd_parse": (_int, [_vp, _vp, _i64, _i64, _vp, _int, _pi64, _pi64]),
    "pfd_nextxy_parse": (_int, [_vp, _vp, _vp, _i64, _i64, _int, _vp, _int, _pi64, _pi64, _pi64]),
    "pfd_load_idxs_ds": (_int, [_vp, _vp, _int, _i64, _i64, _pi64, _pi64]),
    "pfd_order": (_int, [_vp, _pi64, _pi64]),
    "pfd_fetch": (_int, [_vp, _int, _vp, _int]),
    "pfd_accuflux": (_int, [_vp, _vp, _int, C.c_double, _i64, _int, _int, _vp]),
    "pfd_upstream_area_cells": (_int, [_vp, _vp]),
    "pfd_basins": (_int, [_vp, _vp, _i64, _int, _vp, _int, _vp]),
    "pfd_strahler": (_int, [_vp, _vp, _vp]),
    "pfd_hand": (_int, [_vp, _vp, _vp, _int, _vp]),
    "pfd_fillnodata": (_int, [_vp, _vp, _int, C.c_double, _i64, _int, _int, _int, _vp]),
    "pfd_main_upstream": (_int, [_vp, _vp, _int, C.c_double, _vp, _int]),
    "pfd_upstream_count": (_int, [_vp, _vp, _vp]),
    "pfd_upstream_matrix": (_int, [_vp, _vp, _int, _i64, _pi64]),
    "pfd_stream_order_classic": (_int, [_vp, _vp, _int, _vp, _vp]),
    "pfd_stream_distance": (_int, [_vp, _vp, _int, _vp, _vp]),
    "pfd_floodplains": (_int, [_vp, _vp, _vp, _int, _vp]),
    "pfd_upstream_sum": (_int, [_vp, _vp, _int, C.c_double, _i64, _int, _vp]),
    "pfd_subbasins_streamorder": (_int, [_vp, _vp, _vp, _i64, _vp, _pi64]),
    "pfd_subbasins_area": (_int, [_vp, _vp, _int, _vp, _int, C.c_double, _vp, _pi64]),
    "pfd_moving_average": (_int, [_vp, _vp, _int, _vp, _int, _int, _vp, _int, _vp, C.c_double, _vp]),
    "pfd_moving_median": (_int, [_vp, _vp, _int, _int, _vp, _int, _vp, C.c_double, _vp]),
    "pfd_downstream": (_int, [_vp, _vp, _int, _vp]),
    "pfd_trace": (_int, [_vp, _vp, _i64, _int, _vp, _int, _vp, _int, C.c_double, _vp, _vp, _vp, _vp, _vp, _int, _i64]),
    "pfd_inflow_idxs": (_int, [_vp, _vp, _pi64]),
    "pfd_outflow_idxs": (_int, [_vp, _vp, _pi64]),
    "pfd_interbasin_mask": (_int, [_vp, _vp, _vp, _vp]),
    "pfd_region_outlets": (_int, [_vp, _vp, _int, _pi64]),
    "pfd_subbasins_pfafstetter": (_int, [_vp, _vp, _int, _vp, _int, _vp, _int, _vp, _pi64]),
    "pfd_classify_estuary": (_int, [_vp, _vp, _vp, _int, _vp, _int, C.c_double, _vp]),
    "pfd_streams": (_int, [_vp, _vp, _i64, _pi64, _pi64]),
    "pfd_region_slices": (_int, [_vp, _vp, _int, _pi64]),
    "pfd_d8_flow_all": (_int, [_vp, _vp, _i64, _i64, _vp, _int, _vp, _vp, _vp, _pi64, _pi64, _pi64]),
    "pfd_comm_unique_id": (_int, [_vp, _i64]),
    "pfd_comm_init": (_int, [_vp, _int, _int, _vp]),
    "pfd_comm_barrier": (_int, [_vp]),
    "pfd_comm_destroy": (_int, [_vp]),
    "pfd_d8_flow_all_tiled": (_int, [_vp, _vp, _i64, _i64, _int, _int, _i64, _vp, _int, _vp, _vp, _vp, _pi64, _pi64]),
    "pfd_tiled_parse": (_int, [_vp, _vp, _i64, _i64, _int, _int, _i64, _vp, _int, _pi64, _pi64]),
    "pfd_tiled_local": (_int, [_vp, _int, _int, _i64, _vp, C.POINTER(_vp), _pi64]),
    "pfd_tiled_finish": (_int, [_vp, _vp, _vp, _vp]),
    "pfd_sweep_tiled_begin": (_int, [_vp, _int, _vp, _int, _vp, C.c_double, _i64, _int]),
    "pfd_sweep_tiled_round": (_int, [_vp, _pi64]),
    "pfd_sweep_tiled_edges": (_int, [_vp, _int, C.POINTER(_vp), _pi64]),
    "pfd_sweep_tiled_halo": (_int, [_vp, _int, _vp]),
    "pfd_sweep_tiled_end": (_int, [_vp, _vp, _pi64]),
    "pfd_sweep_tiled": (_int, [_vp, _int, _vp, _int, _vp, C.c_double, _i64, _int, _vp, _pi64]),
    "pfd_synth_elevation": (_int, [_vp, _i64, _i64, _i64, _int, _u32, _vp]),
    "pfd_synth_d8": (_int, [_vp, _vp, _i64, _i64, C.c_float, _vp]),
    "pfd_set_option": (_int, [_vp, C.c_char_p, _i64]),
    "pfd_get_info": (_i64, [_vp, C.c_char_p]),
    "pfd_fill_depressions": (_int, [_vp, _vp, _int, _i64, _i64, _int, _vp, _i64, C.c_double, C.c_double, _int, C.c_double, _int, _int,
                             _vp, _vp, _vp]),
    "pfd_synth_d8_block": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _int, _u32, C.c_float, _vp]),
    "pfd_verify_flow": (_int, [_vp, _vp, _int, _vp, _vp, _vp, _pi64]),
    "pfd_verify_strahler": (_int, [_vp, _vp, _vp, _pi64]),
    "pfd_verify_hand": (_int, [_vp, _vp, _vp, _int, _vp, _pi64]),
    "pfd_verify_accuflux": (_int, [_vp, _vp, _int, C.c_double, _i64, _int, _vp, _pi64]),
    "pfd_checksum": (_int, [_vp, _vp, _int, _i64, C.c_uint64, C.POINTER(C.c_uint64)]),
    "pfd_launch_count": (_i64, [_vp]),
    "pfd_timer_start": (_int, [_vp]),
    "pfd_timer_stop": (_int, [_vp, C.POINTER(C.c_double)]),
    "pfd_last_stage_ms": (C.c_double, [_vp, _int]),
}

_lib = None


class PfdError(RuntimeError):
    """Failure inside libpfd_b200 that is not a user-input error (CUDA, OOM, state)."""

    def __init__(self, status, message):
        super().__init__(f"libpfd_b200 [{status}]: {message}")
        self.status = status


def lib():
    """Load the CUDA library; raises ImportError loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built and pyflwdir_b200 has no CPU "
                "fallback. Run `make -C pyflwdir_b200/csrc` (or __graft_entry__.build())."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def device_count():
    return int(lib().pfd_device_count())


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_device_numa_node(device):
    """One process per GPU: run this process on the CPUs of the GPU's NUMA node, so that the pinned buffers it allocates
    afterwards (first touched by this process) sit next to the GPU's PCIe root and the D2H / H2D copies of several ranks
    do not all cross the socket interconnect. Returns the node (or None when the topology is not exposed)."""
    buf = C.create_string_buffer(32)
    if lib().pfd_device_pci_bus_id(int(device), buf, 32) != OK:
        return None
    bus = buf.value.decode().lower()
    try:
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read()) & set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, ValueError):
        return None


def check(status, handle=None):
    """Map a pfd_status to the exception type the reference raises for the same condition."""
    if status == OK:
        return
    msg = lib().pfd_last_error(handle)
    msg = msg.decode() if msg else lib().pfd_status_string(status).decode()
    if status in (ERR_INVALID_ARG, ERR_INVALID_D8, ERR_NO_PITS, ERR_UNSUPPORTED):
        err = ValueError(msg)
        err.status = status
        raise err
    if status == ERR_OOM:
        raise MemoryError(msg)
    raise PfdError(status, msg)


def ptr(a):
    """void* of a numpy array (must be C-contiguous) or of a raw device/pinned address (int)."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    if isinstance(a, C.c_void_p):
        return a
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous")
    return C.c_void_p(a.ctypes.data)


def dtype_code(dt):
    dt = np.dtype(dt)
    if dt == np.bool_:
        return DTYPES[np.dtype(np.uint8)]
    if dt not in DTYPES:
        raise TypeError(f"unsupported dtype {dt} (supported: {[str(k) for k in DTYPES]})")
    return DTYPES[dt]


class PinnedArray:
    """numpy view over page-locked host memory from pfd_host_alloc (fast, truly asynchronous H2D / D2H)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(np.atleast_1d(shape).tolist())
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        check(lib().pfd_host_alloc(max(nbytes, 1), C.byref(p)))
        self._p = p
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._p is not None:
            self.array = None
            lib().pfd_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedPool:
    """Recycled page-locked host buffers behind the numpy arrays the facade returns.

    A device->host copy into fresh pageable memory runs at ~5 GB/s (page faults + driver bounce buffers); into
    page-locked memory it runs at PCIe speed (~55 GB/s). Page-locking is expensive (~0.3 ms/MB), so blocks are kept
    and handed out again: `empty()` returns an ordinary ndarray whose memory goes back to the pool when the array
    (and every view of it) has been garbage-collected. The pool is capped (PFD_PINNED_POOL_MB, default 8192; 0
    disables it); beyond the cap, or if pinning fails, `empty()` falls back to np.empty.
    """

    MIN_BYTES = 1 << 20  # small results are not worth pinning

    def __init__(self):
        self.cap = int(os.environ.get("PFD_PINNED_POOL_MB", "8192")) << 20
        self.total = 0
        self.free = []  # (nbytes, ptr)

    def _acquire(self, nbytes):
        best = None
        for k, (cap, p) in enumerate(self.free):
            if nbytes <= cap <= 2 * nbytes and (best is None or cap < self.free[best][0]):
                best = k
        if best is not None:
            return self.free.pop(best)
        cap = (nbytes + (1 << 21) - 1) >> 21 << 21
        while self.total + cap > self.cap and self.free:  # make room by dropping idle blocks
            c, p = self.free.pop(0)
            lib().pfd_host_free(C.c_void_p(p))
            self.total -= c
        if self.total + cap > self.cap:
            return None
        p = C.c_void_p()
        if lib().pfd_host_alloc(cap, C.byref(p)) != OK:
            return None
        self.total += cap
        return cap, p.value

    def _release(self, cap, p):
        self.free.append((cap, p))

    def empty(self, n, dtype):
        import weakref

        dtype = np.dtype(dtype)
        nbytes = int(n) * dtype.itemsize
        blk = self._acquire(nbytes) if (self.cap > 0 and nbytes >= self.MIN_BYTES) else None
        if blk is None:
            return np.empty(int(n), dtype=dtype)
        cap, p = blk
        buf = (C.c_uint8 * nbytes).from_address(p)
        arr = np.frombuffer(buf, dtype=dtype, count=int(n))
        weakref.finalize(buf, self._release, cap, p)  # views keep `buf` alive through arr.base
        return arr


_pool = None


def out_array(n, dtype):
    """Result array for a device->host copy (page-locked when the pool can provide it)."""
    global _pool
    if _pool is None:
        _pool = PinnedPool()
    return _pool.empty(n, dtype)
