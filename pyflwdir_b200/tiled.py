"""Row-tiled (multi-GPU) solve of ONE D8 raster: parse + rank + upstream_area("cell") + basins() per row block.

One process per GPU (`RowBlockSolver`, exchanges over NCCL inside libpfd_b200), or -- for tests and for a single
GPU -- all row blocks solved in one process with the two exchanges emulated on the host (`solve_emulated`).
Mirrors BASELINE.json config 4 / SURVEY.md §8e; results are bit-identical to the single-GPU path.
"""
import ctypes as C

import numpy as np

from . import _lib
from .pyflwdir import _get_idxs_dtype

TILE = 64  # row blocks (except the last) must be multiples of the solver's tile height


def split_rows(nrow, nranks):
    """Row ranges [(r0, r1), ...] of `nranks` blocks: multiples of 64 rows, as even as possible, last takes the rest.
    Ranks that would get no rows are dropped (fewer blocks than requested for small rasters)."""
    ntile = (nrow + TILE - 1) // TILE
    nranks = max(1, min(nranks, ntile))
    base, extra = divmod(ntile, nranks)
    out, t0 = [], 0
    for g in range(nranks):
        t1 = t0 + base + (1 if g < extra else 0)
        out.append((t0 * TILE, min(t1 * TILE, nrow)))
        t0 = t1
    return out


def block_with_halo(d8, r0, r1):
    """(rows r0-1 .. r1 of d8 clipped to the raster, halo_top, halo_bot)"""
    halo_top = 1 if r0 > 0 else 0
    halo_bot = 1 if r1 < d8.shape[0] else 0
    return np.ascontiguousarray(d8[r0 - halo_top:r1 + halo_bot]), halo_top, halo_bot


class RowBlockSolver:
    """One row block on one GPU."""

    def __init__(self, device=0):
        self._l = _lib.lib()
        h = C.c_void_p()
        _lib.check(self._l.pfd_create(int(device), C.byref(h)))
        self._h = h
        self._dev_bufs = []
        self.nrow = self.ncol = 0
        import os
        # PFD_FUSE_PARSE=0: separate parse pass over the block instead of parsing inside phase A (identical results;
        # used by the parity tests). In the fused path idxs_ds is written by finish() / flow_all().
        if os.environ.get("PFD_FUSE_PARSE", "1") == "0":
            self._ck(self._l.pfd_set_option(self._h, b"fuse_parse", 0))

    def _ck(self, rc):
        _lib.check(rc, self._h)

    def close(self):
        if self._h is not None:
            for p in self._dev_bufs:
                self._l.pfd_dev_free(self._h, p)
            self._dev_bufs = []
            self._l.pfd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        self._ck(self._l.pfd_dev_alloc(self._h, max(int(nbytes), 16), C.byref(p)))
        self._dev_bufs.append(p)
        return p

    # -- NCCL
    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * 128)()
        _lib.check(_lib.lib().pfd_comm_unique_id(buf, 128))
        return bytes(buf)

    def comm_init(self, rank, nranks, uid):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._ck(self._l.pfd_comm_init(self._h, rank, nranks, buf))

    def barrier(self):
        self._ck(self._l.pfd_comm_barrier(self._h))

    # -- steps
    def parse(self, block, halo_top, halo_bot, glob_row0, idx_dtype=None):
        block = np.ascontiguousarray(block, dtype=np.uint8)
        self.nrow = block.shape[0] - halo_top - halo_bot
        self.ncol = block.shape[1]
        idxs = None
        code = 0
        if idx_dtype is not None:
            idxs = np.empty(self.nrow * self.ncol, dtype=idx_dtype)
            code = _lib.dtype_code(idx_dtype)
        nv, npit = C.c_int64(), C.c_int64()
        self._ck(self._l.pfd_tiled_parse(self._h, _lib.ptr(block), self.nrow, self.ncol, halo_top, halo_bot, glob_row0,
                                         _lib.ptr(idxs), code, C.byref(nv), C.byref(npit)))
        return idxs, nv.value, npit.value

    def local(self, rank, nranks, pit_id_offset):
        self._basins_dev = self.dev_alloc(self.nrow * self.ncol * 4)
        tab, n = C.c_void_p(), C.c_int64()
        self._ck(self._l.pfd_tiled_local(self._h, rank, nranks, pit_id_offset, self._basins_dev, C.byref(tab), C.byref(n)))
        return tab, n.value

    def flow_all(self, block, halo_top, halo_bot, glob_row0, idx_dtype):
        """parse + solve of this rank's row block with the exchanges over NCCL (comm_init first; a handle without a
        communicator solves a whole raster alone). Returns (idxs_ds, rank, uparea, basins, n_pits_global)."""
        block = np.ascontiguousarray(block, dtype=np.uint8)
        self.nrow = block.shape[0] - halo_top - halo_bot
        self.ncol = block.shape[1]
        n = self.nrow * self.ncol
        idxs = np.empty(n, dtype=idx_dtype)
        rank = np.empty(n, dtype=np.int32)
        upa = np.empty(n, dtype=np.int32)
        bas = np.empty(n, dtype=np.uint32)
        bas_dev = self.dev_alloc(n * 4)
        npg = C.c_int64()
        self._ck(self._l.pfd_d8_flow_all_tiled(self._h, _lib.ptr(block), self.nrow, self.ncol, halo_top, halo_bot,
                                               glob_row0, _lib.ptr(idxs), _lib.dtype_code(idx_dtype), _lib.ptr(rank),
                                               _lib.ptr(upa), bas_dev, None, C.byref(npg)))
        self._ck(self._l.pfd_memcpy(self._h, _lib.ptr(bas), bas_dev, n * 4))
        shape = (self.nrow, self.ncol)
        return idxs, rank.reshape(shape), upa.reshape(shape), bas.reshape(shape), npg.value

    def read_table(self, tab, n):
        out = np.empty(n, dtype=np.uint32)
        if n:
            self._ck(self._l.pfd_memcpy(self._h, _lib.ptr(out), tab, n * 4))
        return out

    def write_table(self, tab, values):
        values = np.ascontiguousarray(values, dtype=np.uint32)
        if values.size:
            self._ck(self._l.pfd_memcpy(self._h, tab, _lib.ptr(values), values.size * 4))

    def finish(self):
        n = self.nrow * self.ncol
        rank = np.empty(n, dtype=np.int32)
        upa = np.empty(n, dtype=np.int32)
        bas = np.empty(n, dtype=np.uint32)
        self._ck(self._l.pfd_tiled_finish(self._h, _lib.ptr(rank), _lib.ptr(upa), self._basins_dev))
        self._ck(self._l.pfd_memcpy(self._h, _lib.ptr(bas), self._basins_dev, n * 4))
        shape = (self.nrow, self.ncol)
        return rank.reshape(shape), upa.reshape(shape), bas.reshape(shape)


def solve_emulated(d8, nranks, device=0):
    """All row blocks of `d8` on ONE GPU, one after the other, with the two exchanges done on the host.
    Returns dict(idxs_ds, rank, uparea, basins, n_pits) for the whole raster."""
    d8 = np.ascontiguousarray(d8, dtype=np.uint8)
    nrow, ncol = d8.shape
    dtype = _get_idxs_dtype(nrow * ncol)
    blocks = split_rows(nrow, nranks)
    R = len(blocks)
    solvers = [RowBlockSolver(device) for _ in blocks]
    try:
        idxs, npits = [], []
        for s, (r0, r1) in zip(solvers, blocks):
            blk, ht, hb = block_with_halo(d8, r0, r1)
            ids, _, npit = s.parse(blk, ht, hb, r0, dtype)
            idxs.append(ids)
            npits.append(npit)
        offsets = np.concatenate([[0], np.cumsum(npits)[:-1]]).astype(np.int64)  # exchange #1 (all-gather)
        tabs = [s.local(g, R, int(offsets[g])) for g, s in enumerate(solvers)]
        if R > 1:  # exchange #2 (all-reduce, uint32 sum)
            total = np.zeros(tabs[0][1], dtype=np.uint32)
            for s, (tab, n) in zip(solvers, tabs):
                total += s.read_table(tab, n)
            for s, (tab, n) in zip(solvers, tabs):
                s.write_table(tab, total)
        outs = [s.finish() for s in solvers]
        return dict(idxs_ds=np.concatenate(idxs), rank=np.concatenate([o[0] for o in outs]),
                    uparea=np.concatenate([o[1] for o in outs]), basins=np.concatenate([o[2] for o in outs]),
                    n_pits=int(np.sum(npits)), blocks=blocks)
    finally:
        for s in solvers:
            s.close()
