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

    # -- row-block sweeps of the order-sensitive outputs (Strahler, accuflux, HAND)
    KINDS = {"strahler": 0, "accuflux": 1, "hand": 2}

    def _sweep_args(self, kind, data, drain, nodata):
        k = self.KINDS[kind]
        n = self.nrow * self.ncol
        code, dref, mref = 0, None, None
        nodata_f, nodata_i, is_int = 0.0, 0, 0
        if k == 1:
            dref = np.ascontiguousarray(data).reshape(-1)
            code = _lib.dtype_code(dref.dtype)
            from ._device import nodata_args  # numba's typing of `x != nodata` (int nodata: int64 compare)

            nf, ni, nis = nodata_args(nodata)
            nodata_f, nodata_i, is_int = nf.value, ni.value, nis.value
        elif k == 2:
            dref = np.ascontiguousarray(data).reshape(-1)
            if dref.dtype not in (np.float32, np.float64):
                dref = dref.astype(np.float64)
            code = _lib.dtype_code(dref.dtype)
            mref = np.ascontiguousarray(drain).reshape(-1)
            mref = mref.view(np.uint8) if mref.dtype == np.bool_ else (mref == 1).view(np.uint8)
        if dref is not None and dref.size != n:
            raise ValueError('"data" size does not match.')
        out_dtype = np.uint8 if k == 0 else (dref.dtype if k == 1 else np.float64)
        return k, dref, code, mref, nodata_f, nodata_i, is_int, out_dtype

    def sweep_begin(self, kind, data=None, drain=None, nodata=-9999):
        k, dref, code, mref, nf, ni, is_int, self._sweep_dtype = self._sweep_args(kind, data, drain, nodata)
        self._ck(self._l.pfd_sweep_tiled_begin(self._h, k, _lib.ptr(dref), code, _lib.ptr(mref), C.c_double(nf), ni, is_int))

    def sweep_round(self):
        n = C.c_int64()
        self._ck(self._l.pfd_sweep_tiled_round(self._h, C.byref(n)))
        return n.value

    def sweep_edges(self, which):
        """packed [dir | done | value | aux] record of my first (0) / last (1) row, as host bytes"""
        p, nb = C.c_void_p(), C.c_int64()
        self._ck(self._l.pfd_sweep_tiled_edges(self._h, which, C.byref(p), C.byref(nb)))
        buf = np.empty(nb.value, np.uint8)
        self._ck(self._l.pfd_memcpy(self._h, _lib.ptr(buf), p, nb.value))
        return buf

    def sweep_halo(self, which, record):
        self._ck(self._l.pfd_sweep_tiled_halo(self._h, which, _lib.ptr(np.ascontiguousarray(record))))

    def sweep_end(self):
        out = np.empty(self.nrow * self.ncol, self._sweep_dtype)
        res = C.c_int64()
        self._ck(self._l.pfd_sweep_tiled_end(self._h, _lib.ptr(out), C.byref(res)))
        return out.reshape(self.nrow, self.ncol), res.value

    def sweep(self, kind, data=None, drain=None, nodata=-9999):
        """Strahler order / accuflux / HAND of this rank's row block with the halo rounds over NCCL (after flow_all or
        parse on a handle with a communicator). Returns (own rows of the result, number of rounds)."""
        k, dref, code, mref, nf, ni, is_int, out_dtype = self._sweep_args(kind, data, drain, nodata)
        out = np.empty(self.nrow * self.ncol, out_dtype)
        rounds = C.c_int64()
        self._ck(self._l.pfd_sweep_tiled(self._h, k, _lib.ptr(dref), code, _lib.ptr(mref), C.c_double(nf), ni, is_int,
                                         _lib.ptr(out), C.byref(rounds)))
        return out.reshape(self.nrow, self.ncol), rounds.value

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


def sweep_emulated(d8, nranks, kind, data=None, drain=None, nodata=-9999, device=0):
    """Strahler order / accuflux / HAND of `d8` over `nranks` row blocks on ONE GPU, the halo exchange of every round done
    on the host: the same step functions the NCCL path (RowBlockSolver.sweep) drives. Returns (result, rounds)."""
    d8 = np.ascontiguousarray(d8, dtype=np.uint8)
    nrow, ncol = d8.shape
    blocks = split_rows(nrow, nranks)
    R = len(blocks)
    solvers = [RowBlockSolver(device) for _ in blocks]

    def part(a, r0, r1):
        return None if a is None else np.ascontiguousarray(np.asarray(a).reshape(nrow, ncol)[r0:r1])

    def swap():
        recs = [(s.sweep_edges(0) if g > 0 else None, s.sweep_edges(1) if g < R - 1 else None) for g, s in enumerate(solvers)]
        for g, s in enumerate(solvers):
            if g > 0:
                s.sweep_halo(0, recs[g - 1][1])
            if g < R - 1:
                s.sweep_halo(1, recs[g + 1][0])

    try:
        for s, (r0, r1) in zip(solvers, blocks):
            blk, ht, hb = block_with_halo(d8, r0, r1)
            s.parse(blk, ht, hb, r0, None)
            s.sweep_begin(kind, part(data, r0, r1), part(drain, r0, r1), nodata)
        swap()
        rounds = 0
        while True:
            newly = sum(s.sweep_round() for s in solvers)
            swap()
            rounds += 1
            if newly == 0:
                break
        outs = [s.sweep_end() for s in solvers]
        return np.concatenate([o[0] for o in outs]), rounds, sum(o[1] for o in outs)
    finally:
        for s in solvers:
            s.close()
