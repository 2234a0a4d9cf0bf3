"""Shared helpers of the module-level L1/L2 mirrors (core_d8 / core / streams / basins / dem).

The reference's free functions take a flat `idxs_ds` (and usually `seq`) and no raster shape; the device graph needs
the number of columns, so it is inferred from the links (every D8 link is 0, +-1, +-ncol or +-ncol+-1) unless the
caller passes `shape=` / `ncol=` (an extension of the reference signatures)."""
import numpy as np

from . import _device, _lib


def infer_ncol(idxs_ds):
    """Number of raster columns implied by a D8 downstream-index array, or ValueError when it is ambiguous."""
    idxs_ds = np.asarray(idxs_ds)
    n = idxs_ds.size
    i = np.arange(n, dtype=np.int64)
    ds = idxs_ds.astype(np.int64)  # uint64: the missing value 2**64 - 1 wraps to -1
    mv = np.int64(-1) if idxs_ds.dtype.kind == "i" or idxs_ds.dtype.itemsize == 8 else np.int64(np.iinfo(idxs_ds.dtype).max)
    link = (ds != mv) & (ds != i)
    if idxs_ds.dtype == np.uint32:
        link &= idxs_ds != np.uint32(0xFFFFFFFF)
    d = np.abs(ds[link] - i[link])
    big = np.unique(d[d > 1])
    cands = {n} if big.size == 0 else set()
    for v in big.tolist():
        cands.update((v - 1, v, v + 1))
    good = []
    for nc in sorted(c for c in cands if c >= 1 and n % c == 0):
        dr = ds[link] // nc - i[link] // nc
        dc = ds[link] % nc - i[link] % nc
        if np.all(np.abs(dr) <= 1) and np.all(np.abs(dc) <= 1):
            good.append(nc)
    if len(good) != 1:
        raise ValueError("cannot infer the raster width from idxs_ds; pass shape=(nrow, ncol)")
    return good[0]


def resolve_shape(idxs_ds, shape=None, ncol=None):
    n = np.asarray(idxs_ds).size
    if shape is not None:
        shape = tuple(int(v) for v in shape)
        if shape[0] * shape[1] != n:
            raise ValueError(f"shape {shape} does not match size {n}")
        return shape
    if ncol is None:
        ncol = infer_ncol(idxs_ds)
    return (n // int(ncol), int(ncol))


def graph(idxs_ds, shape=None, ncol=None, device=0):
    """DeviceGraph loaded from a downstream-index array."""
    shape = resolve_shape(idxs_ds, shape, ncol)
    g = _device.DeviceGraph(device)
    idxs_ds = np.ascontiguousarray(idxs_ds)
    if idxs_ds.dtype == np.dtype(np.intp) and idxs_ds.dtype not in (np.dtype(np.int32), np.dtype(np.int64)):  # pragma: no cover
        idxs_ds = idxs_ds.astype(np.int64)
    g.load_idxs_ds(idxs_ds, shape)
    return g


def check_seq(g, seq, what, order_sensitive=False):
    """The sweeps run over the device's own "walk" sequence (what `core.idxs_seq` / `FlwdirRaster.idxs_seq` return). A
    caller-supplied `seq` must cover the same cells; when the RESULT depends on the order inside `seq` (float sums over the
    upstream cells, label / segment numbering in sequence order) a different order -- e.g. the `np.argsort(rank)` sequence of
    the reference's test fixtures -- would give different values in the reference, so it is refused instead of silently
    answering for the walk order (first row)."""
    if seq is None:
        return
    seq = np.asarray(seq)
    nn = g.order()[0]
    if seq.size != nn:
        raise NotImplementedError(
            f"{what}: `seq` holds {seq.size} cells but {nn} cells drain to a pit; sweeps over a "
            "partial sequence are outside the accelerated hot path")
    if order_sensitive:
        walk = g.fetch(_lib.ARR_SEQ, np.int64 if seq.dtype.itemsize == 8 else np.int32)
        if not np.array_equal(walk, seq.ravel()):
            raise NotImplementedError(
                f"{what}: the result depends on the order of the cells inside `seq`, and the device sweeps the \"walk\" order "
                "(core.idxs_seq); pass that sequence (FlwdirRaster.idxs_seq), or integer data for accumulations")
