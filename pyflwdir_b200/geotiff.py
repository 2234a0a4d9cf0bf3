"""Single-band GeoTIFF in and out without rasterio / GDAL (the on-disk format either side of the hot
path). The reference has no reader of its own -- its examples open `examples/rhine_d8.tif` / `rhine_elv0.tif` with rasterio
and hand `from_array` the array, the affine transform and `crs.is_geographic` (examples/*.ipynb, README quick start); this
module returns exactly those three things (+ nodata) so that the same two lines work without that dependency:

    d8, transform, latlon, nodata = geotiff.read("rhine_d8.tif")
    flw = pyflwdir_b200.from_array(d8, ftype="d8", transform=transform, latlon=latlon)

Scope: classic (32-bit offset) TIFF, little or big endian, one sample per pixel, strips or tiles, 8/16/32/64-bit integers
and 32/64-bit floats, compression none / LZW (5) / deflate (8, 32946) / PackBits (32773), predictor 1, 2 (horizontal differencing)
or 3 (floating point),
north-up georeferencing by ModelPixelScale + ModelTiepoint or ModelTransformation, GDAL_NODATA. Anything else raises
NotImplementedError naming the tag. (LZW is decoded in pure Python, about 1 MB/s: fine for the example rasters, slow for
gigabyte files -- deflate goes through zlib.) `write` produces an uncompressed or deflate-compressed striped file that `read`,
rasterio and GDAL open."""
import struct
import zlib

import numpy as np

from .gis_utils import Affine

__all__ = ["read", "write"]

_TYPES = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 6: "b", 8: "h", 9: "i", 11: "f", 12: "d", 16: "Q"}
_SIZES = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8, 16: 8}


def _ifd(buf, bo):
    magic, off = struct.unpack(bo + "HI", buf[2:8])
    if magic != 42:
        raise NotImplementedError("only classic TIFF (magic 42) is supported, not BigTIFF")
    (n,) = struct.unpack(bo + "H", buf[off:off + 2])
    tags = {}
    for i in range(n):
        e = off + 2 + 12 * i
        tag, typ, cnt = struct.unpack(bo + "HHI", buf[e:e + 8])
        size = _SIZES.get(typ, 1) * cnt
        pos = e + 8 if size <= 4 else struct.unpack(bo + "I", buf[e + 8:e + 12])[0]
        raw = buf[pos:pos + size]
        if typ == 2:
            val = raw.split(b"\0")[0].decode("ascii", "replace")
        elif typ in (5, 10):
            v = struct.unpack(bo + ("II" if typ == 5 else "ii") * cnt, raw)
            val = tuple(v[2 * k] / v[2 * k + 1] for k in range(cnt))
        elif typ == 7:
            val = raw
        else:
            val = struct.unpack(bo + _TYPES[typ] * cnt, raw)
        tags[tag] = val
    return tags


def _first(tags, tag, default=None):
    v = tags.get(tag)
    if v is None:
        return default
    return v[0] if isinstance(v, tuple) else v


def _unpackbits(data, size):
    out = bytearray()
    i = 0
    while i < len(data) and len(out) < size:
        n = data[i]
        i += 1
        if n < 128:
            out += data[i:i + n + 1]
            i += n + 1
        elif n > 128:
            out += data[i:i + 1] * (257 - n)
            i += 1
    return bytes(out[:size])


def _unlzw(data, size):
    """TIFF LZW (compression 5): MSB-first codes of 9..12 bits, 256 = clear, 257 = end of information, early change."""
    out = bytearray()
    table = [bytes((i,)) for i in range(256)] + [b"", b""]
    nbits, bitbuf, bitcnt, prev = 9, 0, 0, None
    for byte in data:
        bitbuf = (bitbuf << 8) | byte
        bitcnt += 8
        while bitcnt >= nbits:
            bitcnt -= nbits
            code = (bitbuf >> bitcnt) & ((1 << nbits) - 1)
            if code == 257:
                return bytes(out[:size])
            if code == 256:
                table = table[:258]
                nbits, prev = 9, None
                continue
            if prev is None:
                entry = table[code]
            elif code < len(table):
                entry = table[code]
                table.append(prev + entry[:1])
            else:
                entry = prev + prev[:1]
                table.append(entry)
            out += entry
            prev = entry
            if len(table) >= (1 << nbits) - 1 and nbits < 12:
                nbits += 1
            if len(out) >= size:
                return bytes(out[:size])
    return bytes(out[:size])


def read(path):
    """Returns (array 2D, transform Affine, latlon bool, nodata or None)."""
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:2] == b"II":
        bo = "<"
    elif buf[:2] == b"MM":
        bo = ">"
    else:
        raise ValueError(f"{path}: not a TIFF file")
    t = _ifd(buf, bo)
    ncol, nrow = _first(t, 256), _first(t, 257)
    if _first(t, 277, 1) != 1:
        raise NotImplementedError("SamplesPerPixel (277) != 1: only single-band rasters are supported")
    bits, fmt = _first(t, 258, 1), _first(t, 339, 1)
    kind = {1: "u", 2: "i", 3: "f"}.get(fmt)
    if kind is None or bits not in (8, 16, 32, 64) or (kind == "f" and bits < 32):
        raise NotImplementedError(f"BitsPerSample (258) = {bits}, SampleFormat (339) = {fmt}")
    dtype = np.dtype(f"{bo}{kind}{bits // 8}")
    comp, pred = _first(t, 259, 1), _first(t, 317, 1)
    if comp not in (1, 5, 8, 32946, 32773):
        raise NotImplementedError(f"Compression (259) = {comp}: only none / LZW / deflate / PackBits are supported")
    if pred not in (1, 2, 3) or (pred == 2 and kind == "f") or (pred == 3 and kind != "f"):
        raise NotImplementedError(f"Predictor (317) = {pred}")
    tiled = 322 in t
    if tiled:
        bw, bh = _first(t, 322), _first(t, 323)
        offs, cnts = t[324], t[325]
    else:
        bw, bh = ncol, min(_first(t, 278, nrow), nrow)
        offs, cnts = t[273], t[279]
    nbx, nby = -(-ncol // bw), -(-nrow // bh)
    out = np.empty((nrow, ncol), dtype=dtype.newbyteorder("="))
    for by in range(nby):
        for bx in range(nbx):
            k = by * nbx + bx
            raw = buf[offs[k]:offs[k] + cnts[k]]
            rows = bh if tiled else min(bh, nrow - by * bh)
            size = rows * bw * dtype.itemsize
            if comp in (8, 32946):
                raw = zlib.decompress(raw)
            elif comp == 32773:
                raw = _unpackbits(raw, size)
            elif comp == 5:
                raw = _unlzw(raw, size)
            if pred == 3:  # floating-point predictor: bytes of a row differenced, then the byte planes stored most significant first
                b = np.frombuffer(raw[:size], dtype=np.uint8).reshape(rows, bw * dtype.itemsize)
                b = np.cumsum(b, axis=1, dtype=np.uint8).reshape(rows, dtype.itemsize, bw)
                blk = np.ascontiguousarray(b.transpose(0, 2, 1)).view(np.dtype(f">{kind}{bits // 8}")).reshape(rows, bw).astype(out.dtype)
            else:
                blk = np.frombuffer(raw[:size], dtype=dtype).reshape(rows, bw).astype(out.dtype)
            if pred == 2:
                blk = np.cumsum(blk, axis=1, dtype=out.dtype)
            r0, c0 = by * bh, bx * bw
            out[r0:r0 + rows, c0:c0 + bw] = blk[:min(rows, nrow - r0), :min(bw, ncol - c0)]
    if 34264 in t:  # ModelTransformation: row-major 4 x 4
        m = t[34264]
        transform = Affine(m[0], m[1], m[3], m[4], m[5], m[7])
    elif 33550 in t and 33922 in t:
        sx, sy = t[33550][0], t[33550][1]
        i, j, _, x, y, _ = t[33922][:6]
        transform = Affine(sx, 0.0, x - i * sx, 0.0, -sy, y + j * sy)
    else:
        transform = Affine(1.0, 0.0, 0.0, 0.0, -1.0, 0.0)
    latlon = False
    if 34735 in t:  # GeoKeyDirectory: key 1024 GTModelType, 2 = geographic
        g = t[34735]
        for k in range(4, len(g) - 3, 4):
            if g[k] == 1024 and g[k + 1] == 0:
                latlon = g[k + 3] == 2
    nodata = None
    if 42113 in t:
        try:
            nodata = float(t[42113])
            if kind != "f" and nodata == int(nodata):
                nodata = int(nodata)
        except ValueError:
            nodata = None
    return out, transform, latlon, nodata


def write(path, data, transform=Affine(1.0, 0.0, 0.0, 0.0, -1.0, 0.0), latlon=False, nodata=None, compress=True, epsg=None):
    """Writes a 2D array as a striped little-endian GeoTIFF (deflate when `compress`)."""
    a = np.ascontiguousarray(data)
    if a.ndim != 2:
        raise ValueError("data should be a 2D array")
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    if a.dtype.kind not in "uif" or a.dtype.itemsize not in (1, 2, 4, 8) or (a.dtype.kind == "f" and a.dtype.itemsize < 4):
        raise NotImplementedError(f"dtype {a.dtype}")
    a = a.astype(a.dtype.newbyteorder("<"), copy=False)
    nrow, ncol = a.shape
    rps = max(1, min(nrow, (1 << 16) // max(1, ncol * a.dtype.itemsize)))
    strips = [a[r:r + rps].tobytes() for r in range(0, nrow, rps)]
    if compress:
        strips = [zlib.compress(s, 6) for s in strips]
    fmt = {"u": 1, "i": 2, "f": 3}[a.dtype.kind]
    if epsg is None:
        epsg = 4326 if latlon else None
    geokeys = [1, 1, 0, 0, 1024, 0, 1, 2 if latlon else 1, 1025, 0, 1, 1]
    if epsg is not None:
        geokeys += [2048 if latlon else 3072, 0, 1, int(epsg)]
    geokeys[3] = (len(geokeys) - 4) // 4
    entries = []  # (tag, type, count, payload bytes)

    def add(tag, typ, values):
        if typ == 2:
            payload = values.encode("ascii") + b"\0"
            cnt = len(payload)
        else:
            payload = struct.pack("<" + _TYPES[typ] * len(values), *values)
            cnt = len(values)
        entries.append((tag, typ, cnt, payload))

    add(256, 4, [ncol])
    add(257, 4, [nrow])
    add(258, 3, [a.dtype.itemsize * 8])
    add(259, 3, [8 if compress else 1])
    add(262, 3, [1])
    add(273, 4, [0] * len(strips))  # patched below
    add(277, 3, [1])
    add(278, 4, [rps])
    add(279, 4, [len(s) for s in strips])
    add(284, 3, [1])
    add(339, 3, [fmt])
    add(33550, 12, [float(transform[0]), float(-transform[4]), 0.0])
    add(33922, 12, [0.0, 0.0, 0.0, float(transform[2]), float(transform[5]), 0.0])
    add(34735, 3, geokeys)
    if nodata is not None:
        add(42113, 2, repr(float(nodata)) if a.dtype.kind == "f" else str(int(nodata)))
    entries.sort(key=lambda e: e[0])
    ifd_off = 8
    ifd_size = 2 + 12 * len(entries) + 4
    extra_off = ifd_off + ifd_size
    extras = b""
    where = {}
    for tag, typ, cnt, payload in entries:
        if len(payload) > 4:
            if len(extras) % 2:
                extras += b"\0"
            where[tag] = extra_off + len(extras)
            extras += payload
    data_off = extra_off + len(extras)
    data_off += data_off % 2
    strip_offs, pos = [], data_off
    for s in strips:
        strip_offs.append(pos)
        pos += len(s)
    with open(path, "wb") as f:
        f.write(b"II" + struct.pack("<HI", 42, ifd_off))
        f.write(struct.pack("<H", len(entries)))
        patched = {}
        for tag, typ, cnt, payload in entries:
            if tag == 273:
                payload = struct.pack("<" + "I" * len(strip_offs), *strip_offs)
                patched[tag] = payload
            if len(payload) <= 4:
                f.write(struct.pack("<HHI", tag, typ, cnt) + payload.ljust(4, b"\0"))
            else:
                f.write(struct.pack("<HHII", tag, typ, cnt, where[tag]))
        f.write(struct.pack("<I", 0))
        blob = bytearray(extras)
        if 273 in patched and 273 in where:
            o = where[273] - extra_off
            blob[o:o + len(patched[273])] = patched[273]
        f.write(bytes(blob))
        f.write(b"\0" * (data_off - extra_off - len(extras)))
        for s in strips:
            f.write(s)
