"""Host-side GeoTIFF IO (pyflwdir_b200/geotiff.py): round trips, an independent decoder (PIL) on the
files we write, and -- where the reference is mounted -- its example rasters against the committed golden input."""
import os

import numpy as np
import pytest

import _cases as cs
from pyflwdir_b200 import geotiff
from pyflwdir_b200.gis_utils import Affine

EXAMPLES = "/root/reference/examples"


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.int32, np.uint32, np.int64, np.float32, np.float64])
def test_roundtrip(tmp_path, dtype):
    rng = np.random.default_rng(0)
    tr = Affine(0.25, 0.0, -3.5, 0.0, -0.25, 60.0)
    for shape in ((1, 1), (7, 13), (300, 421), (2000, 90)):
        x = (rng.random(shape) * 200 - 50).astype(dtype)
        for compress in (True, False):
            fn = str(tmp_path / "x.tif")
            nodata = 255 if np.dtype(dtype).kind == "u" else -9999
            geotiff.write(fn, x, transform=tr, latlon=True, nodata=nodata, compress=compress)
            y, t2, latlon, nd = geotiff.read(fn)
            assert y.dtype == x.dtype and np.array_equal(x, y)
            assert tuple(t2)[:6] == tuple(tr)[:6] and latlon is True and nd == nodata
            if np.dtype(dtype) in (np.dtype(np.uint8), np.dtype(np.int16), np.dtype(np.int32), np.dtype(np.float32)):
                from PIL import Image  # an independent TIFF decoder reads the same pixels

                assert np.array_equal(np.array(Image.open(fn)), x)


def test_projected_and_errors(tmp_path):
    fn = str(tmp_path / "p.tif")
    geotiff.write(fn, np.arange(6, dtype=np.int32).reshape(2, 3), transform=Affine(30.0, 0.0, 1000.0, 0.0, -30.0, 5000.0), epsg=32631)
    a, t, latlon, nd = geotiff.read(fn)
    assert latlon is False and nd is None and tuple(t)[:6] == (30.0, 0.0, 1000.0, 0.0, -30.0, 5000.0)
    with open(fn, "wb") as f:
        f.write(b"not a tiff")
    with pytest.raises(ValueError, match="not a TIFF"):
        geotiff.read(fn)
    with pytest.raises(ValueError, match="2D"):
        geotiff.write(fn, np.zeros(3))


@pytest.mark.skipif(not os.path.exists(os.path.join(EXAMPLES, "rhine_d8.tif")), reason="/root/reference not mounted")
def test_reference_examples():
    d8, t, latlon, nodata = geotiff.read(os.path.join(EXAMPLES, "rhine_d8.tif"))
    assert np.array_equal(d8, cs.case_d8("rhine"))  # the committed golden input was decoded from this file with PIL
    assert tuple(t)[:6] == cs.RHINE_TRANSFORM and latlon is True
    elv, t2, _, nd = geotiff.read(os.path.join(EXAMPLES, "rhine_elv0.tif"))
    assert elv.dtype == np.float32 and elv.shape == d8.shape and nd == -9999.0 and tuple(t2)[:6] == cs.RHINE_TRANSFORM


@pytest.mark.parametrize("compression", ["tiff_lzw", "tiff_adobe_deflate", "packbits", "raw"])
def test_reads_files_written_by_an_independent_encoder(tmp_path, compression):
    """PIL writes, geotiff.read decodes: LZW (the usual compression of distributed GeoTIFFs), deflate, PackBits, none"""
    from PIL import Image

    rng = np.random.default_rng(1)
    fn = str(tmp_path / "p.tif")
    for dtype in (np.uint8, np.int32, np.float32):
        for shape in ((1, 1), (7, 13), (120, 201), (40, 700)):
            for x in ((rng.random(shape) * 200).astype(dtype), (np.add.outer(np.arange(shape[0]), np.arange(shape[1])) // 7).astype(dtype)):
                Image.fromarray(x).save(fn, compression=compression)
                y = geotiff.read(fn)[0]
                assert y.dtype == x.dtype and np.array_equal(x, y), (dtype, shape, compression)


def test_floating_point_predictor(tmp_path):
    """Predictor 3 (TIFF Technical Note 3), encoded here straight from its definition: per row the bytes of the samples are
    regrouped into planes, most significant byte first, and the resulting byte string is differenced"""
    import struct
    import zlib

    rng = np.random.default_rng(2)
    x = (rng.random((37, 53)) * 1000 - 300).astype(np.float32)
    rows = []
    for r in range(x.shape[0]):
        planes = x[r].astype(">f4").view(np.uint8).reshape(-1, 4).T.reshape(-1).astype(np.int16)  # plane 0 = most significant bytes
        rows.append((np.diff(planes, prepend=0) & 0xFF).astype(np.uint8).tobytes())
    data = zlib.compress(b"".join(rows))
    fn = str(tmp_path / "f.tif")
    tags = [(256, 4, x.shape[1]), (257, 4, x.shape[0]), (258, 3, 32), (259, 3, 8), (262, 3, 1), (273, 4, 8 + 2 + 12 * 11 + 4),
            (277, 3, 1), (278, 4, x.shape[0]), (279, 4, len(data)), (317, 3, 3), (339, 3, 3)]
    with open(fn, "wb") as f:
        f.write(b"II" + struct.pack("<HI", 42, 8) + struct.pack("<H", len(tags)))
        for tag, typ, val in tags:
            f.write(struct.pack("<HHI", tag, typ, 1) + (struct.pack("<I", val) if typ == 4 else struct.pack("<HH", val, 0)))
        f.write(struct.pack("<I", 0) + data)
    y = geotiff.read(fn)[0]
    assert y.dtype == np.float32 and np.array_equal(x, y)
