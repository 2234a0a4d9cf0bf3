"""Region (label raster) post-processing; mirrors /root/reference/pyflwdir/regions.py (region_sum :16-32, region_area :35-55,
region_slices :58-86, region_bounds :89-129, region_outlets :132-163). Slices / bounds / outlets run on the GPU; the two
label statistics are host-side numpy like the reference's scipy.ndimage calls (float64 sums in raster order)."""
import numpy as np

from . import _device, _functional
from . import gis_utils as gis

__all__ = ["region_bounds", "region_slices", "region_outlets", "region_sum", "region_area"]


def region_sum(data, regions):
    """Returns the sum of values in `data` for each unique label (> 0) in `regions` -> (labels, float64 sums).
    Same accumulation as scipy.ndimage.sum(data, regions, index=labels): float64, cells added in raster order."""
    data, regions = np.asarray(data), np.asarray(regions)
    if data.shape != regions.shape:
        raise ValueError("input and labels must have the same shape")
    sel = regions > 0
    lbs, inv = np.unique(regions[sel], return_inverse=True)
    sums = np.bincount(inv.ravel(), weights=data[sel].astype(np.float64, copy=False), minlength=lbs.size)
    return lbs, sums


def region_area(regions, transform=gis.IDENTITY, latlon=False):
    """Returns the area [m2] for each unique label in `regions` -> (labels, areas)"""
    regions = np.asarray(regions)
    area = gis.area_grid(transform=transform, shape=regions.shape, latlon=latlon)
    return region_sum(area, regions)


def region_dissolve(regions, labels=None, idxs=None, transform=gis.IDENTITY, latlon=False, **kwargs):
    """Out of scope (regions post-processing built on gis_utils.spread2d, a priority-queue spread)."""
    raise NotImplementedError("region_dissolve needs gis_utils.spread2d, which is outside the D8 hot path that pyflwdir_b200 "
                              "accelerates; use the reference on the label raster")


def region_slices(regions, device=0):
    """Returns slices for each unique label in `regions` -> (labels, list of (row slice, column slice))"""
    regions = np.asarray(regions)
    if regions.ndim != 2:
        raise ValueError('The "regions" array should be two dimensional')
    g = _device.DeviceGraph(device)
    g.parse_d8(np.zeros(regions.shape, dtype=np.uint8))  # a raster of pits: only the shape matters here
    lbs, sl = g.region_slices(regions)
    if lbs.size == 0:
        raise ValueError("No regions found in data")
    return lbs, [(slice(r0, r1, None), slice(c0, c1, None)) for r0, r1, c0, c1 in sl.tolist()]


def region_bounds(regions, transform=gis.IDENTITY, device=0):
    """Returns the bounding box of each unique label in `regions` -> (labels, [xmin, ymin, xmax, ymax] per label, total)"""
    lbs, slices = region_slices(regions, device=device)
    xres, yres = transform[0], transform[4]
    lons, lats = gis.affine_to_coords(transform, np.shape(regions))
    iy = np.array([0, -1])
    ix = iy.copy()
    if yres < 0:
        iy = iy[::-1]
    if xres < 0:
        ix = ix[::-1]
    dx, dy = np.abs(xres) / 2, np.abs(yres) / 2
    bboxs = []
    for yslice, xslice in slices:
        xmin, xmax = lons[xslice][ix]
        ymin, ymax = lats[yslice][iy]
        bboxs.append([xmin - dx, ymin - dy, xmax + dx, ymax + dy])
    bboxs = np.asarray(bboxs)
    total_bbox = np.hstack([bboxs[:, :2].min(axis=0), bboxs[:, 2:].max(axis=0)])
    return lbs, bboxs, total_bbox


def region_outlets(regions, idxs_ds, seq, shape=None, ncol=None):
    """Returns the linear index of the outlet cell in `regions` -> (labels, outlet indices), sorted by label"""
    regions = np.asarray(regions)
    if shape is None and regions.ndim == 2:
        shape = regions.shape
    g = _functional.graph(idxs_ds, shape, ncol)
    _functional.check_seq(g, seq, "region_outlets", order_sensitive=True)  # ties between outlets follow the sequence
    return g.region_outlets(regions.ravel(), np.asarray(idxs_ds).dtype)
