#!/usr/bin/env python
"""bench.py -- headline benchmark of the D8 flow-network hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size 32768] [--impl ours|reference]

One "step" = the whole hot path on one synthetic D8 raster (SURVEY.md §8d generator, "rough" Perlin fBm with
K = log2(n)-2 octaves, steepest descent, no depression filling):
    parse (d8 -> idxs_ds) + order/rank + upstream_area(cell) + basins()      == pfd_d8_flow_all
Outputs of a step: idxs_ds int32, rank int32, upstream area int32, basins uint32 (all N cells each).

  value : Mcells/s with the d8 raster and all outputs RESIDENT IN HBM, timed with CUDA events on the handle's
          stream around exactly K steps (max over ranks).
  e2e   : the same C-ABI call with PINNED HOST buffers: H2D of the raster and D2H of the four outputs are inside
          the timed region.
  roofline     : the dominant kernel of the step (largest share of device time), algorithmic bytes / event time.
  cpu_baseline : the reference's own numba CPU path (installed by oracle/make_ref.sh into the git-ignored oracle/_ref,
          kind "reference"; else the C oracle port, kind "port") on a --cpu-size^2 sample of the same generator.

N > 1 (launched by torchrun, one process per GPU): ONE raster of (N*size) x size cells is row-tiled across the GPUs,
each rank owning `size` rows (+ one halo row per neighbour): weak scaling with the real exchange step of the path
(pit-count all-gather + one NCCL all-reduce of the boundary tables per step, pfd_d8_flow_all_tiled). torch.distributed
(gloo) is used only to hand out the NCCL unique id and for the barrier / max-reduce of the timings.

--impl reference: times the reference's CPU implementation of the path on the host cores -- the UNMODIFIED numba
reference from oracle/_ref (oracle/make_ref.sh) when present, else the C oracle port -- same metric / config, each step
one pass over a --cpu-size^2 sample of the same generator (the numba kernels are single-threaded).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mcells/s D8 parse+rank+accuflux+basins"
UNIT = "Mcells/s"
# algorithmic bytes per cell (DESIGN.md "Kernels"): compulsory input read + output write
# tile solver kernels: phase A reads dir (1) and writes loc + cnt (8); phase C reads dir + loc + cnt (9) and writes
# rank + basins + uparea (12); the whole fused step has a 17 B/cell lower bound (d8 in, four int32 outputs)
ALG_BYTES = {"parse": 5.0, "bfs": 12.0, "sweep": 8.0, "tile_a": 9.0, "tile_b": 0.0, "tile_c": 21.0}
# fused-parse path (device-resident buffers): phase A reads the raw D8 codes (1) and writes dir (1) + loc + cnt (8);
# phase C reads dir + loc + cnt (9) and writes idxs_ds (4; 8 when int64) + rank + basins + uparea (12)
ALG_BYTES_FUSED = {"tile_a": 10.0, "tile_c": 25.0}
STEP_BYTES_FUSED, STEP_BYTES_UNFUSED = 17.0, 29.0
KERNEL_NAMES = {"parse": "parse_kernel", "bfs": "bfs_kernel", "sweep": "sweep_kernel<AccuUpOp<int>>",
                "tile_a": "tile_phase_a_kernel", "tile_b": "slots_round_kernel", "tile_c": "tile_phase_c_kernel"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.device)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
            self.f.close()
            sm, mx, reasons = [], [], set()
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            if sm:
                out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                       "samples": len(sm)}
            os.unlink(self.path)
        except Exception:
            pass
        return out


def dist_setup(n_gpus):
    """torch.distributed (gloo) purely for barrier / max-reduce across the per-GPU processes."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None, 0, 1, 0
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) off it
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group(backend="gloo")
    return dist, dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", "0"))


def barrier(dist):
    if dist is not None:
        dist.barrier()


def reduce_max(dist, x):
    if dist is None:
        return x
    import torch

    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(dist, x):
    if dist is None:
        return x
    import torch

    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def octaves_for(n):
    return max(1, int(np.log2(n)) - 2)


class Workload:
    """Synthetic raster on the device + resident / pinned output buffers."""

    def __init__(self, size, seed, device, nrow=None):
        from pyflwdir_b200 import _lib

        self.L = _lib
        self.l = _lib.lib()
        self.n = size if nrow is None else nrow   # rows
        self.ncol = size
        self.cells = self.n * size
        h = C.c_void_p()
        _lib.check(self.l.pfd_create(device, C.byref(h)))
        self.h = h
        self.d8_dev = self.dev_alloc(self.cells)
        # same generator call as the row blocks of the multi-GPU arm (bit-identical to synth_elevation + synth_d8)
        self.ck(self.l.pfd_synth_d8_block(h, 0, self.n, size, self.n, size, octaves_for(size), seed, C.c_float(-np.inf), self.d8_dev))
        self.ck(self.l.pfd_set_option(h, b"release_scratch", 1))  # the generator's float plane
        self.idx_dtype = np.int32 if self.cells < 2**31 - 1 else (np.uint32 if self.cells < 2**32 - 2 else np.int64)
        idx_b = np.dtype(self.idx_dtype).itemsize
        self.out_dev = [self.dev_alloc(self.cells * idx_b)] + [self.dev_alloc(self.cells * 4) for _ in range(3)]  # idxs_ds, rank, uparea, basins

    def ck(self, rc):
        self.L.check(rc, self.h)

    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        self.ck(self.l.pfd_dev_alloc(self.h, nbytes, C.byref(p)))
        return p

    def step_resident(self):
        o = self.out_dev
        self.ck(self.l.pfd_d8_flow_all(self.h, self.d8_dev, self.n, self.ncol, o[0], self.L.DTYPES[np.dtype(self.idx_dtype)],
                                       o[1], o[2], o[3], None, None, None))

    def checksums(self, index_offset=0, row0=0, nrow=None):
        """pfd_checksum of (idxs_ds, rank, uparea, basins) over rows [row0, row0 + nrow) of the resident outputs."""
        nrow = self.n if nrow is None else nrow
        out = []
        for p, b in zip(self.out_dev, (np.dtype(self.idx_dtype).itemsize, 4, 4, 4)):
            v = C.c_uint64()
            q = C.c_void_p(p.value + row0 * self.ncol * b)
            self.ck(self.l.pfd_checksum(self.h, q, b, nrow * self.ncol, index_offset, C.byref(v)))
            out.append(int(v.value))
        return out

    def free(self):
        for p in [self.d8_dev] + self.out_dev:
            self.l.pfd_dev_free(self.h, p)
        self.l.pfd_destroy(self.h)

    def make_host(self):
        self.d8_host = self.L.PinnedArray((self.n, self.ncol), np.uint8)
        self.ck(self.l.pfd_memcpy(self.h, self.L.ptr(self.d8_host.array), self.d8_dev, self.cells))
        self.out_host = [self.L.PinnedArray(self.cells, dt) for dt in (self.idx_dtype, np.int32, np.int32, np.uint32)]

    def step_host(self):
        o = [self.L.ptr(a.array) for a in self.out_host]
        self.ck(self.l.pfd_d8_flow_all(self.h, self.L.ptr(self.d8_host.array), self.n, self.ncol, o[0],
                                       self.L.DTYPES[np.dtype(self.idx_dtype)], o[1], o[2], o[3], None, None, None))

    def stage_ms(self):
        g = self.l.pfd_last_stage_ms
        tiles = self.l.pfd_get_info(self.h, b"tiles") == 1
        if tiles:
            st = {"parse": g(self.h, 0), "pits": g(self.h, 1), "tile_a": g(self.h, 6), "tile_b": g(self.h, 7),
                  "tile_c": g(self.h, 8), "total": g(self.h, 4)}
            if st["parse"] == 0.0:  # fused-parse path: no separate parse pass ran
                del st["parse"]
            return st
        return {"parse": g(self.h, 0), "pits": g(self.h, 1), "order": g(self.h, 2), "sweep": g(self.h, 3),
                "total": g(self.h, 4), "bfs": g(self.h, 5)}

    def timer(self, fn, steps):
        self.ck(self.l.pfd_timer_start(self.h))
        for _ in range(steps):
            fn()
        ms = C.c_double()
        self.ck(self.l.pfd_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launches(self):
        return int(self.l.pfd_launch_count(self.h))


class TiledWorkload(Workload):
    """Rank `rank` of `world`: rows [rank*size, (rank+1)*size) of a (world*size) x size raster, one GPU per rank."""

    def __init__(self, size, seed, device, rank, world, dist, strong=False, rows_per_rank=None):
        from pyflwdir_b200 import _lib, tiled

        self.L = _lib
        self.l = _lib.lib()
        self.ncol = size
        if strong:
            blocks = tiled.split_rows(size, world)
            assert len(blocks) == world, "raster too small for this many ranks"
            self.row0, r1 = blocks[rank]
            self.n = r1 - self.row0          # rows owned by this rank
            self.nrow_global = size
        else:
            self.n = size if rows_per_rank is None else rows_per_rank
            self.row0 = rank * self.n
            self.nrow_global = world * self.n
        self.cells = self.n * size
        self.rank, self.world = rank, world
        h = C.c_void_p()
        _lib.check(self.l.pfd_create(device, C.byref(h)))
        self.h = h
        uid = [tiled.RowBlockSolver.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        buf = (C.c_uint8 * 128).from_buffer_copy(uid[0])
        self.ck(self.l.pfd_comm_init(h, rank, world, buf))
        self.ht = 1 if rank > 0 else 0
        self.hb = 1 if rank < world - 1 else 0
        ext = (self.n + self.ht + self.hb) * size
        self.d8_dev = self.dev_alloc(ext)
        self.ck(self.l.pfd_synth_d8_block(h, self.row0 - self.ht, self.n + self.ht + self.hb, size, self.nrow_global, size,
                                          octaves_for(size), seed, C.c_float(-np.inf), self.d8_dev))
        self.ck(self.l.pfd_set_option(h, b"release_scratch", 1))
        self.ext_bytes = ext
        ncells_global = self.nrow_global * size
        self.idx_dtype = np.int32 if ncells_global < 2**31 - 1 else (np.uint32 if ncells_global < 2**32 - 2 else np.int64)
        self.idx_bytes = self.cells * np.dtype(self.idx_dtype).itemsize
        self.out_dev = [self.dev_alloc(self.idx_bytes)] + [self.dev_alloc(self.cells * 4) for _ in range(3)]
        self.n_pits_global = C.c_int64()

    def _call(self, d8, idxs, rank_o, upa_o, bas_dev):
        self.ck(self.l.pfd_d8_flow_all_tiled(self.h, d8, self.n, self.ncol, self.ht, self.hb, self.row0, idxs,
                                             self.L.DTYPES[np.dtype(self.idx_dtype)], rank_o, upa_o, bas_dev, None,
                                             C.byref(self.n_pits_global)))

    def step_resident(self):
        o = self.out_dev
        self._call(self.d8_dev, o[0], o[1], o[2], o[3])

    def make_host(self):
        self.d8_host = self.L.PinnedArray(self.ext_bytes, np.uint8)
        self.ck(self.l.pfd_memcpy(self.h, self.L.ptr(self.d8_host.array), self.d8_dev, self.ext_bytes))
        self.out_host = [self.L.PinnedArray(self.cells, dt) for dt in (self.idx_dtype, np.int32, np.int32, np.uint32)]

    def step_host(self):
        o = [self.L.ptr(a.array) for a in self.out_host]
        self._call(self.L.ptr(self.d8_host.array), o[0], o[1], o[2], self.out_dev[3])
        self.ck(self.l.pfd_memcpy(self.h, o[3], self.out_dev[3], self.cells * 4))  # basins travel through a device buffer

    def free(self):
        for p in [self.d8_dev] + self.out_dev:
            self.l.pfd_dev_free(self.h, p)
        self.l.pfd_comm_destroy(self.h)
        self.l.pfd_destroy(self.h)


def extras(w, size, seed, skip_level=False):
    """Secondary configs of BASELINE.json on the same raster (device-resident, CUDA events, best of 3): the exact
    idxs_seq ordering (level-synchronous BFS), Strahler order, float64 accuflux and HAND level sweeps."""
    l, L, h, n = w.l, w.L, w.h, w.cells
    z_dev = w.dev_alloc(n * 4)
    w.ck(l.pfd_synth_elevation(h, size, size, size, octaves_for(size), seed, z_dev))
    upa = np.empty(n, np.int32)
    w.ck(l.pfd_d8_parse(h, w.d8_dev, size, size, 1, None, 0, None, None, None))
    w.ck(l.pfd_upstream_area_cells(h, L.ptr(upa)))
    drain = (upa > 1000).astype(np.uint8)
    drain_dev = w.dev_alloc(n)
    w.ck(l.pfd_memcpy(h, drain_dev, L.ptr(drain), n))
    f64_dev, out8_dev, out1_dev = w.dev_alloc(n * 8), w.dev_alloc(n * 8), w.dev_alloc(n)

    def best(fn, reps=3, prep=None):
        t = []
        for _ in range(reps):
            if prep:
                prep()
            t.append(w.timer(fn, 1))
        return min(t)

    res = {}
    f32, i32 = L.DTYPES[np.dtype(np.float32)], L.DTYPES[np.dtype(np.int32)]
    calls = {
        "strahler": lambda: w.ck(l.pfd_strahler(h, None, out1_dev)),
        "accuflux_i32": lambda: w.ck(l.pfd_accuflux(h, w.out_dev[2], i32, -9999.0, -9999, 1, 0, w.out_dev[1])),
        "accuflux_f32": lambda: w.ck(l.pfd_accuflux(h, z_dev, f32, -9999.0, 0, 0, 0, w.out_dev[1])),
        "hand": lambda: w.ck(l.pfd_hand(h, drain_dev, z_dev, f32, out8_dev)),
    }
    # HAND as verified path sums (pfd_hand.cuh, the default engine of pfd_hand): no ordering, no passes
    calls["hand"]()
    res["hand_pathsum"] = {"hand_ms": best(calls["hand"]), "engine": int(l.pfd_get_info(h, b"hand_engine")),
                           "note": "engine 1 = path sums accepted by the per-cell proof; 2 / 3 = a hop-by-hop engine answered"}
    res["hand_pathsum"]["hand_mcells_s"] = n / (res["hand_pathsum"]["hand_ms"] / 1e3) / 1e6
    w.ck(l.pfd_set_option(h, b"hand_pathsum", 0))  # the two sections below time the hop-by-hop engines
    # tile-dataflow sweeps (default): no ordering needed
    w.ck(l.pfd_set_option(h, b"tile_sweeps", 2))
    tile = {}
    for k, fn in calls.items():
        fn()
        tile[k + "_ms"] = best(fn)
        tile[k + "_passes"] = int(l.pfd_get_info(h, b"sweep_passes"))
        tile[k + "_mcells_s"] = n / (tile[k + "_ms"] / 1e3) / 1e6
    res["tile_dataflow"] = tile
    if not skip_level:
        # level replays over the BFS order (round-1 path, option tile_sweeps = 0)
        w.ck(l.pfd_set_option(h, b"tile_sweeps", 0))
        reparse = lambda: w.ck(l.pfd_d8_parse(h, w.d8_dev, size, size, 1, None, 0, None, None, None))
        lev = {"order_idxs_seq_ms": best(lambda: w.ck(l.pfd_order(h, None, None)), prep=reparse)}
        lev["nlevels"] = int(l.pfd_get_info(h, b"nlevels"))
        for k, fn in calls.items():
            fn()
            lev[k + "_ms"] = best(fn)
            lev[k + "_mcells_s"] = n / (lev[k + "_ms"] / 1e3) / 1e6
        res["level_replay"] = lev
        w.ck(l.pfd_set_option(h, b"tile_sweeps", 1))
    w.ck(l.pfd_set_option(h, b"hand_pathsum", 1))
    res["note"] = ("device-resident, same raster, best of 3, CUDA events. BASELINE config 3 (accuflux + Strahler) = the headline "
                   "step + strahler; config 5 (HAND) = headline step + hand (hand_pathsum when its engine is 1, else the tile-dataflow "
                   "sweep); level_replay additionally needs order_idxs_seq")
    for p in (z_dev, drain_dev, f64_dev, out8_dev, out1_dev):
        w.ck(l.pfd_dev_free(h, p))
    return res


def extras_widened(w, size, seed, cpu_size):
    """The widened rows of SURVEY.md §8f on the same raster: every entry point timed device-resident (CUDA events, best
    of 3, ordering already cached) next to the oracle port of the same reference function on a `cpu_size`^2 raster of
    the same generator (1 host thread). Returns {name: {"gpu_ms", "gpu_mcells_s", "cpu_mcells_s"}}."""
    import oracle

    l, L, h, n = w.l, w.L, w.h, w.cells
    i32, f32 = L.DTYPES[np.dtype(np.int32)], L.DTYPES[np.dtype(np.float32)]
    z_dev = w.dev_alloc(n * 4)
    w.ck(l.pfd_synth_elevation(h, size, size, size, octaves_for(size), seed, z_dev))
    w.ck(l.pfd_d8_parse(h, w.d8_dev, size, size, 1, None, 0, None, None, None))
    w.ck(l.pfd_order(h, None, None))
    upa_dev, out4_dev, out1_dev, um_dev = w.out_dev[2], w.out_dev[1], w.dev_alloc(n), w.dev_alloc(n * 4)
    w.ck(l.pfd_upstream_area_cells(h, upa_dev))
    upa = np.empty(n, np.int32)
    w.ck(l.pfd_memcpy(h, L.ptr(upa), upa_dev, n * 4))
    mask_dev = w.dev_alloc(n)
    mask_h = (upa > 100).astype(np.uint8)  # (named: L.ptr() does not keep a temporary alive)
    w.ck(l.pfd_memcpy(h, mask_dev, L.ptr(mask_h), n))
    # sparse data for fillnodata: keep ~30 % of the elevations, the rest is nodata
    rng = np.random.default_rng(seed + 5)
    zh = np.empty(n, np.float32)
    w.ck(l.pfd_memcpy(h, L.ptr(zh), z_dev, n * 4))
    gaps = rng.random(n) < 0.7
    sparse_dev = w.dev_alloc(n * 4)
    sparse_h = np.where(gaps, np.float32(-9999.0), zh)
    w.ck(l.pfd_memcpy(h, sparse_dev, L.ptr(sparse_h), n * 4))
    drainh = np.where(upa >= 1000, np.power(upa.astype(np.float64), 0.3), -9999.0).astype(np.float32)
    drainh_dev = w.dev_alloc(n * 4)
    w.ck(l.pfd_memcpy(h, drainh_dev, L.ptr(drainh), n * 4))
    del zh, gaps, drainh, sparse_h, mask_h

    def best(fn, reps=3):
        fn()
        return min(w.timer(fn, 1) for _ in range(reps))

    gpu = {}
    gpu["main_upstream"] = best(lambda: w.ck(l.pfd_main_upstream(h, upa_dev, i32, C.c_double(0.0), um_dev, i32)))
    gpu["upstream_count_masked"] = best(lambda: w.ck(l.pfd_upstream_count(h, mask_dev, out1_dev)))
    gpu["stream_order_classic"] = best(lambda: w.ck(l.pfd_stream_order_classic(h, um_dev, i32, None, out1_dev)))
    gpu["accuflux_ds_f32"] = best(lambda: w.ck(l.pfd_accuflux(h, z_dev, f32, C.c_double(-9999.0), 0, 0, 1, out4_dev)))
    gpu["fillnodata_up_f32"] = best(lambda: w.ck(l.pfd_fillnodata(h, sparse_dev, f32, C.c_double(-9999.0), 0, 0, 0, 0, out4_dev)))
    gpu["fillnodata_down_max_f32"] = best(lambda: w.ck(l.pfd_fillnodata(h, sparse_dev, f32, C.c_double(-9999.0), 0, 0, 1, 0, out4_dev)))
    gpu["stream_distance_cells"] = best(lambda: w.ck(l.pfd_stream_distance(h, mask_dev, 0, None, out4_dev)))
    gpu["floodplains"] = best(lambda: w.ck(l.pfd_floodplains(h, drainh_dev, z_dev, f32, out1_dev)))
    ldd_dev = w.dev_alloc(n)
    w.ck(l.pfd_fetch(h, L.ARR_LDD, ldd_dev, 0))
    gpu["ldd_parse"] = best(lambda: w.ck(l.pfd_ldd_parse(h, ldd_dev, size, size, w.out_dev[0], i32, None, None)))
    gpu["to_array_d8"] = best(lambda: w.ck(l.pfd_fetch(h, L.ARR_D8, out1_dev, 0)))
    # later rows: arithmetics, subbasins, local / region post-processing, NEXTXY codec
    f64 = L.DTYPES[np.dtype(np.float64)]
    u32 = L.DTYPES[np.dtype(np.uint32)]
    so_dev = w.dev_alloc(n)
    w.ck(l.pfd_strahler(h, None, so_dev))
    k64 = C.c_int64()
    gpu["upstream_sum_f32"] = best(lambda: w.ck(l.pfd_upstream_sum(h, z_dev, f32, C.c_double(-9999.0), 0, 0, out4_dev)))
    gpu["subbasins_streamorder"] = best(lambda: w.ck(l.pfd_subbasins_streamorder(h, so_dev, None, -2, out4_dev, C.byref(k64))))
    gpu["subbasins_area"] = best(lambda: w.ck(l.pfd_subbasins_area(h, um_dev, i32, upa_dev, i32, C.c_double(5000.0), out4_dev, C.byref(k64))))
    gpu["moving_average_n3_f32"] = best(lambda: w.ck(l.pfd_moving_average(h, z_dev, f32, None, 0, 3, um_dev, i32, None, C.c_double(-9999.0), out4_dev)))
    gpu["moving_median_n3_f32"] = best(lambda: w.ck(l.pfd_moving_median(h, z_dev, f32, 3, um_dev, i32, None, C.c_double(-9999.0), out4_dev)))
    gpu["downstream_f32"] = best(lambda: w.ck(l.pfd_downstream(h, z_dev, f32, out4_dev)))
    region_h = np.zeros((size, size), np.uint8)
    region_h[size // 4: 3 * size // 4, size // 5: 4 * size // 5] = 1
    region_dev = w.dev_alloc(n)
    w.ck(l.pfd_memcpy(h, region_dev, L.ptr(region_h), n))
    gpu["outflow_idxs"] = best(lambda: w.ck(l.pfd_outflow_idxs(h, region_dev, C.byref(k64))))
    gpu["inflow_idxs"] = best(lambda: w.ck(l.pfd_inflow_idxs(h, region_dev, C.byref(k64))))
    gpu["interbasin_mask_stream"] = best(lambda: w.ck(l.pfd_interbasin_mask(h, region_dev, mask_dev, out1_dev)))
    bas_dev = w.out_dev[3]
    w.ck(l.pfd_basins(h, None, 0, i32, None, u32, bas_dev))
    gpu["region_outlets_basins"] = best(lambda: w.ck(l.pfd_region_outlets(h, bas_dev, u32, C.byref(k64))))
    gpu["region_slices_basins"] = best(lambda: w.ck(l.pfd_region_slices(h, bas_dev, u32, C.byref(k64))))
    starts_h = np.ascontiguousarray(np.random.default_rng(seed + 9).integers(0, n, size=4096), dtype=np.int64)
    cnt_h, end_h, dist_h = np.empty(4096, np.int64), np.empty(4096, np.int64), np.empty(4096, np.float64)
    gpu["snap_4096_starts"] = best(lambda: w.ck(l.pfd_trace(h, L.ptr(starts_h), 4096, 0, None, 0, mask_dev, 0, C.c_double(0.0), None,
                                                           L.ptr(cnt_h), L.ptr(end_h), L.ptr(dist_h), None, 0, 0)))
    i64 = L.DTYPES[np.dtype(np.int64)]
    pf_dev = w.dev_alloc(n * 8)
    n2 = C.c_int64()
    gpu["subbasins_pfafstetter_d2"] = best(lambda: w.ck(l.pfd_subbasins_pfafstetter(h, um_dev, i32, upa_dev, i32, mask_dev, 2, pf_dev, C.byref(k64))))
    gpu["streams_mask_len25"] = best(lambda: w.ck(l.pfd_streams(h, mask_dev, 25, C.byref(k64), C.byref(n2))))
    est_h = np.zeros(n, np.int8)
    est_dev = w.dev_alloc(n)
    w.ck(l.pfd_memcpy(h, est_dev, L.ptr(est_h), n))
    gpu["classify_estuary_f32"] = best(lambda: w.ck(l.pfd_classify_estuary(h, est_dev, z_dev, f32, z_dev, f32, C.c_double(1e-2), out1_dev)))
    w.ck(l.pfd_dev_free(h, pf_dev))
    w.ck(l.pfd_dev_free(h, est_dev))
    nxy_dev = w.dev_alloc(n * 8)
    w.ck(l.pfd_fetch(h, L.ARR_NEXTXY, nxy_dev, 0))
    gpu["nextxy_parse"] = best(lambda: w.ck(l.pfd_nextxy_parse(h, nxy_dev, C.c_void_p(nxy_dev.value + n * 4), size, size, 1, w.out_dev[0], i32, None, None, None)))
    for pdev in (z_dev, out1_dev, um_dev, mask_dev, sparse_dev, drainh_dev, ldd_dev, so_dev, region_dev, nxy_dev):
        w.ck(l.pfd_dev_free(h, pdev))
    # the step before the path: dem.fill_depressions of the synthetic terrain in metres (float32,
    # z * 700 + 2000), device-resident in and out, on at most 8192^2 cells
    fsz = min(size, 8192)
    fz = (oracle.synth_elevation(fsz, fsz, seed=seed, octaves=octaves_for(fsz), nref=fsz) * np.float32(700.0) + np.float32(2000.0)).astype(np.float32)
    fin_dev, fout_dev, fd8_dev = w.dev_alloc(fz.size * 4), w.dev_alloc(fz.size * 4), w.dev_alloc(fz.size)
    w.ck(l.pfd_memcpy(h, fin_dev, L.ptr(fz), fz.size * 4))
    fstats = np.zeros(12, np.int64)
    fill_call = lambda: w.ck(l.pfd_fill_depressions(h, fin_dev, f32, fsz, fsz, 0, None, 0, C.c_double(-9999.0), C.c_double(-1.0), 0,
                                                    C.c_double(0.0), 8, 0, fout_dev, fd8_dev, L.ptr(fstats)))
    fill_ms = best(fill_call, reps=2)
    fstats0 = fstats.copy()
    # the same terrain on a regional slope (3 m per cell along the diagonal): depressions stay local, as in a real DEM -- without
    # it the fBm surface has basin-scale sinks whose lakes (millions of cells) are replayed serially
    fzt = (fz + (np.arange(fsz, dtype=np.float32)[:, None] + np.arange(fsz, dtype=np.float32)[None, :]) * np.float32(1.5)).astype(np.float32)
    w.ck(l.pfd_memcpy(h, fin_dev, L.ptr(fzt), fzt.size * 4))
    fill_t_ms = best(fill_call, reps=2)
    fstats1 = fstats.copy()
    for pdev in (fin_dev, fout_dev, fd8_dev):
        w.ck(l.pfd_dev_free(h, pdev))

    # CPU port on a bounded sample of the same generator
    oracle.build()
    d8, _ = host_raster(cpu_size, seed)
    m = d8.size
    zc = oracle.synth_elevation(cpu_size, cpu_size, seed=seed, octaves=octaves_for(cpu_size), nref=cpu_size).ravel()
    ids, pits, _ = oracle.core_d8.from_array(d8, dtype=np.int32)
    seq = oracle.core.idxs_seq(ids, pits)
    upc = oracle.streams.accuflux(ids, seq, np.ones(m, np.int32), -9999)
    cmask = upc > 100
    sparse = np.where(np.random.default_rng(seed + 5).random(m) < 0.7, np.float32(-9999.0), zc)

    def cpu_time(fn):
        t0 = time.perf_counter()
        r = fn()
        return time.perf_counter() - t0, r

    cpu = {}
    cpu["main_upstream"], um = cpu_time(lambda: oracle.core.main_upstream(ids, upc, 0.0))
    cpu["upstream_count_masked"], _ = cpu_time(lambda: oracle.core.upstream_count(ids, mask=cmask))
    cpu["stream_order_classic"], _ = cpu_time(lambda: oracle.streams.stream_order(ids, seq, um))
    cpu["accuflux_ds_f32"], _ = cpu_time(lambda: oracle.streams.accuflux_ds(ids, seq, zc, -9999.0))
    cpu["fillnodata_up_f32"], _ = cpu_time(lambda: oracle.core.fillnodata_upstream_any(ids, seq, sparse, -9999.0))
    cpu["fillnodata_down_max_f32"], _ = cpu_time(lambda: oracle.core.fillnodata_downstream(ids, seq, sparse, -9999.0, "max"))
    cpu["stream_distance_cells"], _ = cpu_time(lambda: oracle.streams.stream_distance(ids, seq, cpu_size, mask=cmask, real_length=False))
    cpu["floodplains"], _ = cpu_time(lambda: oracle.dem.floodplains(ids, seq, zc, upc, 1000.0, 0.3))
    ldd = oracle.core_ldd.to_array(ids, d8.shape) if hasattr(oracle, "core_ldd") else None
    if ldd is not None:
        cpu["ldd_parse"], _ = cpu_time(lambda: oracle.core_ldd.from_array(ldd, dtype=np.int32))
    cpu["to_array_d8"], _ = cpu_time(lambda: oracle.core_d8.to_array(ids, d8.shape))
    so = oracle.streams.strahler_order(ids, seq)
    cpu["upstream_sum_f32"], _ = cpu_time(lambda: oracle.arithmetics.upstream_sum(ids, zc, -9999.0))
    cpu["subbasins_streamorder"], _ = cpu_time(lambda: oracle.basins.subbasins_streamorder(ids, seq, so, None, -2))
    cpu["subbasins_area"], _ = cpu_time(lambda: oracle.basins.subbasins_area(ids, seq, um, upc, 5000.0))
    cpu["moving_average_n3_f32"], _ = cpu_time(lambda: oracle.arithmetics.moving_average(zc, None, 3, ids, um, None, -9999.0))
    cpu["moving_median_n3_f32"], _ = cpu_time(lambda: oracle.arithmetics.moving_median(zc, 3, ids, um, None, -9999.0))
    creg = np.zeros(d8.shape, np.uint8)
    creg[cpu_size // 4: 3 * cpu_size // 4, cpu_size // 5: 4 * cpu_size // 5] = 1
    creg = creg.ravel()
    cpu["outflow_idxs"], _ = cpu_time(lambda: oracle.core.outflow_idxs(ids, seq, creg))
    cpu["inflow_idxs"], _ = cpu_time(lambda: oracle.core.inflow_idxs(ids, seq, creg))
    cpu["interbasin_mask_stream"], _ = cpu_time(lambda: oracle.basins.interbasin_mask(ids, seq, creg, cmask))
    cbas = oracle.basins.basins(ids, pits, seq)
    cpu["region_outlets_basins"], _ = cpu_time(lambda: oracle.regions.region_outlets(cbas, ids, seq))
    cpu["subbasins_pfafstetter_d2"], _ = cpu_time(lambda: oracle.basins.subbasins_pfafstetter(pits, ids, seq, um, upc, mask=cmask, depth=2))
    cpu["streams_mask_len25"], _ = cpu_time(lambda: oracle.streams.streams(ids, seq, cmask, 25))
    cpu["classify_estuary_f32"], _ = cpu_time(lambda: oracle.rivers.classify_estuary(ids, seq, pits, zc, zc, zc, -1e9, 1e-2))
    cnxy = oracle.core_nextxy.to_array(ids, d8.shape)
    cpu["nextxy_parse"], _ = cpu_time(lambda: oracle.core_nextxy.from_array(cnxy, dtype=np.int32))
    res = {}
    for k, ms in gpu.items():
        res[k] = {"gpu_ms": ms, "gpu_mcells_s": n / (ms / 1e3) / 1e6,
                  "cpu_mcells_s": (m / cpu[k] / 1e6) if k in cpu else None}
    # fill_depressions: the CPU arm is the numba reference itself when oracle/_ref is there (else the C port), 1024^2 sample
    csz = min(cpu_size, 1024)
    fzc = np.ascontiguousarray(fz[:csz, :csz])
    fill_kind, fill_fn = "port", oracle.dem.fill_depressions
    try:
        from oracle import reference as _ref
        if _ref.available():
            _ref.load()
            from pyflwdir import dem as _rdem
            _rdem.fill_depressions(fzc[:64, :64].copy())  # JIT
            fill_kind, fill_fn = "reference", _rdem.fill_depressions
    except Exception:
        pass
    t_fill, _ = cpu_time(lambda: fill_fn(fzc.copy()))
    def fill_entry(ms, st, t_cpu, what):
        return {"gpu_ms": ms, "gpu_mcells_s": fz.size / (ms / 1e3) / 1e6, "cpu_mcells_s": fzc.size / t_cpu / 1e6, "cpu_kind": fill_kind,
                "terrain": what, "gpu_raster": f"{fsz}^2 float32", "cpu_raster": f"{csz}^2", "level_passes": int(st[0]),
                "label_passes": int(st[1]), "tied_cells": int(st[2]), "tie_components": int(st[3]), "largest_component": int(st[7]),
                "levels_ms": st[10] / 1e3, "ties_ms": st[11] / 1e3}

    res["fill_depressions"] = fill_entry(fill_ms, fstats0, t_fill, "rough fBm in metres, no regional slope: basin-scale sinks")
    t_fill_t, _ = cpu_time(lambda: fill_fn(np.ascontiguousarray(fzt[:csz, :csz])))
    res["fill_depressions_sloped"] = fill_entry(fill_t_ms, fstats1, t_fill_t, "the same + 3 m per cell of regional slope: local depressions")
    res["note"] = (f"GPU: {size}^2 raster, device-resident buffers, ordering cached, best of 3; CPU: oracle port of the same "
                   f"reference function on a {cpu_size}^2 raster of the same generator, 1 thread (the reference's "
                   "stream_distance / floodplains are interpreted Python and far slower than this C port)")
    return res


CPU_STAGES = ("from_array", "rank", "idxs_seq", "accuflux", "basins")


def cpu_modules():
    """(kind, core_d8, core, streams, basins): the reference's numba modules when installed, else the oracle port."""
    import oracle
    from oracle import reference

    if reference.available() and os.environ.get("PFD_BENCH_CPU", "reference") != "port":
        try:
            pf = reference.load()
            import importlib

            mods = [importlib.import_module(f"pyflwdir.{m}") for m in ("core_d8", "core", "streams", "basins")]
            return ("reference",) + tuple(mods)
        except Exception as exc:  # numba / reference import problem on this box: say so and time the port
            print(f"bench.py: reference import failed ({exc!r}); timing the oracle port", file=sys.stderr)
    oracle.build()
    return "port", oracle.core_d8, oracle.core, oracle.streams, oracle.basins


def cpu_path(mods, d8):
    """One pass of the reference's CPU path on `d8`: seconds per stage. The reference needs idxs_seq for its accuflux /
    basins sweeps (pyflwdir.py:292-297), so it is part of its path; the GPU path needs no ordering for these outputs."""
    kind, core_d8, core, streams, basins = mods
    t = {}
    t0 = time.perf_counter()
    ids, pits, _ = core_d8.from_array(d8, dtype=np.int32)
    t["from_array"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    core.rank(ids, mv=np.int32(-1)) if kind == "reference" else core.rank(ids)
    t["rank"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    seq = core.idxs_seq(ids, pits, mv=np.int32(-1)) if kind == "reference" else core.idxs_seq(ids, pits)
    t["idxs_seq"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    streams.accuflux(ids, seq, np.ones(d8.size, np.int32), -9999)
    t["accuflux"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    basins.basins(ids, pits, seq)
    t["basins"] = time.perf_counter() - t0
    t["total"] = sum(t.values())
    return t


def cpu_warm(mods):
    """JIT-compile the numba kernels on a 256^2 raster outside any timed region."""
    d8, _ = host_raster(256, 1)
    cpu_path(mods, d8)


def cpu_describe(mods, cs, st):
    kind = mods[0]
    what = ("UNMODIFIED reference (Deltares/pyflwdir numba kernels, JIT warm, single-threaded by construction)"
            if kind == "reference" else "oracle C port of the numba kernels")
    return (f"{cs}x{cs} sample of the same generator/seed, {what}, 1 thread of {os.cpu_count()} host cores; s/stage: " +
            ", ".join(f"{k}={v:.3f}" for k, v in st.items()))


def host_raster(size, seed):
    """Synthetic raster built on the HOST (oracle generator, bit-identical to the CUDA one): the reference arm touches
    nothing of the GPU path."""
    import oracle

    oracle.build()
    z = oracle.synth_elevation(size, size, seed=seed, octaves=octaves_for(size), nref=size)
    return oracle.synth_d8(z), "synthetic (host generator)"


def workload_config(args, world=1, w=None):
    """The `config` object: identical for both arms of one invocation (ours / --impl reference)."""
    if world == 1:
        what = f"synthetic {args.size}x{args.size} D8 raster"
        par = "single GPU"
    else:
        nrow_global = args.size if args.scaling == "strong" else world * args.size
        what = f"ONE synthetic {nrow_global}x{args.size} D8 raster row-tiled over {world} GPUs"
        par = f"row-tiled x{world}: pit-count all-gather + 1 NCCL all-reduce of boundary tables per step"
    return {
        "workload": what + f" (rough Perlin fBm, {octaves_for(args.size)} octaves, steepest descent, no depression filling): "
                           "parse->idxs_ds + rank + upstream_area(cell) + basins (BASELINE.json configs[2] size, configs[1] "
                           "outputs)",
        "seed": args.seed,
        "cpu_sample": f"{args.cpu_size}x{args.cpu_size} raster of the same generator per CPU step (throughput is flat in "
                      "size, BASELINE.md section 2); the CPU path also computes idxs_seq, which its sweeps need",
        "l2": "per-step working set ~26 B/cell x N cells (>= 1.7 GB at 8192^2) exceeds the 126 MB L2; no flush needed",
        "parallelism": par,
    }


def run_reference(args):
    """--impl reference: the reference's CPU path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    mods = cpu_modules()
    cpu_warm(mods)
    d8, data = host_raster(args.cpu_size, args.seed)
    cells = d8.size
    for _ in range(args.warmup):
        cpu_path(mods, d8)
    t0 = time.perf_counter()
    stages = None
    for _ in range(args.steps):
        st = cpu_path(mods, d8)
        stages = st if stages is None or st["total"] < stages["total"] else stages
    dt = time.perf_counter() - t0
    value = cells * args.steps / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": mods[0], "sample": cpu_describe(mods, args.cpu_size, stages),
                         "accuflux_plus_basins_mcells_s": cells / (stages["accuflux"] + stages["basins"]) / 1e6},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=32768,
                    help="raster is size x size (default: the north-star size, BASELINE.json configs[2]; configs[1] is 8192); "
                         "N > 1: rows per GPU (weak) or the whole raster (strong)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-size", type=int, default=4096,
                    help="side of the raster the CPU arm / cpu_baseline samples per step (same generator; both arms use it)")
    ap.add_argument("--solver", default="tiles", choices=["tiles", "bfs"], help="rank/basins/uparea solver")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = every GPU owns `size` rows of a (N*size) x size raster (default); "
                         "strong = ONE size x size raster split into N row blocks (BASELINE config 4 with --size 65536)")
    ap.add_argument("--no-fuse", action="store_true", help="run the separate parse pass even on device-resident buffers")
    ap.add_argument("--no-e2e", action="store_true", help="skip the pinned-host end-to-end arm (very large rasters)")
    ap.add_argument("--verify-full", action="store_true",
                    help="N>1: compare the FULL-size outputs with a single-GPU solve of the whole raster on rank 0 (needs the "
                         "raster to fit one GPU; the multi-rank buffers are freed first)")
    ap.add_argument("--extras", nargs="?", const="sweeps", default=None, choices=["none", "tile", "sweeps", "all"],
                    help="also time the secondary configs: Strahler / accuflux / HAND sweeps (tile = tile-dataflow only, sweeps = "
                         "also the level replays + ordering, all = plus every widened entry point)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    dist, rank, world, local_rank = dist_setup(args.gpus)
    from pyflwdir_b200 import _lib

    if _lib.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device -- pyflwdir_b200 has no CPU fallback")
    device = local_rank % _lib.device_count()
    numa_node = _lib.bind_to_device_numa_node(device) if world > 1 else None  # before any pinned allocation
    if world > 1:
        w = TiledWorkload(args.size, args.seed, device, rank, world, dist, strong=(args.scaling == "strong"))
    else:
        w = Workload(args.size, args.seed, device)
        w.ck(w.l.pfd_set_option(w.h, b"tiles", 1 if args.solver == "tiles" else 0))
        w.ck(w.l.pfd_set_option(w.h, b"fuse_parse", 0 if args.no_fuse else 1))
    cells = w.cells

    # ---- device-resident arm. The clock sampler (rank 0) runs from before the warm-up until after a >= 1.2 s
    # window of back-to-back steps that directly follows the timed region (the timed region itself is only
    # K x ~2 ms, shorter than one nvidia-smi sampling period).
    sampler = ClockSampler(device)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        w.step_resident()
    barrier(dist)
    l0 = w.launches()
    stage_acc = {}
    w.ck(w.l.pfd_timer_start(w.h))
    for _ in range(args.steps):
        w.step_resident()
        for k, v in w.stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    ms = C.c_double()
    w.ck(w.l.pfd_timer_stop(w.h, C.byref(ms)))
    launches = w.launches() - l0
    t_probe = time.perf_counter()
    while time.perf_counter() - t_probe < 1.2:
        w.step_resident()
    barrier(dist)
    clocks = sampler.stop() if rank == 0 else None
    ms_total = reduce_max(dist, ms.value)
    cells_total = reduce_sum(dist, float(cells))
    value = cells_total * args.steps / (ms_total / 1e3) / 1e6
    stage_avg = {k: v / args.steps for k, v in stage_acc.items()}

    # ---- end-to-end arm (pinned host buffers through the same C-ABI call)
    e2e = None
    if not args.no_e2e:
        w.make_host()
        # the host link under the same concurrency: every rank copies its uparea output D2H at the same time (plain
        # cudaMemcpy into pinned memory). e2e is bound by this number, not by the kernels (0.3 s of copies vs 0.02 s of GPU work)
        probe_bytes = cells * 4
        w.ck(w.l.pfd_memcpy(w.h, w.L.ptr(w.out_host[2].array), w.out_dev[2], probe_bytes))
        barrier(dist)
        t0 = time.perf_counter()
        for _ in range(3):
            w.ck(w.l.pfd_memcpy(w.h, w.L.ptr(w.out_host[2].array), w.out_dev[2], probe_bytes))
        probe_s = reduce_max(dist, (time.perf_counter() - t0) / 3)
        d2h_gbs_per_gpu = probe_bytes / probe_s / 1e9
        for _ in range(2):
            w.step_host()
        barrier(dist)
        e2e_ms = reduce_max(dist, w.timer(w.step_host, args.steps))
        barrier(dist)
        idx_b = np.dtype(getattr(w, "idx_dtype", np.int32)).itemsize
        e2e = {"value": cells_total * args.steps / (e2e_ms / 1e3) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(getattr(w, "ext_bytes", cells)), "d2h_bytes_per_step": int((12 + idx_b) * cells),
               "ms_per_step": e2e_ms / args.steps, "note": "bytes are per GPU", "numa_node_rank0": numa_node,
               "host_link": {"d2h_gbs_per_gpu_all_ranks_concurrent": d2h_gbs_per_gpu,
                             "copy_bound_ms_per_step": (int(getattr(w, "ext_bytes", cells)) + (12 + idx_b) * cells) / d2h_gbs_per_gpu / 1e6,
                             "note": "plain pinned cudaMemcpy D2H, every rank at once: the e2e step is bound by it"}}
    launches_total = int(reduce_sum(dist, launches))

    # ---- roofline of the dominant kernel
    peak, peak_src = peaks()
    kern = max((k for k in ("parse", "bfs", "sweep", "tile_a", "tile_c") if stage_avg.get(k, 0.0) > 0.0), key=lambda k: stage_avg[k])
    alg_bytes = dict(ALG_BYTES)
    stage_avg = {k: v for k, v in stage_avg.items() if v > 0.0}  # stages that did not run in this path report 0
    fused = world == 1 and not args.no_fuse and "parse" not in stage_avg and "tile_a" in stage_avg
    if fused:
        alg_bytes.update(ALG_BYTES_FUSED)
        alg_bytes["tile_c"] += np.dtype(getattr(w, "idx_dtype", np.int32)).itemsize - 4
    achieved = alg_bytes[kern] * cells / (stage_avg[kern] / 1e3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(f"{kern}{'_fused' if fused else ''}_{args.size}")
    except Exception:
        pass
    per_kernel = {}
    for k in ("parse", "tile_a", "tile_c", "bfs", "sweep"):
        if stage_avg.get(k):
            gbs = alg_bytes[k] * cells / (stage_avg[k] / 1e3) / 1e9
            per_kernel[KERNEL_NAMES[k]] = {"algorithmic_bytes_per_cell": alg_bytes[k], "ms": stage_avg[k], "achieved": gbs,
                                           "frac": gbs / peak}
    roofline = {"bound": "hbm", "kernel": KERNEL_NAMES[kern], "per_kernel": per_kernel,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_cell": alg_bytes[kern], "fused_parse": fused,
                "kernel_ms": stage_avg[kern], "stage_ms": stage_avg,
                "step": {"algorithmic_bytes_per_cell": STEP_BYTES_FUSED, "unfused_bytes_per_cell": STEP_BYTES_UNFUSED,
                         "achieved_gbs_per_gpu": STEP_BYTES_FUSED * cells / (ms_total / args.steps / 1e3) / 1e9,
                         "frac_per_gpu": STEP_BYTES_FUSED * cells / (ms_total / args.steps / 1e3) / 1e9 / peak}}

    # BASELINE configs 3 / 5 on the same raster (N = 1): Strahler order, float accuflux and HAND by the tile-dataflow sweeps.
    # --extras sweeps adds the level replays over the BFS order, --extras all every widened entry point.
    extra = None
    if world == 1 and args.extras != "none":
        mode = args.extras or "tile"
        extra = extras(w, args.size, args.seed, skip_level=mode == "tile")
        t = extra["tile_dataflow"]
        step_ms = ms_total / args.steps
        hp = extra["hand_pathsum"]
        hand_ms = hp["hand_ms"] if hp["engine"] == 1 else t["hand_ms"]  # what a pfd_hand call costs with the default options
        extra["baseline_configs"] = {
            "config3_accuflux_plus_strahler_ms": step_ms + t["strahler_ms"],
            "config3_mcells_s": cells / ((step_ms + t["strahler_ms"]) / 1e3) / 1e6,
            "config5_accuflux_mask_hand_ms": step_ms + hand_ms,
            "config5_mcells_s": cells / ((step_ms + hand_ms) / 1e3) / 1e6,
            "config5_hand_engine": "path sums (verified)" if hp["engine"] == 1 else "tile-dataflow sweep",
            "note": f"this {args.size}^2 raster: headline step (parse + rank + accuflux + basins) + the sweep; BASELINE quotes "
                    "config 5 at 16384^2 (run with --size 16384)"}
        if mode == "all":
            extra["widened"] = extras_widened(w, args.size, args.seed, min(args.size, 2048))

    # ---- parity of what was just timed (size-independent properties at full size; N > 1: the same NCCL path on a
    # reduced raster against a single-GPU solve of that raster)
    parity = parity_check(w, dist, rank, world, device, args)

    # ---- CPU baseline (rank 0, N = 1 only): the reference's own numba path when oracle/_ref is there
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        mods = cpu_modules()
        cpu_warm(mods)
        d8, _ = host_raster(args.cpu_size, args.seed)
        st = min((cpu_path(mods, d8) for _ in range(3)), key=lambda t: t["total"])  # best of 3
        cpu = {"value": d8.size / st["total"] / 1e6, "unit": UNIT, "cores": 1, "kind": mods[0],
               "sample": cpu_describe(mods, args.cpu_size, st),
               "accuflux_plus_basins_mcells_s": d8.size / (st["accuflux"] + st["basins"]) / 1e6}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(args, world, w),
            "e2e": e2e, "parity_ok": parity["ok"], "parity": parity,
            "gpu_launches": launches_total, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0 if parity["ok"] else 3


def gather_objects(dist, obj):
    if dist is None:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def parity_check(w, dist, rank, world, device, args):
    """Checks on the outputs the timed region produced (device-resident arm).
    (a) full size, every N: conservation laws that tie the four outputs together across ALL ranks --
        sum of upstream_area over the pits (rank == 0) == number of cells that drain to a pit (rank >= 0);
        number of pits == max basin id == global pit count.
    (b) N > 1: the same pfd_d8_flow_all_tiled / NCCL path on a reduced raster (world x 1024 rows x 4096 columns, or the
        full raster with --verify-full) against rank 0's single-GPU pfd_d8_flow_all of the WHOLE raster, by position-
        dependent checksums of idxs_ds / rank / upstream_area / basins (additive over row blocks)."""
    res = {"ok": True}
    w.step_resident()
    # (a)
    host = []
    for p, dt in zip(w.out_dev[1:], (np.int32, np.int32, np.uint32)):
        a = np.empty(w.cells, dt)
        w.ck(w.l.pfd_memcpy(w.h, w.L.ptr(a), p, w.cells * 4))
        host.append(a)
    rk, upa, bas = host
    is_pit = rk == 0
    mine = {"ranked": int(np.count_nonzero(rk >= 0)), "pit_upa": int(upa[is_pit].astype(np.int64).sum()),
            "pits": int(np.count_nonzero(is_pit)), "max_basin": int(bas.max()),
            "npits_reported": int(getattr(w, "n_pits_global", C.c_int64(-1)).value)}
    del host, rk, upa, bas, is_pit
    allr = gather_objects(dist, mine)
    ranked, pit_upa, pits = (sum(r[k] for r in allr) for k in ("ranked", "pit_upa", "pits"))
    max_basin = max(r["max_basin"] for r in allr)
    res["conservation"] = {"cells_draining_to_pits": ranked, "sum_uparea_at_pits": pit_upa, "pits": pits, "max_basin_id": max_basin}
    ok = ranked == pit_upa and pits == max_basin and ranked > 0
    if world > 1:
        ok = ok and all(r["npits_reported"] == pits for r in allr)
    res["ok"] = bool(ok)
    # (b)
    if world > 1:
        full = args.verify_full
        if full:
            cs = w.checksums(index_offset=w.row0 * w.ncol)
            shape = (w.nrow_global, w.ncol)
            w.free()
        else:
            rows = 1024
            shape = (world * rows, 4096)
            tw = TiledWorkload(shape[1], args.seed + 1, device, rank, world, dist, strong=False, rows_per_rank=rows)
            tw.step_resident()
            cs = tw.checksums(index_offset=tw.row0 * tw.ncol)
            tw.free()
        allc = gather_objects(dist, cs)
        total = [sum(c[i] for c in allc) % (1 << 64) for i in range(4)]
        want = [None]
        if rank == 0:
            sw = Workload(shape[1], args.seed if full else args.seed + 1, device, nrow=shape[0])
            sw.ck(sw.l.pfd_set_option(sw.h, b"tiles", 1))
            sw.step_resident()
            want = [sw.checksums()]
            sw.free()
        if dist is not None:
            dist.broadcast_object_list(want, src=0)
        match = total == want[0]
        res["vs_single_gpu"] = {"raster": f"{shape[0]}x{shape[1]}", "full_size": bool(full), "outputs": ["idxs_ds", "rank", "uparea", "basins"],
                                "checksums_match": bool(match)}
        res["ok"] = bool(res["ok"] and match)
    return res


if __name__ == "__main__":
    sys.exit(main())
