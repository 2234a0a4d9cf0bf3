/*
 * pfd_b200.h -- C ABI of libpfd_b200.so: the B200 (sm_100a) implementation of pyflwdir's D8 flow-network
 * hot path (parse -> order -> accumulate / delineate / stream order / HAND).
 *
 * The reference (Deltares/pyflwdir v0.5.12) has no FFI: its seam is the Python call boundary between the L3
 * object API and the L1/L2 numba kernels. Each entry point below replaces one of those numba
 * kernels; the reference interface it stands in for is cited as (file:line) relative to /root/reference.
 * INTEGRATION.md shows the ctypes binding a maintainer would add on the reference side.
 *
 * Conventions
 *  - plain C: pointers + sizes only. Every array argument may be a HOST pointer (pageable or pinned) or a
 *    DEVICE pointer of the handle's device; the library detects which (cudaPointerGetAttributes) and stages
 *    host buffers itself. Outputs are written in full; calls are synchronous on return.
 *  - return value: 0 = PFD_OK, otherwise a pfd_status code; pfd_last_error(h) has the message. There is NO
 *    CPU fallback: without a usable CUDA device every compute call fails with PFD_ERR_CUDA.
 *  - one handle = one raster on one GPU (owns its device buffers and a CUDA stream). Calls on different
 *    handles are independent; calls on one handle must not overlap.
 *  - cell indices are row-major linear indices, exactly the reference's `idxs_ds` convention
 *    (pit: idxs_ds[i]==i, nodata: idxs_ds[i]==mv with mv=-1 / 0xFFFFFFFF; pyflwdir/core.py:12,
 *    pyflwdir/flwdir.py:112-117).
 */
#ifndef PFD_B200_H
#define PFD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pfd_handle pfd_handle;

typedef enum pfd_status {
    PFD_OK = 0,
    PFD_ERR_CUDA = 1,        /* CUDA runtime / driver error (includes "no device") */
    PFD_ERR_INVALID_ARG = 2, /* bad pointer, size, dtype or flag */
    PFD_ERR_INVALID_D8 = 3,  /* raster holds a value outside core_d8._all (pyflwdir/core_d8.py:19) */
    PFD_ERR_NO_PITS = 4,     /* "Invalid FlwdirRaster: no pits found" (pyflwdir/flwdir.py:126-127) */
    PFD_ERR_STATE = 5,       /* call sequence error (e.g. sweep before parse) */
    PFD_ERR_UNSUPPORTED = 6, /* e.g. more than 2^32 cells, downstream index outside the 8 neighbours */
    PFD_ERR_OOM = 7,         /* device or pinned-host allocation failed */
    PFD_ERR_NCCL = 8         /* NCCL failure in the multi-GPU path */
} pfd_status;

/* element types of caller arrays */
typedef enum pfd_dtype {
    PFD_I8 = 0, PFD_U8 = 1, PFD_I16 = 2, PFD_U16 = 3, PFD_I32 = 4, PFD_U32 = 5,
    PFD_I64 = 6, PFD_U64 = 7, PFD_F32 = 8, PFD_F64 = 9
} pfd_dtype;

/* which cached array pfd_fetch copies out */
typedef enum pfd_array {
    PFD_ARR_IDXS_DS = 0,   /* idx dtype, N          -- core_d8.from_array()[0]      */
    PFD_ARR_PITS = 1,      /* idx dtype, n_pits     -- core_d8.from_array()[1] / core.pit_indices */
    PFD_ARR_PIT_IS_OUTLET = 2, /* uint8, n_pits: 1 where the pit's D8 code is 0/255 (pyflwdir.py:193) */
    PFD_ARR_SEQ = 3,       /* idx dtype, nnodes     -- core.idxs_seq ("walk")        */
    PFD_ARR_RANK = 4,      /* int32, N              -- core.rank()[0]                */
    PFD_ARR_N_UPSTREAM = 5,/* int8, N               -- core.upstream_count           */
    PFD_ARR_D8 = 6,        /* uint8, N              -- core_d8.to_array              */
    PFD_ARR_LEVEL_OFFSETS = 7, /* int64, nlevels+1: start of every rank level inside SEQ */
    PFD_ARR_LDD = 8,       /* uint8, N              -- core_ldd.to_array             */
    PFD_ARR_SUBBASIN_OUTLETS = 9, /* idx dtype, n cells selected by the last pfd_subbasins_* / pfd_inflow_idxs / pfd_outflow_idxs /
                                   * pfd_region_outlets call (their index-array return value) */
    PFD_ARR_REGION_LABELS = 10, /* int64, n labels of the last pfd_region_outlets / pfd_region_slices call */
    PFD_ARR_REGION_SLICES = 11, /* int32 [n][4] (row start, row stop, col start, col stop) of the last pfd_region_slices call */
    PFD_ARR_NEXTXY = 12,        /* int32 [2][N] -- core_nextxy.to_array (nextx plane, then nexty plane) */
    PFD_ARR_STREAM_OFFSETS = 13,/* int64 [n_streams + 1]: start of every segment of the last pfd_streams call inside STREAM_CELLS */
    PFD_ARR_STREAM_CELLS = 14   /* idx dtype [n_cells]: the segments of the last pfd_streams call, back to back */
} pfd_array;

/* ---- library / device ------------------------------------------------------------------------------- */
const char* pfd_version(void);
int pfd_device_count(void);                       /* 0 when no CUDA device is usable */
const char* pfd_status_string(int status);
/* PCI bus id ("0000:1b:00.0") of CUDA device `device`: lets a one-process-per-GPU host bind itself to the GPU's NUMA node
 * before it allocates pinned buffers (pyflwdir_b200/_lib.py bind_to_device_numa_node). */
int pfd_device_pci_bus_id(int device, char* out, int capacity);

/* ---- handle ------------------------------------------------------------------------------------------ */
int pfd_create(int device, pfd_handle** out);
void pfd_destroy(pfd_handle* h);
const char* pfd_last_error(const pfd_handle* h);  /* message of the last failing call on h (h may be NULL) */

/* Pinned host memory + device memory for callers that want resident / zero-staging buffers. */
int pfd_host_alloc(size_t bytes, void** out);
int pfd_host_free(void* p);
int pfd_dev_alloc(pfd_handle* h, size_t bytes, void** out);
int pfd_dev_free(pfd_handle* h, void* p);
int pfd_memcpy(pfd_handle* h, void* dst, const void* src, size_t bytes); /* any direction, synchronous */
int pfd_synchronize(pfd_handle* h);

/* ---- parse ------------------------------------------------------------------------------------------- */
/*
 * Replaces core_d8.from_array (pyflwdir/core_d8.py:42-67) + core_d8.isvalid/check_values (:105-122), as
 * called from pyflwdir.from_array (pyflwdir/pyflwdir.py:183-193).
 *   d8          : nrow x ncol uint8, C-contiguous (host or device)
 *   check_values: !=0 -> fail with PFD_ERR_INVALID_D8 if a value is not one of the 11 legal codes. Illegal
 *                 codes are refused even when 0 (the reference would silently mis-parse them).
 *   idxs_ds_out : optional (may be NULL). If given, the downstream-index array is ALSO written here in
 *                 `idx_dtype` (PFD_I32 / PFD_U32 / PFD_I64) by the same kernel that parses the raster.
 *   n_valid / n_pits / n_outlets : counts of non-nodata cells, pits, pits whose code is 0/255.
 * The handle keeps the parsed flow graph on the device (1 B direction + 1 B upstream mask per cell).
 */
int pfd_d8_parse(pfd_handle* h, const uint8_t* d8, int64_t nrow, int64_t ncol, int check_values,
                 void* idxs_ds_out, int idx_dtype, int64_t* n_valid, int64_t* n_pits, int64_t* n_outlets);

/* core_ldd.from_array (pyflwdir/core_ldd.py:41-66): PCRaster LDD codes 1..9 (5 = pit), 255 = nodata; same kernel. */
int pfd_ldd_parse(pfd_handle* h, const uint8_t* ldd, int64_t nrow, int64_t ncol, void* idxs_ds_out, int idx_dtype,
                  int64_t* n_valid, int64_t* n_pits);

/* core_nextxy.from_array (pyflwdir/core_nextxy.py:24-67): CaMa-Flood NEXTXY raster = two int32 planes with the one-based
 * column (nextx) and row (nexty) of the downstream cell; -9 / -10 = pit, -9999 = nodata. check_values != 0: fail with
 * PFD_ERR_INVALID_D8 unless core_nextxy.isvalid (:86-103) holds. n_outlets = pits whose nextx is -9 / -10
 * (pyflwdir.py:193). A link that leaves the 8 neighbours of its cell is refused with PFD_ERR_UNSUPPORTED (the device
 * graph stores the slot of the downstream neighbour in one byte). pfd_fetch(PFD_ARR_NEXTXY) is core_nextxy.to_array. */
int pfd_nextxy_parse(pfd_handle* h, const int32_t* nextx, const int32_t* nexty, int64_t nrow, int64_t ncol, int check_values,
                     void* idxs_ds_out, int idx_dtype, int64_t* n_valid, int64_t* n_pits, int64_t* n_outlets);

/*
 * Constructor path FlwdirRaster(idxs_ds, shape, "d8", ...) (pyflwdir/pyflwdir.py:211-273): load a graph from
 * a downstream-index array whose links all stay inside the 8-neighbourhood.
 */
int pfd_load_idxs_ds(pfd_handle* h, const void* idxs_ds, int idx_dtype, int64_t nrow, int64_t ncol,
                     int64_t* n_valid, int64_t* n_pits);

/* ---- order -------------------------------------------------------------------------------------------- */
/*
 * Replaces core.idxs_seq (pyflwdir/core.py:87-117, incl. upstream_matrix :67-84) and core.rank (:17-47):
 * a level-synchronous BFS from the pits that reproduces the reference's "walk" sequence exactly and yields
 * rank (= level) on the way. Idempotent; every sweep calls it implicitly.
 *   nnodes  : number of cells that drain to a pit (= seq.size = #(rank >= 0))
 *   nlevels : max rank + 1
 */
int pfd_order(pfd_handle* h, int64_t* nnodes, int64_t* nlevels);

/* Copy a cached array out (host or device destination); the pfd_array entries above name the reference array each one
 * is (core_d8.from_array, core.idxs_seq, core.rank, core.upstream_count, core_d8 / core_ldd / core_nextxy.to_array under
 * pyflwdir/). idx_dtype is used for the index-typed arrays. */
int pfd_fetch(pfd_handle* h, int which, void* out, int idx_dtype);

/* ---- sweeps ------------------------------------------------------------------------------------------ */
/*
 * streams.accuflux / accuflux_ds (pyflwdir/streams.py:15-41, 44-70) over the "walk" sequence.
 *   data, out : N elements of `dtype`; nodata compared as numba does (integer vs integer as int64, anything
 *               with a float as float64): pass both representations + is_int.
 *   direction : 0 = "up" (accumulate upstream values), 1 = "down".
 * Bit-exact for floats: a cell pulls its upstream neighbours in descending linear index.
 */
int pfd_accuflux(pfd_handle* h, const void* data, int dtype, double nodata_f, int64_t nodata_i,
                 int nodata_is_int, int direction, void* out);

/* FlwdirRaster.upstream_area(unit="cell") (pyflwdir/pyflwdir.py:770-801): int32 cell counts, -9999 on nodata. */
int pfd_upstream_area_cells(pfd_handle* h, int32_t* out);

/*
 * basins.basins (pyflwdir/basins.py:12-18) -> core.fillnodata_upstream (pyflwdir/core.py:120-146).
 *   outlets == NULL : all pits, ids 1..n_pits (uint32) -- `ids`/`ids_dtype` ignored, out is uint32.
 *   else            : n_outlets linear indices (idx_dtype) with ids of `ids_dtype` (any 1/2/4/8-byte integer).
 */
int pfd_basins(pfd_handle* h, const void* outlets, int64_t n_outlets, int idx_dtype, const void* ids,
               int ids_dtype, void* out);

/* streams.strahler_order (pyflwdir/streams.py:228-269); mask may be NULL, else N bytes (non-zero = stream). */
int pfd_strahler(pfd_handle* h, const uint8_t* mask, uint8_t* out);

/*
 * dem.height_above_nearest_drain (pyflwdir/dem.py:299-330): drain = N bytes (==1 marks a drain cell), elevtn
 * = N values of PFD_F32 or PFD_F64; out = N float64, -9999.0 outside the sequence.
 */
int pfd_hand(pfd_handle* h, const uint8_t* drain, const void* elevtn, int elev_dtype, double* out);

/* ---- next rows: same sweeps, other functors ------------------------------------------------ */
/* core.fillnodata_upstream / fillnodata_downstream (pyflwdir/core.py:120-146, 149-188), Flwdir.fillnodata
 * (pyflwdir/flwdir.py:360-392). direction 0 = "up" (fill from the first downstream valid cell), 1 = "down"
 * (fill from upstream, merging at confluences with how = 0 max / 1 min / 2 sum). */
int pfd_fillnodata(pfd_handle* h, const void* data, int dtype, double nodata_f, int64_t nodata_i, int nodata_is_int,
                   int direction, int how, void* out);
/* core.main_upstream (pyflwdir/core.py:191-219): index of the upstream cell with the largest uparea (> upa_min),
 * mv (-1) where there is none. uparea: int32 / uint32 / int64 / float32 / float64. */
int pfd_main_upstream(pfd_handle* h, const void* uparea, int dtype, double upa_min, void* out, int idx_dtype);
/* core.upstream_count (pyflwdir/core.py:50-61) with an optional mask of the cells that count (N bytes or NULL) */
int pfd_upstream_count(pfd_handle* h, const uint8_t* mask, int8_t* out);
/* core.upstream_matrix (pyflwdir/core.py:67-84): out[N][d] = upstream cells of every cell in ascending linear index, padded
 * with mv (-1 / all ones); d = largest in-degree, returned in *d_out. out == NULL: size query only. */
int pfd_upstream_matrix(pfd_handle* h, void* out, int idx_dtype, int64_t d_capacity, int64_t* d_out);
/* streams.stream_order, "classic" / Hack (pyflwdir/streams.py:191-225); mask may be NULL */
int pfd_stream_order_classic(pfd_handle* h, const void* idxs_us_main, int idx_dtype, const uint8_t* mask, uint8_t* out);

/* streams.stream_distance (pyflwdir/streams.py:272-315, interpreted Python in the reference): distance to the
 * outlet or to the next downstream cell of `mask` (N bytes or NULL). real_length != 0: float32 metres, hop lengths
 * from hop_table[nrow][3][2] = float32(gis_utils.distance) per (row, row delta + 1, |column delta|) built by the
 * host; else int32 cell counts. out is -9999 outside the sequence. */
int pfd_stream_distance(pfd_handle* h, const uint8_t* mask, int real_length, const float* hop_table, void* out);

/* dem.floodplains (pyflwdir/dem.py:333-379, interpreted Python in the reference): floodplain mask from a HAND
 * threshold that scales with the drain's upstream area (h ~ A**b). drainh_init: N float32 holding
 * float32(uparea ** b) at the drain cells (uparea >= upa_min) and -9999 elsewhere -- the power is evaluated by the
 * host exactly as the reference does; elevtn: N float32 / float64; out: N int8 (-1 outside the sequence). */
int pfd_floodplains(pfd_handle* h, const float* drainh_init, const void* elevtn, int elev_dtype, int8_t* out);

/* arithmetics.upstream_sum (pyflwdir/arithmetics.py:150-169), Flwdir.upstream_sum (pyflwdir/flwdir.py:412-433): per cell
 * the sum of the values of its direct upstream neighbours, with the reference's nodata rule (a cell whose own or whose
 * downstream value is nodata gets nodata ASSIGNED when the scan reaches it, which wipes earlier additions). data / out:
 * N elements of `dtype` (any 1/2/4/8-byte integer, float32, float64); nodata as in pfd_accuflux. Bit-exact for floats. */
int pfd_upstream_sum(pfd_handle* h, const void* data, int dtype, double nodata_f, int64_t nodata_i, int nodata_is_int,
                     void* out);
/* basins.subbasins_streamorder (pyflwdir/basins.py:67-103), FlwdirRaster.subbasins_streamorder (pyflwdir/pyflwdir.py:601-629):
 * subbasins of every stream segment with order >= min_sto (min_sto < 0: relative to the maximum order). strord: N uint8;
 * mask: N bytes or NULL; subbas_out: N int32 (0 = no subbasin); ids follow the reference (outlets numbered along
 * seq[::-1]). The outlet cells (2nd return value) are fetched with pfd_fetch(PFD_ARR_SUBBASIN_OUTLETS). */
int pfd_subbasins_streamorder(pfd_handle* h, const uint8_t* strord, const uint8_t* mask, int64_t min_sto, int32_t* subbas_out,
                              int64_t* n_outlets);

/* basins.subbasins_area (pyflwdir/basins.py:194-233), FlwdirRaster.subbasins_area (pyflwdir/pyflwdir.py:665-692): moving
 * upstream from the outlets a new subbasin starts at tributaries / interbasins with more than area_min contributing area.
 * idxs_us_main: N indices (core.main_upstream, e.g. from pfd_main_upstream); uparea: N int32 / int64 / float32 / float64;
 * subbas_out: N uint32; outlet cells via pfd_fetch(PFD_ARR_SUBBASIN_OUTLETS). */
int pfd_subbasins_area(pfd_handle* h, const void* idxs_us_main, int idx_dtype, const void* uparea, int dtype, double area_min,
                       uint32_t* subbas_out, int64_t* n_outlets);

/* arithmetics.moving_average / moving_median (pyflwdir/arithmetics.py:67-147) over core._window (pyflwdir/core.py:368-398),
 * Flwdir.moving_average / moving_median (pyflwdir/flwdir.py:435-505): per cell the weighted mean / nan-median over the n
 * cells upstream along idxs_us_main, the cell itself and the n cells downstream (strord != NULL: downstream only while the
 * stream order does not grow). data / out: N float32 or float64; weights: NULL (ones) or N float64 (wdtype = PFD_F64, the only dtype the
 * reference's numba typing accepts); 0 <= n <= 64;
 * nodata may be NaN. Bit-exact (float64 accumulation in window order; numba's own quick-select for the median). */
int pfd_moving_average(pfd_handle* h, const void* data, int dtype, const void* weights, int wdtype, int n,
                       const void* idxs_us_main, int idx_dtype, const uint8_t* strord, double nodata, void* out);
int pfd_moving_median(pfd_handle* h, const void* data, int dtype, int n, const void* idxs_us_main, int idx_dtype,
                      const uint8_t* strord, double nodata, void* out);

/* ---- local traces and region post-processing ---------------------------------------------------------- */
/* Flwdir.downstream (pyflwdir/flwdir.py:394-410): out[i] = data[idxs_ds[i]] for cells with a downstream link, data[i] for
 * pits and nodata cells. data / out: N values of `dtype` (any of pfd_dtype; moved by size). */
int pfd_downstream(pfd_handle* h, const void* data, int dtype, void* out);

/* core.path / core.snap (pyflwdir/core.py:400-480) over core._trace (:309-364): from every start cell follow the
 * downstream links (direction = 0) or idxs_us_main (direction = 1; N indices of idx_dtype) until a pit / headwater, a
 * mask cell (mask: NULL or N uint8, the hit is included) or until the next hop would exceed max_length
 * (has_max_length = 0 <=> None). hop_table: NULL (unit "cell", every hop counts 1.0) or nrow*3*2 float64 =
 * gis_utils.distance (pyflwdir/gis_utils.py:452-486) per (row of the cell, row delta + 1, column delta != 0).
 * starts: n0 int64. Per start: counts_out (cells in the trace, start included), ends_out (last cell = core.snap's
 * index), dists_out (float64; core.snap casts to float32); any of them may be NULL. paths_out: NULL, or room for
 * paths_capacity >= sum(counts) indices of path_dtype, traces back to back in start order (core.path). Fails with
 * PFD_ERR_INVALID_ARG for an index outside the raster and PFD_ERR_UNSUPPORTED for a trace that never ends. */
int pfd_trace(pfd_handle* h, const int64_t* starts, int64_t n0, int direction, const void* idxs_us_main, int idx_dtype,
              const uint8_t* mask, int has_max_length, double max_length, const double* hop_table, int64_t* counts_out,
              int64_t* ends_out, double* dists_out, void* paths_out, int path_dtype, int64_t paths_capacity);

/* core.inflow_idxs / core.outflow_idxs (pyflwdir/core.py:483-514): most upstream cells draining INTO / most downstream
 * cells INSIDE the region (N uint8), in the reference's order (seq[::-1] / seq). *n_out = count; the indices are
 * fetched with pfd_fetch(PFD_ARR_SUBBASIN_OUTLETS). */
int pfd_inflow_idxs(pfd_handle* h, const uint8_t* region, int64_t* n_out);
int pfd_outflow_idxs(pfd_handle* h, const uint8_t* region, int64_t* n_out);

/* basins.interbasin_mask (pyflwdir/basins.py:23-64): most downstream contiguous area within region (N uint8), optionally
 * reduced to the cells that drain to a stream cell (stream: NULL or N uint8). out: N uint8 (0 / 1). */
int pfd_interbasin_mask(pfd_handle* h, const uint8_t* region, const uint8_t* stream, uint8_t* out);

/* regions.region_outlets (pyflwdir/regions.py:132-163): outlet cell(s) of every region label > 0, sorted by label with
 * numba's argsort (ties keep the reference's order). regions: N values of PFD_I32 / PFD_U32 / PFD_I64 / PFD_U64.
 * *n_out = count; pfd_fetch(PFD_ARR_REGION_LABELS) -> int64 labels, pfd_fetch(PFD_ARR_SUBBASIN_OUTLETS) -> cells. */
int pfd_region_outlets(pfd_handle* h, const void* regions, int dtype, int64_t* n_out);

/* regions.region_slices (pyflwdir/regions.py:58-86, scipy.ndimage.find_objects): bounding rows / columns of every label
 * > 0 present in regions (dtype as above; labels up to 2^28). *n_labels = count; pfd_fetch(PFD_ARR_REGION_LABELS) ->
 * ascending int64 labels (np.unique), pfd_fetch(PFD_ARR_REGION_SLICES) -> int32 [n][4]. region_bounds (:89-129) turns
 * these into coordinates on the host. */
int pfd_region_slices(pfd_handle* h, const void* regions, int dtype, int64_t* n_labels);

/* streams.streams (pyflwdir/streams.py:131-188): linear indices per stream segment between two confluences of the masked
 * network (mask: NULL = all cells, or N uint8), in the reference's order (segment starts found while walking seq[::-1]);
 * segments longer than max_len > 0 cells are split into pieces that share their end points; a zero-length [pit, pit]
 * segment follows every segment that ends in a pit. *n_streams / *n_cells = number of segments / of indices; fetch with
 * pfd_fetch(PFD_ARR_STREAM_OFFSETS) (int64 [n_streams + 1]) and pfd_fetch(PFD_ARR_STREAM_CELLS) (idx dtype). */
int pfd_streams(pfd_handle* h, const uint8_t* mask, int64_t max_len, int64_t* n_streams, int64_t* n_cells);

/* basins.subbasins_pfafstetter (pyflwdir/basins.py:106-191): Pfafstetter coding of every basin down to `depth` levels
 * (1..8): per label the four largest tributaries (odd digits) and the interbasins between them (even digits), on the
 * classic stream order of the masked network (mask: NULL or N uint8, e.g. uparea >= upa_min). idxs_us_main: N indices;
 * uparea: N values of PFD_I32 / PFD_I64 / PFD_F32 / PFD_F64. pfafbas_out: N int64 (the reference's int32 % int64).
 * *n_outlets = number of subbasin outlet cells, fetched with pfd_fetch(PFD_ARR_SUBBASIN_OUTLETS) in the reference's
 * order (pits first, then label by label). Ties between equal upstream areas are broken like numba's argsort does. */
int pfd_subbasins_pfafstetter(pfd_handle* h, const void* idxs_us_main, int idx_dtype, const void* uparea, int dtype,
                              const uint8_t* mask, int depth, int64_t* pfafbas_out, int64_t* n_outlets);

/* rivers.classify_estuary (pyflwdir/rivers.py:11-53): estuaries by width convergence. est_init: N int8, 1 at the pits with
 * elevtn <= max_elevtn (selected by the caller, rivers.py:38-39), 0 elsewhere; rivdst / rivwth: N float32 or float64;
 * out: N int8 (>= 1 where estuary, 2 at the upstream end). */
int pfd_classify_estuary(pfd_handle* h, const int8_t* est_init, const void* rivdst, int dst_dtype, const void* rivwth,
                         int wth_dtype, double min_convergence, int8_t* out);

/* ---- fused headline pass ------------------------------------------------------------------------------ */
/*
 * parse + order + rank + upstream_area(cell) + basins() in one call (BASELINE.json metric): core_d8.from_array
 * (pyflwdir/core_d8.py:42-67), core.rank (pyflwdir/core.py:17-47), streams.accuflux of ones with -9999 on nodata
 * (pyflwdir/streams.py:15-41, pyflwdir/pyflwdir.py:770-801) and basins.basins over all pits (pyflwdir/basins.py:12-18).
 * Any output may be NULL. Equivalent to pfd_d8_parse + pfd_order + pfd_fetch(RANK) + pfd_upstream_area_cells +
 * pfd_basins(NULL).
 */
int pfd_d8_flow_all(pfd_handle* h, const uint8_t* d8, int64_t nrow, int64_t ncol, void* idxs_ds_out,
                    int idx_dtype, int32_t* rank_out, int32_t* uparea_out, uint32_t* basins_out,
                    int64_t* n_valid, int64_t* n_pits, int64_t* nnodes);

/* ---- row-tiled multi-GPU solve (BASELINE config 4) ------------------------------------- */
/*
 * One process per GPU; rank g owns a block of consecutive rows (a multiple of 64 rows except for the last rank)
 * of ONE raster. d8_block holds halo_top + nrow_owned + halo_bot rows: the owned rows plus one halo row of D8
 * codes from each existing neighbour (core_d8.from_array's "downstream cell is nodata?" test needs it).
 * The solve is the tile solver of pfd_d8_flow_all with two exchanges: the pit counts (global basin ids) and ONE
 * all-reduce (uint32 sum) of the boundary tables, 4 x 2(R-1) x ncol entries. Outputs hold the owned rows only;
 * idxs_ds are GLOBAL linear indices. Integer outputs => bit-identical to the single-GPU result.
 *
 * pfd_d8_flow_all_tiled does everything over the handle's NCCL communicator (pfd_comm_init). The three step
 * functions expose the same computation with the exchanges left to the caller (tests emulate R ranks on one GPU):
 *   pfd_tiled_parse  -> n_pits of the block;   caller: pit_id_offset = sum of n_pits of the lower ranks
 *   pfd_tiled_local  -> *table_dev (device, uint32 x *table_len);   caller: all-reduce(sum) it across ranks
 *   pfd_tiled_finish -> outputs.   basins_out (if wanted) must be a device buffer, same pointer in both calls.
 */
int pfd_comm_unique_id(void* out128, int64_t capacity);     /* ncclGetUniqueId -> 128 bytes, on rank 0 */
int pfd_comm_init(pfd_handle* h, int rank, int nranks, const void* unique_id128);
int pfd_comm_barrier(pfd_handle* h);
int pfd_comm_destroy(pfd_handle* h);
int pfd_d8_flow_all_tiled(pfd_handle* h, const uint8_t* d8_block, int64_t nrow_owned, int64_t ncol, int halo_top,
                          int halo_bot, int64_t glob_row0, void* idxs_ds_out, int idx_dtype, int32_t* rank_out,
                          int32_t* uparea_out, uint32_t* basins_out, int64_t* n_valid, int64_t* n_pits_global);
int pfd_tiled_parse(pfd_handle* h, const uint8_t* d8_block, int64_t nrow_owned, int64_t ncol, int halo_top,
                    int halo_bot, int64_t glob_row0, void* idxs_ds_out, int idx_dtype, int64_t* n_valid,
                    int64_t* n_pits);
int pfd_tiled_local(pfd_handle* h, int rank, int nranks, int64_t pit_id_offset, uint32_t* basins_out,
                    void** table_dev, int64_t* table_len);
int pfd_tiled_finish(pfd_handle* h, int32_t* rank_out, int32_t* uparea_out, uint32_t* basins_out);

/* ---- row-block (multi-GPU) sweeps of the order-sensitive outputs ------------------------------------------- */
/*
 * streams.strahler_order (kind 0, unmasked; pyflwdir/streams.py:228-269), streams.accuflux up (kind 1, any dtype;
 * streams.py:15-41) and dem.height_above_nearest_drain (kind 2; pyflwdir/dem.py:299-330) across the row blocks of ONE
 * raster, on handles that hold a row block (pfd_tiled_parse / pfd_d8_flow_all_tiled). These outputs cannot be
 * re-associated, so every rank sweeps its block extended by the neighbours' edge rows and the ranks swap the values +
 * done flags of their edge rows between rounds until a round resolves nothing anywhere (SURVEY.md §8e "neighbour
 * send/recv rounds"). Bit-identical to the single-GPU call. data: accuflux data / HAND elevtn of the OWN rows; drain:
 * HAND only. pfd_sweep_tiled drives the rounds over NCCL (ncclSend / ncclRecv between row neighbours + one all-reduce
 * of the progress counter per round); the step functions let a caller emulate the exchange (tests):
 *   begin -> swap(edges -> halo) -> { round -> swap(edges -> halo) }* -> end
 * pfd_sweep_tiled_edges(which = 0 my first row | 1 my last row) packs [dir | done | value | aux] of that row;
 * pfd_sweep_tiled_halo(which = 0 the row above my block | 1 the row below) installs a neighbour's record.
 * Rasters with loops are refused by the up-sweeps (PFD_ERR_UNSUPPORTED): use the single-GPU call. The HAND down-sweep starts at
 * every drain cell, so a drain cell that drains to no pit (above a loop; outside the reference's sequence) gets 0 instead of
 * -9999 here: callers with such rasters use the single-GPU pfd_hand (pyflwdir_b200.multigpu does, from the rank it already has).
 */
int pfd_sweep_tiled_begin(pfd_handle* h, int kind, const void* data, int dtype, const uint8_t* drain, double nodata_f,
                          int64_t nodata_i, int nodata_is_int);
int pfd_sweep_tiled_round(pfd_handle* h, int64_t* newly_resolved);
int pfd_sweep_tiled_edges(pfd_handle* h, int which, void** buf_dev, int64_t* nbytes);
int pfd_sweep_tiled_halo(pfd_handle* h, int which, const void* record);
int pfd_sweep_tiled_end(pfd_handle* h, void* out, int64_t* resolved);
int pfd_sweep_tiled(pfd_handle* h, int kind, const void* data, int dtype, const uint8_t* drain, double nodata_f,
                    int64_t nodata_i, int nodata_is_int, void* out, int64_t* rounds);

/* ---- the step before the path ----------------------------------------------------------- */
/*
 * dem.fill_depressions (pyflwdir/dem.py:17-143; pyflwdir.from_dem, pyflwdir/pyflwdir.py:51-102, is this followed by
 * pfd_d8_parse): depression-filled elevation and the D8 codes of the reference's priority flood, bit for bit. Does not
 * touch the raster parsed on the handle.
 *   elevtn / elevtn_out : nrow*ncol values of PFD_F32 or PFD_F64. Integer rasters are passed as PFD_F64 with int_delv = 1
 *                         (numba's typing of the loop: arithmetic in float64, the raise truncated to the integer type).
 *   outlets_mode        : 0 = outlets="edge", 1 = outlets="min", 2 = idxs_pit (n_pit int64 linear indices).
 *   nodata              : may be NaN. has_elv_max / elv_max: the optional elv_max of outlets="edge".
 *   connectivity        : 4 or 8.  max_depth: must be negative (fill everything); >= 0 -> PFD_ERR_UNSUPPORTED.
 *   stats (may be NULL) : int64[12] = relaxation passes (levels, tie labels), cells in tie components, tie components,
 *                         cells no outlet reaches, initial outlets, final band (float32 bits), largest component, largest
 *                         key drift seen (float32 bits), number of label + replay rounds, microseconds spent on the levels / on the rest.
 * PFD_ERR_INVALID_ARG with the reference's message when elv_max leaves no outlet ("No initial outlet cells found.").
 * Float32 rasters: raising a cell (z1 += z0 - z1) is inexact, the reference's heap keys drift around the pour level; the
 * replay follows them exactly (csrc/pfd_fill.cuh), widening the band of "tied" levels until it covers the drift.
 */
int pfd_fill_depressions(pfd_handle* h, const void* elevtn, int elev_dtype, int64_t nrow, int64_t ncol, int outlets_mode,
                         const int64_t* idxs_pit, int64_t n_pit, double nodata, double max_depth, int has_elv_max,
                         double elv_max, int connectivity, int int_delv, void* elevtn_out, uint8_t* d8_out, int64_t* stats);

/* ---- synthetic input (bench / tests) ---------------------------------------------------- */
/* z: nrow*ncol float32 (device or host) elevation; d8 from z by strict steepest descent. */
int pfd_synth_elevation(pfd_handle* h, int64_t nrow, int64_t ncol, int64_t nref, int octaves, uint32_t seed,
                        float* z_out);
int pfd_synth_d8(pfd_handle* h, const float* z, int64_t nrow, int64_t ncol, float sea_level, uint8_t* d8_out);
/* rows [row0, row0+nrow) of the nrow_global x ncol synthetic D8 raster (row blocks of the multi-GPU bench) */
int pfd_synth_d8_block(pfd_handle* h, int64_t row0, int64_t nrow, int64_t ncol, int64_t nrow_global, int64_t nref,
                       int octaves, uint32_t seed, float sea_level, uint8_t* d8_out);

/* ---- options / introspection ------------------------------------------------------------------------- */
/* "tiles" = 1 (default): rank / basins() / upstream_area("cell") come from the tile-hierarchical solver
 * (pfd_tiles.cuh); 0: from the level-synchronous BFS + sweeps. Results are identical bit for bit.
 * "tile_sweeps" = 1 (default): accuflux / Strahler / HAND fallback by the tile-dataflow sweeps unless the BFS ordering is already
 * cached; 2: always; 0: level replays over the BFS order. "hand_pathsum" = 1 (default): pfd_hand tries the re-associated path sums
 * first (accepted only when every cell satisfies the reference's statement bit for bit, pfd_hand.cuh), and pfd_basins with custom
 * outlets / pfd_fillnodata(direction up) use the same path summaries (no ordering needed); 0: the sweeps only.
 * "fuse_parse", "sweep_max_passes", "release_scratch": see pfd_api.cu. pfd_get_info: "hand_engine" (1 path sums, 2 tile sweep,
 * 3 level replay answered the last pfd_hand), "sweep_passes", "tile_rounds", "nnodes", "n_pits", "nlevels", ... */
int pfd_set_option(pfd_handle* h, const char* name, int64_t value);
/* "tiles", "tile_rounds", "nlevels", "nnodes", "n_pits", "n_valid", "num_sms"; -1 if unknown */
int64_t pfd_get_info(const pfd_handle* h, const char* name);

/* ---- instrumentation ---------------------------------------------------------------------------------- */
/* Position-dependent 64-bit checksum of `count` elements of `elem_bytes` (1, 2, 4 or 8) bytes (device or host array):
 * sum over i of (v[i] + 1) * ((i + index_offset) * 0x9E3779B97F4A7C15 | 1) mod 2^64. Additive over consecutive blocks,
 * so the checksums of the row blocks of a multi-GPU run add up to the checksum of the single-GPU array (bench.py's
 * parity_ok, the full-size tests). No reference counterpart: verification plumbing. */
/* Element-wise verification of FINISHED outputs against their defining recurrences (csrc/pfd_verify.cuh): one independent
 * pass re-evaluates, for every cell, the reference's per-cell statement from the finished values of the cell's graph
 * neighbours and compares bit for bit -- a proof of the whole array at any raster size (the oracle only runs at small sizes).
 * n_bad counts the violating cells (0 = the array is exactly what the reference returns for the raster parsed on h).
 *   pfd_verify_flow     n_bad[5] = idxs_ds (core_d8.py:42-67), rank (core.py:17-47), basins (core.py:120-146), upstream
 *                       area in cells (streams.py:15-41 via pyflwdir.py:790-800), pit numbering (basins.py:14-16); any array may be NULL
 *   pfd_verify_strahler streams.strahler_order (streams.py:228-269)        pfd_verify_hand  dem.height_above_nearest_drain (dem.py:299-330)
 *   pfd_verify_accuflux streams.accuflux, any dtype, in the walk order (descending upstream index) */
int pfd_verify_flow(pfd_handle* h, const void* idxs_ds, int idx_dtype, const int32_t* rank, const int32_t* uparea,
                    const uint32_t* basins, int64_t* n_bad);
int pfd_verify_strahler(pfd_handle* h, const uint8_t* mask, const uint8_t* strord, int64_t* n_bad);
int pfd_verify_hand(pfd_handle* h, const uint8_t* drain, const void* elevtn, int elev_dtype, const double* hand, int64_t* n_bad);
int pfd_verify_accuflux(pfd_handle* h, const void* data, int dtype, double nodata_f, int64_t nodata_i, int nodata_is_int,
                        const void* accu, int64_t* n_bad);
int pfd_checksum(pfd_handle* h, const void* data, int elem_bytes, int64_t count, uint64_t index_offset, uint64_t* out);
/* kernels launched by this handle since creation (bench.py's gpu_launches) */
int64_t pfd_launch_count(const pfd_handle* h);
/* CUDA-event stopwatch on the handle's stream: device time between the two calls, in ms */
int pfd_timer_start(pfd_handle* h);
int pfd_timer_stop(pfd_handle* h, double* ms);
/* device time [ms] (CUDA events on the handle's stream) of the most recent call, by stage:
 * 0 parse kernel, 1 pit compaction (3 kernels), 2 order (init + BFS + bookkeeping), 3 last sweep kernel,
 * 4 whole pfd_d8_flow_all call, 5 BFS kernel alone, 6 / 7 / 8 tile solver phase A / B (all rounds) / C */
double pfd_last_stage_ms(const pfd_handle* h, int stage);

#ifdef __cplusplus
}
#endif
#endif /* PFD_B200_H */
