// pfd_common.cuh -- shared device helpers, the handle, scratch management.
#pragma once

#include <cuda_runtime.h>
#include <cooperative_groups.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <initializer_list>
#include <string>
#include <vector>

#include "../../include/pfd_b200.h"

namespace cg = cooperative_groups;

// ---------------------------------------------------------------------------------------------------------
// Device-side flow graph encoding (HBM layout):
//   dir[i]    uint8: 0..7 = slot of the downstream neighbour, 8 = pit with code 0/255 ("outlet"),
//                    9 = forced pit (flows off the raster or into nodata), 255 = nodata
//   upmask[i] uint8: bit k set <=> the neighbour in slot k drains into cell i
// Neighbour slots are numbered in ASCENDING linear index: 0 NW, 1 N, 2 NE, 3 W, 4 E, 5 SW, 6 S, 7 SE.
// The neighbour in slot k drains into the centre iff its own dir == 7 - k.
// ---------------------------------------------------------------------------------------------------------
#define PFD_DIR_PIT 8
#define PFD_DIR_FPIT 9
#define PFD_DIR_NODATA 255

// D8 code (core_d8.py:15 _ds) of "flows to slot k"
__host__ __device__ __forceinline__ constexpr unsigned pfd_slot_code(int k) {
    // NW 32, N 64, NE 128, W 16, E 1, SW 8, S 4, SE 2
    return (k == 0) ? 32u : (k == 1) ? 64u : (k == 2) ? 128u : (k == 3) ? 16u : (k == 4) ? 1u : (k == 5) ? 8u
                                                                                         : (k == 6) ? 4u : 2u;
}

// PCRaster LDD code (core_ldd.py:13 _ds = [[7,8,9],[4,5,6],[1,2,3]]) of "flows to slot k"; pit = 5, nodata = 255
__host__ __device__ __forceinline__ constexpr unsigned pfd_slot_code_ldd(int k) {
    return (k == 0) ? 7u : (k == 1) ? 8u : (k == 2) ? 9u : (k == 3) ? 4u : (k == 4) ? 6u : (k == 5) ? 1u : (k == 6) ? 2u : 3u;
}
// FT: 0 = D8, 1 = LDD
template <int FT>
__host__ __device__ __forceinline__ constexpr unsigned pfd_code(int k) { return FT == 0 ? pfd_slot_code(k) : pfd_slot_code_ldd(k); }
template <int FT>
__host__ __device__ __forceinline__ constexpr unsigned pfd_nodata_code() { return FT == 0 ? 247u : 255u; }

// row / column delta of slot k (packed 2-bit LUTs: value + 1)
__host__ __device__ __forceinline__ int pfd_slot_dr(int k) { return (int)((0xA940u >> (2 * k)) & 3u) - 1; }
__host__ __device__ __forceinline__ int pfd_slot_dc(int k) { return (int)((0x9224u >> (2 * k)) & 3u) - 1; }
__host__ __device__ __forceinline__ int64_t pfd_slot_off(int k, int64_t ncol) {
    return (int64_t)pfd_slot_dr(k) * ncol + pfd_slot_dc(k);
}

typedef uint32_t cell_t;  // internal cell index: rasters up to 2^32 cells


__device__ __forceinline__ uint32_t splat4(uint32_t b) { return b * 0x01010101u; }

// ---------------------------------------------------------------------------------------------------------
// core_d8.from_array / core_ldd.from_array (core_d8.py:42-67, core_ldd.py:41-66) for FOUR horizontally adjacent cells
// at once with byte-SIMD compares. Inputs: the 3x3 neighbourhood of 32-bit words around the word `w` that holds the
// four codes (a* = row above, b0 / b2 = left / right word, c* = row below); cells outside the raster must read as
// the nodata code, which makes "flows off the raster" and "flows into nodata" the same test (core_d8.py:58-61).
// Outputs: dirw (dir byte per cell), upw (upstream mask per cell). Returns the per-byte mask of LEGAL codes
// (core_d8.py:115-122): 0xFFFFFFFF when all four are legal.
// ---------------------------------------------------------------------------------------------------------
template <int FT>
__device__ __forceinline__ uint32_t pfd_parse_word(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b0, uint32_t w, uint32_t b2,
                                                   uint32_t c0, uint32_t c1, uint32_t c2, uint32_t& dirw_out, uint32_t& upw_out) {
    // neighbour words per slot: byte j = code of the slot-k neighbour of my cell j
    uint32_t nb[8];
    nb[0] = __byte_perm(a0, a1, 0x6543);  // NW: shift right by one cell
    nb[1] = a1;                           // N
    nb[2] = __byte_perm(a1, a2, 0x4321);  // NE: shift left by one cell
    nb[3] = __byte_perm(b0, w, 0x6543);   // W
    nb[4] = __byte_perm(w, b2, 0x4321);   // E
    nb[5] = __byte_perm(c0, c1, 0x6543);  // SW
    nb[6] = c1;                           // S
    nb[7] = __byte_perm(c1, c2, 0x4321);  // SE

    const uint32_t nodata = __vcmpeq4(w, splat4(pfd_nodata_code<FT>()));
    uint32_t pit, legal;
    if (FT == 0) {
        pit = __vcmpeq4(w, 0u) | __vcmpeq4(w, splat4(255u));
        // legal codes: 0 or a power of two, 247, 255 (core_d8.py:19)
        legal = __vcmpeq4(w & __vsub4(w, splat4(1u)), 0u) | nodata | pit;
    } else {
        pit = __vcmpeq4(w, splat4(5u));
        // legal codes: 1..9 and 255 (core_ldd.py:17)
        legal = __vcmpltu4(__vsub4(w, splat4(1u)), splat4(9u)) | nodata;
    }
    uint32_t dirw = 0, forced = 0, upw = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t sel = __vcmpeq4(w, splat4(pfd_code<FT>(k)));
        dirw |= sel & splat4((uint32_t)k);
        forced |= sel & __vcmpeq4(nb[k], splat4(pfd_nodata_code<FT>()));
        upw |= __vcmpeq4(nb[k], splat4(pfd_code<FT>(7 - k))) & splat4(1u << k);
    }
    dirw = (dirw & ~forced) | (forced & splat4(PFD_DIR_FPIT));
    dirw |= pit & splat4(PFD_DIR_PIT);
    dirw |= nodata;  // 0xFF
    upw &= ~nodata;
    dirw_out = dirw;
    upw_out = upw;
    return legal;
}

// L2-only loads for data that other SMs write inside the same persistent kernel
template <typename T>
__device__ __forceinline__ T ld_cg(const T* p) {
    return __ldcg(p);
}
template <>
__device__ __forceinline__ int8_t ld_cg<int8_t>(const int8_t* p) {
    return (int8_t)__ldcg((const signed char*)p);
}
template <>
__device__ __forceinline__ uint8_t ld_cg<uint8_t>(const uint8_t* p) {
    return (uint8_t)__ldcg((const unsigned char*)p);
}
template <>
__device__ __forceinline__ int16_t ld_cg<int16_t>(const int16_t* p) {
    return (int16_t)__ldcg((const short*)p);
}
template <>
__device__ __forceinline__ uint16_t ld_cg<uint16_t>(const uint16_t* p) {
    return (uint16_t)__ldcg((const unsigned short*)p);
}
template <>
__device__ __forceinline__ int64_t ld_cg<int64_t>(const int64_t* p) {
    return (int64_t)__ldcg((const long long*)p);
}
template <>
__device__ __forceinline__ uint64_t ld_cg<uint64_t>(const uint64_t* p) {
    return (uint64_t)__ldcg((const unsigned long long*)p);
}

// ---------------------------------------------------------------------------------------------------------
// Host side: handle
// ---------------------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

enum { PFD_STAGE_PARSE = 0, PFD_STAGE_PITS = 1, PFD_STAGE_ORDER = 2, PFD_STAGE_SWEEP = 3, PFD_STAGE_TOTAL = 4, PFD_STAGE_BFS = 5,
       PFD_STAGE_TILE_A = 6, PFD_STAGE_TILE_B = 7, PFD_STAGE_TILE_C = 8, PFD_NSTAGE = 9 };

struct pfd_handle {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    std::string err;

    // raster
    int64_t nrow = 0, ncol = 0, n = 0;
    bool parsed = false, ordered = false, have_rank = false, have_basins = false;
    int64_t n_valid = 0, n_pits = 0, n_outlets = 0, nnodes = 0, nlevels = 0;

    // graph + order (device)
    DevBuf dir, upmask;       // uint8 [npad]
    DevBuf pits;              // cell_t [n_pits]
    DevBuf pit_outlet;        // uint8 [n_pits]
    DevBuf seq;               // cell_t [n]
    DevBuf bseq;              // uint32 [n]  basin id aligned with seq positions
    DevBuf rank;              // int32 [n]
    DevBuf basins;            // uint32 [n] default basins (all pits, ids 1..n_pits)
    DevBuf level_off;         // int64 [level_cap + 1]
    int64_t level_cap = 0;
    DevBuf bfs_state;         // BfsState
    DevBuf chunk_status;      // uint64 [n / CHUNK + 2]
    DevBuf blk_counts;        // uint32 pit count per PC_CHUNK
    DevBuf blk_offsets;       // uint64 exclusive scan of blk_counts
    DevBuf counters;          // uint64 [8]: n_valid, n_pits, n_outlets, parse flags, load flags, basins flags
    DevBuf segs;              // SweepSeg schedule of the level replays
    DevBuf tslots;            // reduced-graph (tile ring) arrays of the tile solver
    DevBuf mg_counts;
    DevBuf ts_done, ts_lists;  // tile-dataflow sweeps: done bitmap; work lists + stamps + control block
    int tile_sweeps = 1;       // option "tile_sweeps"
    int ts_max_passes = 0;     // option "sweep_max_passes" (profiling)
    int sweep_passes = 0;      // passes of the last tile-dataflow sweep
    int64_t sweep_visits = 0;  // tile visits of the last tile-dataflow sweep
    // row-block (multi-GPU) tile sweeps: the block extended by one foreign row above and below
    struct SweepShard {
        bool active = false;
        int kind = 0, dtype = 0, vsz = 0, asz = 0, next_pass = 1;
        const void* data = nullptr;      // accuflux data / (unused) of the own rows, device
        const uint8_t* drain = nullptr;  // HAND drain mask of the own rows, device
        double nodata_f = 0.0;
        long long nodata_i = 0;
        int nodata_is_int = 0;
        unsigned long long resolved = 0;  // cells resolved so far (all rounds)
    } sw;
    DevBuf sw_dir, sw_out, sw_aux, sw_fdone, sw_edge[4];  // ext rows; [0,1] = send top / bottom, [2,3] = receive top / bottom
    bool fill_attr_set[2] = {false, false}, hand_attr_set[48] = {};  // dynamic shared memory opt-ins (per device, so per handle)
    DevBuf fill_bufs[12];      // pfd_fill_depressions (pfd_fill.cuh): levels, labels, heap pool ... (kept between calls)
    DevBuf hand_root, hand_sum, hand_slots;  // path-sum HAND (pfd_hand.cuh): per-cell (root, segment sum), ring nodes
    int hand_pathsum = 1;      // option "hand_pathsum": 1 = try the re-associated path sums first (verified, else hop by hop)
    int hand_fin = 0;          // which ring-node buffer holds the final state of the last path-sum attempt
    int hand_engine = 0;       // info "hand_engine": what produced the last pfd_hand result (1 path sums, 2 tile sweep, 3 level replay)
    DevBuf verify;            // VerifyCounts of the pfd_verify_* entry points
    DevBuf btab, bgraph;       // row-tiled multi-GPU solve: boundary tables, boundary graph state
    int64_t dir_off = 0;       // offset of the first OWNED row inside dir (halo row of a row block)
    bool tiled = false;        // parsed as a row block of a larger raster (only the tiled entry points apply)
    int mg_rank = 0, mg_nranks = 0, mg_halo_top = 0, mg_halo_bot = 0;
    uint32_t* mg_basins = nullptr;
    bool mg_fused = false;       // row block parsed inside phase A (fused): idxs_ds is written by pfd_tiled_finish
    void* mg_idxs_user = nullptr; // caller's idxs_ds buffer of the fused row-block path (host or device)
    int mg_idx_dtype = 0;
    int64_t mg_glob_row0 = 0;
    unsigned long long* h_gather = nullptr;  // page-locked: per-rank {n_valid, n_pits, n_outlets, flags} of exchange #1
    int h_gather_ranks = 0;
    void* nccl_comm = nullptr;
    DevBuf sub_idxs;           // cell_t [n_sub] outlets of the last pfd_subbasins_streamorder call
    int64_t n_sub = 0;
    DevBuf sub_labels;         // int64 [n_sub] labels of the last pfd_region_outlets / pfd_region_slices call
    DevBuf sub_slices;         // int4 [n_sub] (row start, row stop, col start, col stop) of the last pfd_region_slices call
    bool have_sub_labels = false, have_sub_slices = false;
    DevBuf stream_off, stream_cells;  // int64 [n_streams + 1] / cell_t [n_stream_cells] of the last pfd_streams call
    int64_t n_streams = -1, n_stream_cells = 0;
    DevBuf tile_loc;           // uint2 [n]: per cell (local terminal | hops << 12, in-tile subtree size)
    DevBuf uparea;            // int32 [n] cached cell-count upstream area (tile solver)
    bool have_uparea = false;
    bool c_attr_set[3] = {false, false, false};  // dynamic shared memory opt-in done for tile_phase_c_kernel<.., IDXMODE>
    bool have_upmask = false;  // false after the fused-parse path: derived from dir on demand (ensure_upmask)
    unsigned long long* h_counters = nullptr;  // page-locked mirror of `counters` (read while the GPU keeps working)
    int fuse_parse = 1;       // option "fuse_parse": pfd_d8_flow_all on device buffers parses inside the tile solver
    int use_tiles = 1;        // option "tiles": 1 = tile-hierarchical solver for rank/basins/uparea, 0 = BFS + sweeps
    int tile_rounds = 0;      // reduced-graph doubling rounds of the last solve
    int nsegs = 0;
    int64_t max_level_size = 0;
    std::vector<long long> h_level_off;

    // staging / scratch
    DevBuf scratch[6];

    // instrumentation
    int64_t launches = 0;
    double stage_ms[PFD_NSTAGE] = {};
    bool stage_used[PFD_NSTAGE] = {};
    cudaEvent_t ev_start[PFD_NSTAGE] = {};
    cudaEvent_t ev_stop[PFD_NSTAGE] = {};
    cudaEvent_t ev_timer[2] = {};
    cudaEvent_t ev_total[2] = {};
    cudaEvent_t ev_copy = nullptr;
    cudaStream_t copy_stream = nullptr;  // D2H copies that overlap compute (pfd_d8_flow_all)
    bool copy_pending = false;
};

static thread_local std::string g_last_error;

static int pfd_fail(pfd_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    g_last_error = msg;
    return code;
}

#define PFD_CUDA(h, call)                                                                          \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            int code__ = (e__ == cudaErrorMemoryAllocation) ? PFD_ERR_OOM : PFD_ERR_CUDA;          \
            return pfd_fail((h), code__,                                                           \
                            std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + \
                                ":" + std::to_string(__LINE__) + ")");                             \
        }                                                                                          \
    } while (0)

#define PFD_TRY(expr)                 \
    do {                              \
        int rc__ = (expr);            \
        if (rc__ != PFD_OK) return rc__; \
    } while (0)

static int pfd_reserve(pfd_handle* h, DevBuf& b, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return PFD_OK;
    if (b.p) {
        PFD_CUDA(h, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    PFD_CUDA(h, cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    return PFD_OK;
}

static void pfd_release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

static bool pfd_is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static inline size_t pfd_dtype_size(int dt) {
    switch (dt) {
    case PFD_I8: case PFD_U8: return 1;
    case PFD_I16: case PFD_U16: return 2;
    case PFD_I32: case PFD_U32: case PFD_F32: return 4;
    case PFD_I64: case PFD_U64: case PFD_F64: return 8;
    default: return 0;
    }
}

// Input staging: returns a device pointer holding `bytes` of `src` (src itself when already on device).
static int pfd_stage_in(pfd_handle* h, const void* src, size_t bytes, int slot, const void** dev) {
    if (pfd_is_device_ptr(src)) {
        *dev = src;
        return PFD_OK;
    }
    PFD_TRY(pfd_reserve(h, h->scratch[slot], bytes));
    PFD_CUDA(h, cudaMemcpyAsync(h->scratch[slot].p, src, bytes, cudaMemcpyHostToDevice, h->stream));
    *dev = h->scratch[slot].p;
    return PFD_OK;
}

// Output staging: device pointer to compute into (dst itself when on device).
static int pfd_stage_out(pfd_handle* h, void* dst, size_t bytes, int slot, void** dev) {
    if (pfd_is_device_ptr(dst)) {
        *dev = dst;
        return PFD_OK;
    }
    PFD_TRY(pfd_reserve(h, h->scratch[slot], bytes));
    *dev = h->scratch[slot].p;
    return PFD_OK;
}

static int pfd_finish_out(pfd_handle* h, void* dst, const void* dev, size_t bytes) {
    if (dev != dst) PFD_CUDA(h, cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, h->stream));
    return PFD_OK;
}

#define PFD_LAUNCH_CHECK(h)                     \
    do {                                        \
        (h)->launches++;                        \
        PFD_CUDA((h), cudaGetLastError());      \
    } while (0)
