// pfd_api.cu -- extern "C" entry points of libpfd_b200.so (see include/pfd_b200.h for the contract).
#include "pfd_common.cuh"
#include "pfd_order.cuh"
#include "pfd_parse.cuh"
#include "pfd_sweeps.cuh"
#include "pfd_compact.cuh"
#include "pfd_local.cuh"
#include "pfd_tiles.cuh"
#include "pfd_tilesweep.cuh"
#include "pfd_verify.cuh"
#include "pfd_synth.h"

#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <new>

#define PFD_VERSION "0.1.0"
#define MAX_CELLS (1ll << 32)

// ---------------------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------------------
struct StageTimer {
    pfd_handle* h;
    int stage;
    StageTimer(pfd_handle* h_, int s) : h(h_), stage(s) { cudaEventRecord(h->ev_start[s], h->stream); h->stage_used[s] = true; }
    ~StageTimer() { cudaEventRecord(h->ev_stop[stage], h->stream); }
};

static void stage_reset(pfd_handle* h) {
    for (int s = 0; s < PFD_NSTAGE; ++s) h->stage_used[s] = false;
}

// call after the stream is synchronised
static void stage_collect(pfd_handle* h) {
    for (int s = 0; s < PFD_NSTAGE; ++s) {
        if (!h->stage_used[s]) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev_start[s], h->ev_stop[s]) == cudaSuccess) h->stage_ms[s] = ms;
        else cudaGetLastError();
    }
}

static int grid_for(int64_t n, int threads, int per_thread = 1, int64_t cap = 1 << 20) {
    int64_t g = (n + (int64_t)threads * per_thread - 1) / ((int64_t)threads * per_thread);
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

static int check_handle(pfd_handle* h) {
    if (!h) return pfd_fail(nullptr, PFD_ERR_INVALID_ARG, "null handle");
    PFD_CUDA(h, cudaSetDevice(h->device));
    return PFD_OK;
}

template <typename IDX>
__global__ void widen_cells_kernel(const cell_t* __restrict__ in, int64_t n, IDX* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (IDX)in[i];
}

// ---------------------------------------------------------------------------------------------------------
// library / handle
// ---------------------------------------------------------------------------------------------------------
extern "C" const char* pfd_version(void) { return PFD_VERSION " (sm_100a)"; }

extern "C" int pfd_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int pfd_device_pci_bus_id(int device, char* out, int capacity) {
    if (!out || capacity < 16) return pfd_fail(nullptr, PFD_ERR_INVALID_ARG, "pfd_device_pci_bus_id: need a buffer of >= 16 bytes");
    if (cudaDeviceGetPCIBusId(out, capacity, device) != cudaSuccess) {
        cudaGetLastError();
        return pfd_fail(nullptr, PFD_ERR_CUDA, "pfd_device_pci_bus_id: no such device");
    }
    return PFD_OK;
}

extern "C" const char* pfd_status_string(int s) {
    switch (s) {
    case PFD_OK: return "ok";
    case PFD_ERR_CUDA: return "CUDA error";
    case PFD_ERR_INVALID_ARG: return "invalid argument";
    case PFD_ERR_INVALID_D8: return "invalid D8 data";
    case PFD_ERR_NO_PITS: return "no pits found";
    case PFD_ERR_STATE: return "invalid call sequence";
    case PFD_ERR_UNSUPPORTED: return "unsupported";
    case PFD_ERR_OOM: return "out of memory";
    case PFD_ERR_NCCL: return "NCCL error";
    default: return "unknown";
    }
}

extern "C" const char* pfd_last_error(const pfd_handle* h) { return h ? h->err.c_str() : g_last_error.c_str(); }

extern "C" int pfd_create(int device, pfd_handle** out) {
    if (!out) return pfd_fail(nullptr, PFD_ERR_INVALID_ARG, "pfd_create: out is null");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return pfd_fail(nullptr, PFD_ERR_CUDA,
                        std::string("pfd_create: no usable CUDA device (") + cudaGetErrorString(e) +
                            "); libpfd_b200 has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return pfd_fail(nullptr, PFD_ERR_INVALID_ARG, "pfd_create: bad device ordinal");
    pfd_handle* h = new (std::nothrow) pfd_handle();
    if (!h) return pfd_fail(nullptr, PFD_ERR_OOM, "pfd_create: host allocation failed");
    h->device = device;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        std::string msg = std::string("pfd_create: ") + cudaGetErrorString(cudaGetLastError());
        delete h;
        return pfd_fail(nullptr, PFD_ERR_CUDA, msg);
    }
    h->num_sms = prop.multiProcessorCount;
    if (!prop.cooperativeLaunch) {
        cudaStreamDestroy(h->stream);
        delete h;
        return pfd_fail(nullptr, PFD_ERR_UNSUPPORTED, "pfd_create: device lacks cooperative launch");
    }
    for (int s = 0; s < PFD_NSTAGE; ++s) {
        cudaEventCreate(&h->ev_start[s]);
        cudaEventCreate(&h->ev_stop[s]);
    }
    cudaEventCreate(&h->ev_timer[0]);
    cudaEventCreate(&h->ev_timer[1]);
    cudaEventCreate(&h->ev_total[0]);
    cudaEventCreate(&h->ev_total[1]);
    cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    *out = h;
    return PFD_OK;
}

extern "C" int pfd_comm_destroy(pfd_handle* h);

extern "C" void pfd_destroy(pfd_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    pfd_comm_destroy(h);
    cudaStreamSynchronize(h->stream);
    DevBuf* bufs[] = {&h->dir, &h->upmask, &h->pits, &h->pit_outlet, &h->seq, &h->bseq, &h->rank, &h->basins,
                      &h->level_off, &h->bfs_state, &h->chunk_status, &h->blk_counts, &h->blk_offsets, &h->counters,
                      &h->segs, &h->tslots, &h->uparea, &h->tile_loc, &h->btab, &h->bgraph, &h->mg_counts, &h->sub_idxs,
                      &h->sub_labels, &h->sub_slices, &h->stream_off, &h->stream_cells, &h->verify, &h->ts_done, &h->ts_lists,
                      &h->hand_root, &h->hand_sum, &h->hand_slots, &h->sw_dir, &h->sw_out, &h->sw_aux, &h->sw_fdone};
    for (DevBuf* b : bufs) pfd_release(*b);
    for (DevBuf& b : h->scratch) pfd_release(b);
    for (DevBuf& b : h->fill_bufs) pfd_release(b);
    for (int s = 0; s < PFD_NSTAGE; ++s) {
        cudaEventDestroy(h->ev_start[s]);
        cudaEventDestroy(h->ev_stop[s]);
    }
    cudaEventDestroy(h->ev_timer[0]);
    cudaEventDestroy(h->ev_timer[1]);
    cudaEventDestroy(h->ev_total[0]);
    cudaEventDestroy(h->ev_total[1]);
    cudaEventDestroy(h->ev_copy);
    if (h->h_counters) cudaFreeHost(h->h_counters);
    if (h->h_gather) cudaFreeHost(h->h_gather);
    cudaStreamSynchronize(h->copy_stream);
    cudaStreamDestroy(h->copy_stream);
    cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int pfd_host_alloc(size_t bytes, void** out) {
    if (!out) return PFD_ERR_INVALID_ARG;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return pfd_fail(nullptr, PFD_ERR_OOM, std::string("pfd_host_alloc: ") + cudaGetErrorString(e));
    }
    return PFD_OK;
}

extern "C" int pfd_host_free(void* p) {
    if (p && cudaFreeHost(p) != cudaSuccess) {
        cudaGetLastError();
        return PFD_ERR_CUDA;
    }
    return PFD_OK;
}

extern "C" int pfd_dev_alloc(pfd_handle* h, size_t bytes, void** out) {
    PFD_TRY(check_handle(h));
    if (!out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_dev_alloc: out is null");
    PFD_CUDA(h, cudaMalloc(out, bytes ? bytes : 16));
    return PFD_OK;
}

extern "C" int pfd_dev_free(pfd_handle* h, void* p) {
    PFD_TRY(check_handle(h));
    if (p) PFD_CUDA(h, cudaFree(p));
    return PFD_OK;
}

extern "C" int pfd_memcpy(pfd_handle* h, void* dst, const void* src, size_t bytes) {
    PFD_TRY(check_handle(h));
    PFD_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    return PFD_OK;
}

extern "C" int pfd_synchronize(pfd_handle* h) {
    PFD_TRY(check_handle(h));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    return PFD_OK;
}

extern "C" int64_t pfd_launch_count(const pfd_handle* h) { return h ? h->launches : 0; }

// position-dependent checksum (verification plumbing, include/pfd_b200.h)
template <typename T>
__global__ void checksum_kernel(const T* __restrict__ v, int64_t n, unsigned long long off, unsigned long long* __restrict__ out) {
    unsigned long long acc = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc += ((unsigned long long)v[i] + 1ull) * ((((unsigned long long)i + off) * 0x9E3779B97F4A7C15ull) | 1ull);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

extern "C" int pfd_checksum(pfd_handle* h, const void* data, int elem_bytes, int64_t count, uint64_t index_offset, uint64_t* out) {
    PFD_TRY(check_handle(h));
    if (!data || !out || count < 0 || (elem_bytes != 1 && elem_bytes != 2 && elem_bytes != 4 && elem_bytes != 8))
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_checksum: bad argument");
    PFD_TRY(pfd_reserve(h, h->counters, 8 * sizeof(unsigned long long)));
    unsigned long long* ctr = (unsigned long long*)h->counters.p + 6;
    PFD_CUDA(h, cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), h->stream));
    const void* dev = nullptr;
    PFD_TRY(pfd_stage_in(h, data, (size_t)count * elem_bytes, 2, &dev));
    if (count > 0) {
        const int g = grid_for(count, 256, 8, 148 * 16);
        if (elem_bytes == 1) checksum_kernel<uint8_t><<<g, 256, 0, h->stream>>>((const uint8_t*)dev, count, index_offset, ctr);
        else if (elem_bytes == 2) checksum_kernel<uint16_t><<<g, 256, 0, h->stream>>>((const uint16_t*)dev, count, index_offset, ctr);
        else if (elem_bytes == 4) checksum_kernel<uint32_t><<<g, 256, 0, h->stream>>>((const uint32_t*)dev, count, index_offset, ctr);
        else checksum_kernel<unsigned long long><<<g, 256, 0, h->stream>>>((const unsigned long long*)dev, count, index_offset, ctr);
        PFD_LAUNCH_CHECK(h);
    }
    unsigned long long v = 0;
    PFD_CUDA(h, cudaMemcpyAsync(&v, ctr, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    *out = v;
    return PFD_OK;
}

// CUDA-event bracket on the handle's stream (bench.py times K steps between start and stop)
extern "C" int pfd_timer_start(pfd_handle* h) {
    PFD_TRY(check_handle(h));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    PFD_CUDA(h, cudaEventRecord(h->ev_timer[0], h->stream));
    return PFD_OK;
}

extern "C" int pfd_timer_stop(pfd_handle* h, double* ms) {
    PFD_TRY(check_handle(h));
    PFD_CUDA(h, cudaEventRecord(h->ev_timer[1], h->stream));
    PFD_CUDA(h, cudaEventSynchronize(h->ev_timer[1]));
    float f = 0.f;
    PFD_CUDA(h, cudaEventElapsedTime(&f, h->ev_timer[0], h->ev_timer[1]));
    if (ms) *ms = (double)f;
    return PFD_OK;
}

extern "C" double pfd_last_stage_ms(const pfd_handle* h, int stage) {
    if (!h || stage < 0 || stage >= PFD_NSTAGE) return 0.0;
    return h->stage_used[stage] ? h->stage_ms[stage] : 0.0;  // 0: the stage did not run in the last call
}

// ---------------------------------------------------------------------------------------------------------
// parse
// ---------------------------------------------------------------------------------------------------------
static void invalidate(pfd_handle* h) {
    h->parsed = h->ordered = h->have_rank = h->have_basins = h->have_uparea = h->have_upmask = false;
    h->n_valid = h->n_pits = h->n_outlets = h->nnodes = h->nlevels = h->n_sub = 0;
    h->n_streams = -1;
}

// pit list in ascending index order + n_valid / n_pits / n_outlets from the dir bytes; fails on the parse flag
static int pits_stage(pfd_handle* h, int64_t npad, int ftype) {
    const int64_t nblk = npad / PC_CHUNK;
    {
        StageTimer t(h, PFD_STAGE_PITS);
        PFD_TRY(pfd_reserve(h, h->blk_counts, (size_t)nblk * sizeof(uint32_t)));
        PFD_TRY(pfd_reserve(h, h->blk_offsets, (size_t)(nblk + 1) * sizeof(unsigned long long)));
        pit_count_kernel<<<(unsigned)nblk, 256, 0, h->stream>>>((const uint8_t*)h->dir.p, (uint32_t*)h->blk_counts.p,
                                                               (unsigned long long*)h->counters.p);
        PFD_LAUNCH_CHECK(h);
        scan_counts_kernel<<<1, 1024, 0, h->stream>>>((const uint32_t*)h->blk_counts.p, nblk,
                                                     (unsigned long long*)h->blk_offsets.p);
        PFD_LAUNCH_CHECK(h);
        unsigned long long hc[4];
        PFD_CUDA(h, cudaMemcpyAsync(hc, h->counters.p, sizeof(hc), cudaMemcpyDeviceToHost, h->stream));
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
        const unsigned int flags = (unsigned int)hc[3];
        if (flags & 1u) {
            invalidate(h);
            return pfd_fail(h, PFD_ERR_INVALID_D8, ftype == 0 ? "raster holds values outside the D8 code set {0,1,2,4,8,16,32,64,128,247,255}"
                                                   : ftype == 1 ? "raster holds values outside the LDD code set {1..9,255}"
                                                                : "NEXTXY data is invalid (core_nextxy.isvalid)");
        }
        h->n_valid = (int64_t)hc[0];
        h->n_pits = (int64_t)hc[1];
        h->n_outlets = (int64_t)hc[2];
        PFD_TRY(pfd_reserve(h, h->pits, (size_t)std::max<int64_t>(h->n_pits, 1) * sizeof(cell_t)));
        PFD_TRY(pfd_reserve(h, h->pit_outlet, (size_t)std::max<int64_t>(h->n_pits, 1)));
        if (h->n_pits > 0) {
            pit_scatter_kernel<<<(unsigned)nblk, 256, 0, h->stream>>>((const uint8_t*)h->dir.p,
                                                                     (const unsigned long long*)h->blk_offsets.p,
                                                                     (cell_t*)h->pits.p, (uint8_t*)h->pit_outlet.p);
            PFD_LAUNCH_CHECK(h);
        }
    }
    return PFD_OK;
}

// d8_dev: device pointer. idxs_dev: device pointer or null.
// d8_dev holds halo_top + nrow + halo_bot rows (halo rows only feed the forced-pit test of the block's edge rows);
// idxs_dev (optional) receives the nrow owned rows as global linear indices (first owned row = glob_row0).
static int parse_device(pfd_handle* h, const uint8_t* d8_dev, int64_t nrow_owned, int64_t ncol, void* idxs_dev, int idx_dtype,
                        int halo_top = 0, int halo_bot = 0, int64_t glob_row0 = 0, int ftype = 0) {
    const int64_t nrow = nrow_owned + halo_top + halo_bot;
    const int64_t n = nrow * ncol;
    const int64_t npad = (n + PC_CHUNK - 1) / PC_CHUNK * PC_CHUNK;
    PFD_TRY(pfd_reserve(h, h->dir, (size_t)npad));
    PFD_TRY(pfd_reserve(h, h->upmask, (size_t)npad));
    PFD_TRY(pfd_reserve(h, h->counters, 8 * sizeof(unsigned long long)));
    PFD_CUDA(h, cudaMemsetAsync(h->counters.p, 0, 8 * sizeof(unsigned long long), h->stream));
    if (npad > n) PFD_CUDA(h, cudaMemsetAsync((uint8_t*)h->dir.p + n, 0xFF, (size_t)(npad - n), h->stream));
    unsigned int* flag = reinterpret_cast<unsigned int*>((unsigned long long*)h->counters.p + 3);
    {
        StageTimer t(h, PFD_STAGE_PARSE);
        dim3 grid((unsigned)((ncol + PT_W - 1) / PT_W), (unsigned)((nrow + PT_H - 1) / PT_H));
        if (grid.y > 65535u) return pfd_fail(h, PFD_ERR_UNSUPPORTED, "pfd_d8_parse: more than 2097120 rows");
        const int idxmode = !idxs_dev ? 0 : (idx_dtype == PFD_I64 ? 2 : 1);
        const bool aligned = (ncol % 4 == 0) && ((uintptr_t)d8_dev % 4 == 0) && (!idxs_dev || (uintptr_t)idxs_dev % 16 == 0);
        uint8_t* dir = (uint8_t*)h->dir.p;
        uint8_t* upm = (uint8_t*)h->upmask.p;
#define LAUNCH_PARSE(A, M, F) \
    parse_kernel<A, M, F><<<grid, 256, 0, h->stream>>>(d8_dev, nrow, ncol, dir, upm, idxs_dev, flag, (int64_t)halo_top, nrow_owned, glob_row0)
#define LAUNCH_PARSE_FT(F)                                \
    do {                                                  \
        if (aligned) {                                    \
            if (idxmode == 0) LAUNCH_PARSE(true, 0, F);   \
            else if (idxmode == 1) LAUNCH_PARSE(true, 1, F); \
            else LAUNCH_PARSE(true, 2, F);                \
        } else {                                          \
            if (idxmode == 0) LAUNCH_PARSE(false, 0, F);  \
            else if (idxmode == 1) LAUNCH_PARSE(false, 1, F); \
            else LAUNCH_PARSE(false, 2, F);               \
        }                                                 \
    } while (0)
        if (ftype == 0) LAUNCH_PARSE_FT(0);
        else LAUNCH_PARSE_FT(1);
#undef LAUNCH_PARSE_FT
#undef LAUNCH_PARSE
        PFD_LAUNCH_CHECK(h);
        // halo rows must not contribute pits / valid cells: blank them after the owned rows were parsed
        if (halo_top) PFD_CUDA(h, cudaMemsetAsync(h->dir.p, 0xFF, (size_t)ncol, h->stream));
        if (halo_bot) PFD_CUDA(h, cudaMemsetAsync((uint8_t*)h->dir.p + (n - ncol), 0xFF, (size_t)ncol, h->stream));
    }
    PFD_TRY(pits_stage(h, npad, ftype));
    h->nrow = nrow_owned;
    h->ncol = ncol;
    h->n = nrow_owned * ncol;
    h->dir_off = (int64_t)halo_top * ncol;
    h->tiled = (halo_top || halo_bot || glob_row0 != 0);
    h->parsed = true;
    h->have_upmask = true;
    h->ordered = h->have_rank = h->have_basins = h->have_uparea = false;
    return PFD_OK;
}

static int check_shape(pfd_handle* h, int64_t nrow, int64_t ncol, const char* who) {
    if (nrow <= 0 || ncol <= 0) return pfd_fail(h, PFD_ERR_INVALID_ARG, std::string(who) + ": empty raster");
    if (nrow > MAX_CELLS / ncol) return pfd_fail(h, PFD_ERR_UNSUPPORTED, std::string(who) + ": more than 2^32 cells");
    return PFD_OK;
}

// overlap_idxs_copy: the D2H copy of idxs_ds runs on the handle's copy stream (the caller joins it later)
static int parse_impl(pfd_handle* h, const uint8_t* d8, int64_t nrow, int64_t ncol, void* idxs_ds_out, int idx_dtype,
                      bool overlap_idxs_copy = false, int ftype = 0) {
    PFD_TRY(check_shape(h, nrow, ncol, "pfd_d8_parse"));
    if (!d8) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_d8_parse: d8 is null");
    if (idxs_ds_out && idx_dtype != PFD_I32 && idx_dtype != PFD_U32 && idx_dtype != PFD_I64)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_d8_parse: idx_dtype must be int32, uint32 or int64");
    const int64_t n = nrow * ncol;
    if (idxs_ds_out && idx_dtype == PFD_I32 && n >= 2147483647ll)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_d8_parse: int32 indices cannot address this raster");
    invalidate(h);
    const void* d8_dev = nullptr;
    PFD_TRY(pfd_stage_in(h, d8, (size_t)n, 0, &d8_dev));
    void* idxs_dev = nullptr;
    const size_t ibytes = (size_t)n * pfd_dtype_size(idx_dtype);
    if (idxs_ds_out) PFD_TRY(pfd_stage_out(h, idxs_ds_out, ibytes, 1, &idxs_dev));
    PFD_TRY(parse_device(h, (const uint8_t*)d8_dev, nrow, ncol, idxs_dev, idx_dtype, 0, 0, 0, ftype));
    if (idxs_ds_out && idxs_dev != idxs_ds_out) {
        if (overlap_idxs_copy) {
            // parse_device ended with a stream synchronisation (pit counters), so the staging buffer is complete
            PFD_CUDA(h, cudaMemcpyAsync(idxs_ds_out, idxs_dev, ibytes, cudaMemcpyDeviceToHost, h->copy_stream));
            h->copy_pending = true;
        } else {
            PFD_TRY(pfd_finish_out(h, idxs_ds_out, idxs_dev, ibytes));
        }
    }
    return PFD_OK;
}

extern "C" int pfd_d8_parse(pfd_handle* h, const uint8_t* d8, int64_t nrow, int64_t ncol, int check_values,
                            void* idxs_ds_out, int idx_dtype, int64_t* n_valid, int64_t* n_pits, int64_t* n_outlets) {
    (void)check_values;  // illegal codes are always refused
    PFD_TRY(check_handle(h));
    stage_reset(h);
    PFD_TRY(parse_impl(h, d8, nrow, ncol, idxs_ds_out, idx_dtype));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    if (n_valid) *n_valid = h->n_valid;
    if (n_pits) *n_pits = h->n_pits;
    if (n_outlets) *n_outlets = h->n_outlets;
    return PFD_OK;
}

// core_ldd.from_array (pyflwdir/core_ldd.py:41-66): same kernel, PCRaster LDD code table
extern "C" int pfd_ldd_parse(pfd_handle* h, const uint8_t* ldd, int64_t nrow, int64_t ncol, void* idxs_ds_out, int idx_dtype,
                             int64_t* n_valid, int64_t* n_pits) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    PFD_TRY(parse_impl(h, ldd, nrow, ncol, idxs_ds_out, idx_dtype, false, 1));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    if (n_valid) *n_valid = h->n_valid;
    if (n_pits) *n_pits = h->n_pits;
    return PFD_OK;
}

// core_nextxy.from_array (pyflwdir/core_nextxy.py:24-67): CaMa-Flood NEXTXY rasters. The device graph stores one byte per
// cell (slot of the downstream neighbour), so every link has to stay inside the 8-neighbourhood; a raster with longer
// links (e.g. across the date line) is refused with PFD_ERR_UNSUPPORTED.
extern "C" int pfd_nextxy_parse(pfd_handle* h, const int32_t* nextx, const int32_t* nexty, int64_t nrow, int64_t ncol, int check_values,
                                void* idxs_ds_out, int idx_dtype, int64_t* n_valid, int64_t* n_pits, int64_t* n_outlets) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    PFD_TRY(check_shape(h, nrow, ncol, "pfd_nextxy_parse"));
    if (!nextx || !nexty) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_nextxy_parse: null array");
    if (idxs_ds_out && idx_dtype != PFD_I32 && idx_dtype != PFD_U32 && idx_dtype != PFD_I64)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_nextxy_parse: idx_dtype must be int32, uint32 or int64");
    const int64_t n = nrow * ncol;
    if (idxs_ds_out && idx_dtype == PFD_I32 && n >= 2147483647ll)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_nextxy_parse: int32 indices cannot address this raster");
    invalidate(h);
    const void *x_dev = nullptr, *y_dev = nullptr;
    PFD_TRY(pfd_stage_in(h, nextx, (size_t)n * 4, 0, &x_dev));
    PFD_TRY(pfd_stage_in(h, nexty, (size_t)n * 4, 2, &y_dev));
    void* idxs_dev = nullptr;
    const size_t ibytes = (size_t)n * pfd_dtype_size(idx_dtype);
    if (idxs_ds_out) PFD_TRY(pfd_stage_out(h, idxs_ds_out, ibytes, 1, &idxs_dev));
    const int64_t npad = (n + PC_CHUNK - 1) / PC_CHUNK * PC_CHUNK;
    PFD_TRY(pfd_reserve(h, h->dir, (size_t)npad));
    PFD_TRY(pfd_reserve(h, h->upmask, (size_t)npad));
    PFD_TRY(pfd_reserve(h, h->counters, 8 * sizeof(unsigned long long)));
    PFD_CUDA(h, cudaMemsetAsync(h->counters.p, 0, 8 * sizeof(unsigned long long), h->stream));
    if (npad > n) PFD_CUDA(h, cudaMemsetAsync((uint8_t*)h->dir.p + n, 0xFF, (size_t)(npad - n), h->stream));
    unsigned int* flag = reinterpret_cast<unsigned int*>((unsigned long long*)h->counters.p + 3);
    unsigned int* flag2 = reinterpret_cast<unsigned int*>((unsigned long long*)h->counters.p + 4);
    {
        StageTimer t(h, PFD_STAGE_PARSE);
        const int g = grid_for(n, 256, 2);
        const int idxmode = !idxs_dev ? 0 : (idx_dtype == PFD_I64 ? 2 : 1);
#define LAUNCH_NX(M)                                                                                                    \
    nextxy_parse_kernel<M><<<g, 256, 0, h->stream>>>((const int32_t*)x_dev, (const int32_t*)y_dev, nrow, ncol, check_values, \
                                                     (uint8_t*)h->dir.p, idxs_dev, flag, flag2)
        if (idxmode == 0) LAUNCH_NX(0);
        else if (idxmode == 1) LAUNCH_NX(1);
        else LAUNCH_NX(2);
#undef LAUNCH_NX
        PFD_LAUNCH_CHECK(h);
    }
    PFD_TRY(pits_stage(h, npad, 2));
    unsigned int hflag2 = 0;
    PFD_CUDA(h, cudaMemcpyAsync(&hflag2, flag2, sizeof(hflag2), cudaMemcpyDeviceToHost, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    if (hflag2 & 2u) {
        invalidate(h);
        return pfd_fail(h, PFD_ERR_UNSUPPORTED, "pfd_nextxy_parse: a downstream link leaves the 8 neighbours of its cell (or points at the cell itself)");
    }
    h->nrow = nrow;
    h->ncol = ncol;
    h->n = n;
    h->dir_off = 0;
    h->tiled = false;
    h->parsed = true;
    h->have_upmask = false;  // derived from dir on demand (ensure_upmask)
    h->ordered = h->have_rank = h->have_basins = h->have_uparea = false;
    if (idxs_ds_out) PFD_TRY(pfd_finish_out(h, idxs_ds_out, idxs_dev, ibytes));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    if (n_valid) *n_valid = h->n_valid;
    if (n_pits) *n_pits = h->n_pits;
    if (n_outlets) *n_outlets = h->n_outlets;
    return PFD_OK;
}

extern "C" int pfd_load_idxs_ds(pfd_handle* h, const void* idxs_ds, int idx_dtype, int64_t nrow, int64_t ncol,
                                int64_t* n_valid, int64_t* n_pits) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    PFD_TRY(check_shape(h, nrow, ncol, "pfd_load_idxs_ds"));
    if (!idxs_ds) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_load_idxs_ds: idxs_ds is null");
    const size_t isz = pfd_dtype_size(idx_dtype);
    if (idx_dtype != PFD_I32 && idx_dtype != PFD_U32 && idx_dtype != PFD_I64 && idx_dtype != PFD_U64)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_load_idxs_ds: idx_dtype must be a 32/64-bit integer");
    invalidate(h);
    const int64_t n = nrow * ncol;
    const void* idev = nullptr;
    PFD_TRY(pfd_stage_in(h, idxs_ds, (size_t)n * isz, 1, &idev));
    PFD_TRY(pfd_reserve(h, h->scratch[0], (size_t)n));
    PFD_TRY(pfd_reserve(h, h->counters, 8 * sizeof(unsigned long long)));
    PFD_CUDA(h, cudaMemsetAsync(h->counters.p, 0, 8 * sizeof(unsigned long long), h->stream));
    unsigned int* flag = reinterpret_cast<unsigned int*>((unsigned long long*)h->counters.p + 4);
    uint8_t* d8 = (uint8_t*)h->scratch[0].p;
    const int g = grid_for(n, 256, 4);
    switch (idx_dtype) {
    case PFD_I32: idxs_to_d8_kernel<int32_t><<<g, 256, 0, h->stream>>>((const int32_t*)idev, n, ncol, d8, flag); break;
    case PFD_U32: idxs_to_d8_kernel<uint32_t><<<g, 256, 0, h->stream>>>((const uint32_t*)idev, n, ncol, d8, flag); break;
    case PFD_I64: idxs_to_d8_kernel<int64_t><<<g, 256, 0, h->stream>>>((const int64_t*)idev, n, ncol, d8, flag); break;
    default: idxs_to_d8_kernel<uint64_t><<<g, 256, 0, h->stream>>>((const uint64_t*)idev, n, ncol, d8, flag); break;
    }
    PFD_LAUNCH_CHECK(h);
    unsigned int hflag = 0;
    PFD_CUDA(h, cudaMemcpyAsync(&hflag, flag, sizeof(hflag), cudaMemcpyDeviceToHost, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    if (hflag & 2u)
        return pfd_fail(h, PFD_ERR_UNSUPPORTED, "Invalid data downstream index outside 8 neighbors.");
    PFD_TRY(parse_device(h, d8, nrow, ncol, nullptr, 0));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    if (n_valid) *n_valid = h->n_valid;
    if (n_pits) *n_pits = h->n_pits;
    return PFD_OK;
}


// ---------------------------------------------------------------------------------------------------------
// tile-hierarchical solver (pfd_tiles.cuh): rank / basins() / upstream_area("cell") without ordering the cells
// ---------------------------------------------------------------------------------------------------------
__global__ void count_ranked_kernel(const int32_t* __restrict__ rank, int64_t n, unsigned long long* __restrict__ out) {
    unsigned long long c = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        c += (rank[i] >= 0) ? 1u : 0u;
    c = __reduce_add_sync(0xFFFFFFFFu, (unsigned)c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

static bool tiles_usable(const pfd_handle* h) { return h->use_tiles && h->n_pits > 0 && h->n_pits < (1ll << 31); }

template <class K>
static int coop_grid(pfd_handle* h, K kernel, int threads, int64_t max_useful_blocks, int* grid);

struct TileCtx {
    long long ntx = 0, nty = 0, nslots = 0;
    size_t arr = 0;
    SlotBuf B[2];       // double-buffered reduced-graph state (B[1].acc == B[0].acc)
    uint32_t* recv[2] = {nullptr, nullptr};
    SlotBuf init;       // copy of the state after phase A (multi-rank: the local reduced graph is solved twice)
    uint32_t *term = nullptr, *term_h = nullptr, *sbasin = nullptr;
    int32_t* srank = nullptr;
    unsigned int* flag = nullptr;
    const uint8_t* dir = nullptr;  // first OWNED row
};

static int tiles_setup(pfd_handle* h, TileCtx& T, bool with_init_copy) {
    if (!h->parsed) return pfd_fail(h, PFD_ERR_STATE, "no raster parsed on this handle");
    T.ntx = (h->ncol + TL_W - 1) / TL_W;
    T.nty = (h->nrow + TL_H - 1) / TL_H;
    if (T.nty > 65535) return pfd_fail(h, PFD_ERR_UNSUPPORTED, "tile solver: more than 4194240 rows");
    T.nslots = T.ntx * (T.nty + 2) * TL_RING;  // + one halo tile row above and below
    if (T.nslots >= (1ll << 30)) return pfd_fail(h, PFD_ERR_UNSUPPORTED, "tile solver: too many ring slots");
    T.arr = (size_t)T.nslots * sizeof(uint32_t);
    const int narr = with_init_copy ? 17 : 13;
    PFD_TRY(pfd_reserve(h, h->tslots, narr * T.arr + 256));
    uint32_t* base = (uint32_t*)h->tslots.p;
    int a = 0;
    auto take = [&]() { return base + (size_t)T.nslots * a++; };
    for (int side = 0; side < 2; ++side) {
        T.B[side].nxt = take();
        T.B[side].rh = take();
        T.B[side].ch = take();
        T.recv[side] = take();
    }
    T.B[0].acc = T.B[1].acc = take();
    T.term = take();
    T.term_h = take();
    T.srank = (int32_t*)take();
    T.sbasin = take();
    if (with_init_copy) {
        T.init.nxt = take();
        T.init.rh = take();
        T.init.ch = take();
        T.init.acc = take();
    }
    T.flag = (unsigned int*)(base + (size_t)T.nslots * a);
    T.dir = (const uint8_t*)h->dir.p + h->dir_off;
    PFD_TRY(pfd_reserve(h, h->tile_loc, (size_t)h->n * sizeof(uint2)));  // (loc, cnt) per cell
    return PFD_OK;
}

// phase A: per-tile local solve -> loc / cnt per cell, ring nodes + inflow weights in T.B[0]
static int tiles_phase_a(pfd_handle* h, TileCtx& T, uint32_t* basin_dev, unsigned long long pit_id_offset) {
    StageTimer t(h, PFD_STAGE_TILE_A);
    const dim3 grid((unsigned)T.ntx, (unsigned)T.nty);
    PFD_CUDA(h, cudaMemsetAsync(T.B[0].acc, 0, T.arr, h->stream));
    halo_slots_init_kernel<<<grid_for(2 * T.ntx * TL_RING, 256), 256, 0, h->stream>>>(T.B[0], T.term, T.term_h, T.ntx, T.nty);
    PFD_LAUNCH_CHECK(h);
    if (basin_dev && h->n_pits > 0) {
        stash_pit_ids_kernel<<<grid_for(h->n_pits, 256, 1, 148 * 16), 256, 0, h->stream>>>(
            (const cell_t*)h->pits.p, h->n_pits, h->dir_off, pit_id_offset, basin_dev);
        PFD_LAUNCH_CHECK(h);
    }
    PhaseAArgs A{};
    A.dir = T.dir, A.nrow = h->nrow, A.ncol = h->ncol, A.ntx = T.ntx, A.pit_ids = basin_dev;
    A.loccnt = (uint2*)h->tile_loc.p, A.W = T.B[0].acc, A.s_nxt = T.B[0].nxt, A.s_rh = T.B[0].rh, A.s_ch = T.B[0].ch;
    A.s_term = T.term, A.s_term_h = T.term_h;
    A.al4 = (h->ncol % 4 == 0) && ((uintptr_t)T.dir % 4 == 0);  // tile_loc is cudaMalloc-aligned
    tl_launch_phase_a<false>(grid, h->stream, A);
    PFD_LAUNCH_CHECK(h);
    return PFD_OK;
}

// phase B: all doubling rounds over `n` reduced-graph nodes in one cooperative launch. The final node state is on
// side (*rounds & 1) where rounds = (int*)(flags + 3) lives on the device; acc is side-independent.
static int slots_solve(pfd_handle* h, SlotBuf* B, uint32_t** recv, long long n, unsigned int* flags, int mark_inert,
                       SlotProtect prot) {
    int grid = 1;
    PFD_TRY(coop_grid(h, slots_solve_kernel, 256, (n + 255) / 256, &grid));
    PFD_CUDA(h, cudaMemsetAsync(flags, 0, 4 * sizeof(unsigned int), h->stream));
    PFD_CUDA(h, cudaMemsetAsync(recv[0], 0, (size_t)n * sizeof(uint32_t), h->stream));
    PFD_CUDA(h, cudaMemsetAsync(recv[1], 0, (size_t)n * sizeof(uint32_t), h->stream));
    int* rounds_dev = (int*)(flags + 3);
    void* args[] = {(void*)&B[0], (void*)&B[1], (void*)&recv[0], (void*)&recv[1], (void*)&n,
                    (void*)&flags,  (void*)&rounds_dev, (void*)&mark_inert, (void*)&prot};
    PFD_CUDA(h, cudaLaunchCooperativeKernel((void*)slots_solve_kernel, dim3(grid), dim3(256), args, 0, h->stream));
    h->launches++;
    return PFD_OK;
}

static SlotProtect tiles_protect(const pfd_handle* h, const TileCtx& T, bool first_multirank_pass) {
    SlotProtect p;
    p.per_row = T.ntx * TL_RING;
    p.nty = T.nty;
    p.top = first_multirank_pass ? h->mg_halo_top : 0;
    p.bot = first_multirank_pass ? h->mg_halo_bot : 0;
    return p;
}

// pit_stash != null (fused-parse paths): pit terminals carry the pit's local cell index; the id sits in the basin buffer
static int tiles_phase_b(pfd_handle* h, TileCtx& T, const uint32_t* pit_stash = nullptr) {
    StageTimer t(h, PFD_STAGE_TILE_B);
    PFD_TRY(slots_solve(h, T.B, T.recv, T.nslots, T.flag, 1, tiles_protect(h, T, false)));
    slots_finalize_kernel<<<grid_for(T.nslots, 256, 2, 148 * 16), 256, 0, h->stream>>>(
        T.B[0], T.B[1], (const int*)(T.flag + 3), T.term, T.term_h, T.nslots, T.srank, T.sbasin, pit_stash, h->ncol, T.ntx, T.nty);
    PFD_LAUNCH_CHECK(h);
    return PFD_OK;
}

// idxs_dev (optional, fused-parse path): idxs_ds in `idx_dtype` is written by the same kernel
static int tiles_phase_c(pfd_handle* h, TileCtx& T, int32_t* rank_dev, uint32_t* basin_dev, int32_t* uparea_dev,
                         void* idxs_dev = nullptr, int idx_dtype = PFD_I32, long long idx_base = 0) {
    StageTimer t(h, PFD_STAGE_TILE_C);
    const dim3 grid((unsigned)T.ntx, (unsigned)T.nty);
    auto al16 = [](const void* p) { return !p || (uintptr_t)p % 16 == 0; };
    PhaseCArgs A{};
    A.dir = T.dir, A.nrow = h->nrow, A.ncol = h->ncol, A.ntx = T.ntx, A.nty = T.nty;
    A.loccnt = (const uint2*)h->tile_loc.p, A.inflow = T.B[0].acc, A.s_rank = T.srank, A.s_basin = T.sbasin;
    A.rank_out = rank_dev, A.basin_out = basin_dev, A.uparea_out = uparea_dev, A.idxs_out = idxs_dev;
    A.idx_base = idx_base;
    A.al4 = (h->ncol % 4 == 0) && ((uintptr_t)T.dir % 4 == 0) && al16(rank_dev) && al16(basin_dev) && al16(uparea_dev) &&
            al16(idxs_dev);
    const size_t smem = sizeof(TileSharedC);
#define LAUNCH_C(M)                                                                                                   \
    do {                                                                                                              \
        if (!h->c_attr_set[M]) { /* the attribute is per device (context) and per instantiation */                   \
            PFD_CUDA(h, cudaFuncSetAttribute(tile_phase_c_kernel<TLC_THREADS, TLC_MINBLOCKS, M>,                      \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
            h->c_attr_set[M] = true;                                                                                  \
        }                                                                                                             \
        tile_phase_c_kernel<TLC_THREADS, TLC_MINBLOCKS, M><<<grid, TLC_THREADS, smem, h->stream>>>(A);                 \
    } while (0)
    if (!idxs_dev) LAUNCH_C(0);
    else if (idx_dtype == PFD_I64) LAUNCH_C(2);
    else LAUNCH_C(1);
#undef LAUNCH_C
    PFD_LAUNCH_CHECK(h);
    return PFD_OK;
}

// Single-GPU solve. Any of the three device outputs may be null.
static int tiles_solve(pfd_handle* h, int32_t* rank_dev, uint32_t* basin_dev, int32_t* uparea_dev) {
    TileCtx T;
    PFD_TRY(tiles_setup(h, T, false));
    PFD_TRY(tiles_phase_a(h, T, basin_dev, 0));
    PFD_TRY(tiles_phase_b(h, T));
    PFD_TRY(tiles_phase_c(h, T, rank_dev, basin_dev, uparea_dev));
    return PFD_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Fused headline path (pfd_d8_flow_all with device-resident buffers): the raster is read ONCE as raw D8 codes by
// phase A, which derives + writes `dir` itself; idxs_ds is written by phase C next to rank / basins / uparea. The
// separate parse pass (7 B/cell) disappears. Pits are numbered (count -> scan -> scatter -> stash) from the `dir`
// phase A wrote; the host needs the pit count to size the pit list, and waits for it WHILE the reduced-graph solve
// (which does not depend on the pit numbering) runs on the GPU. `upmask` is not produced here: ensure_upmask()
// derives it from `dir` the first time an ordering / sweep entry point asks for it.
// ---------------------------------------------------------------------------------------------------------
__global__ void upmask_from_dir_kernel(const uint8_t* __restrict__ dir, long long nrow, long long ncol, uint8_t* __restrict__ upmask) {
    const long long n = nrow * ncol;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / ncol, c = i % ncol;
        uint32_t m = 0;
        if (dir[i] != PFD_DIR_NODATA) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const long long rr = r + pfd_slot_dr(k), cc = c + pfd_slot_dc(k);
                if (rr >= 0 && rr < nrow && cc >= 0 && cc < ncol && __ldg(dir + rr * ncol + cc) == (uint8_t)(7 - k)) m |= 1u << k;
            }
        }
        upmask[i] = (uint8_t)m;
    }
}

static int ensure_upmask(pfd_handle* h) {
    if (!h->parsed) return pfd_fail(h, PFD_ERR_STATE, "no raster parsed on this handle");
    if (h->have_upmask) return PFD_OK;
    const int64_t npad = (h->n + PC_CHUNK - 1) / PC_CHUNK * PC_CHUNK;
    PFD_TRY(pfd_reserve(h, h->upmask, (size_t)npad));
    upmask_from_dir_kernel<<<grid_for(h->n, 256, 4), 256, 0, h->stream>>>((const uint8_t*)h->dir.p, h->nrow, h->ncol,
                                                                          (uint8_t*)h->upmask.p);
    PFD_LAUNCH_CHECK(h);
    h->have_upmask = true;
    return PFD_OK;
}

static int flow_all_fused(pfd_handle* h, const uint8_t* d8_dev, int64_t nrow, int64_t ncol, void* idxs_dev, int idx_dtype,
                          int32_t* rank_dev, uint32_t* basin_dev, int32_t* uparea_dev) {
    const int64_t n = nrow * ncol;
    const int64_t npad = (n + PC_CHUNK - 1) / PC_CHUNK * PC_CHUNK;
    PFD_TRY(pfd_reserve(h, h->dir, (size_t)npad));
    PFD_TRY(pfd_reserve(h, h->counters, 8 * sizeof(unsigned long long)));
    if (!h->h_counters) PFD_CUDA(h, cudaHostAlloc((void**)&h->h_counters, 8 * sizeof(unsigned long long), cudaHostAllocDefault));
    PFD_CUDA(h, cudaMemsetAsync(h->counters.p, 0, 8 * sizeof(unsigned long long), h->stream));
    if (npad > n) PFD_CUDA(h, cudaMemsetAsync((uint8_t*)h->dir.p + n, 0xFF, (size_t)(npad - n), h->stream));
    unsigned int* flag = reinterpret_cast<unsigned int*>((unsigned long long*)h->counters.p + 3);
    h->nrow = nrow;
    h->ncol = ncol;
    h->n = n;
    h->dir_off = 0;
    h->tiled = false;
    h->parsed = true;  // tiles_setup checks it; invalidate() on any failure below
    h->have_upmask = false;
    TileCtx T;
    int rc = tiles_setup(h, T, false);
    if (rc != PFD_OK) {
        invalidate(h);
        return rc;
    }
    const dim3 grid((unsigned)T.ntx, (unsigned)T.nty);
    {
        StageTimer t(h, PFD_STAGE_TILE_A);
        PFD_CUDA(h, cudaMemsetAsync(T.B[0].acc, 0, T.arr, h->stream));
        halo_slots_init_kernel<<<grid_for(2 * T.ntx * TL_RING, 256), 256, 0, h->stream>>>(T.B[0], T.term, T.term_h, T.ntx, T.nty);
        PFD_LAUNCH_CHECK(h);
        PhaseAArgs A{};
        A.nrow = nrow, A.ncol = ncol, A.ntx = T.ntx;
        A.loccnt = (uint2*)h->tile_loc.p, A.W = T.B[0].acc, A.s_nxt = T.B[0].nxt, A.s_rh = T.B[0].rh, A.s_ch = T.B[0].ch;
        A.s_term = T.term, A.s_term_h = T.term_h;
        A.d8 = d8_dev, A.dir_out = (uint8_t*)h->dir.p, A.invalid_flag = flag;
        A.al4 = (ncol % 4 == 0) && ((uintptr_t)d8_dev % 4 == 0);
        tl_launch_phase_a<true>(grid, h->stream, A);
        PFD_LAUNCH_CHECK(h);
    }
    const int64_t nblk = npad / PC_CHUNK;
    {
        StageTimer t(h, PFD_STAGE_PITS);
        PFD_TRY(pfd_reserve(h, h->blk_counts, (size_t)nblk * sizeof(uint32_t)));
        PFD_TRY(pfd_reserve(h, h->blk_offsets, (size_t)(nblk + 1) * sizeof(unsigned long long)));
        pit_count_kernel<<<(unsigned)nblk, 256, 0, h->stream>>>((const uint8_t*)h->dir.p, (uint32_t*)h->blk_counts.p,
                                                               (unsigned long long*)h->counters.p);
        PFD_LAUNCH_CHECK(h);
        scan_counts_kernel<<<1, 1024, 0, h->stream>>>((const uint32_t*)h->blk_counts.p, nblk,
                                                     (unsigned long long*)h->blk_offsets.p);
        PFD_LAUNCH_CHECK(h);
        PFD_CUDA(h, cudaMemcpyAsync(h->h_counters, h->counters.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                    h->stream));
        PFD_CUDA(h, cudaEventRecord(h->ev_copy, h->stream));
    }
    {
        // stage B = reduced-graph solve, pit list + id stash (queued once the host knows the pit count), finalize
        StageTimer t(h, PFD_STAGE_TILE_B);
        PFD_TRY(slots_solve(h, T.B, T.recv, T.nslots, T.flag, 1, tiles_protect(h, T, false)));
        PFD_CUDA(h, cudaEventSynchronize(h->ev_copy));  // the solve keeps the GPU busy meanwhile
        const unsigned int flags = (unsigned int)h->h_counters[3];
        h->n_valid = (int64_t)h->h_counters[0];
        h->n_pits = (int64_t)h->h_counters[1];
        h->n_outlets = (int64_t)h->h_counters[2];
        if ((flags & 1u) || h->n_pits == 0 || h->n_pits >= (1ll << 31)) {
            const int64_t np = h->n_pits;
            cudaStreamSynchronize(h->stream);
            invalidate(h);
            if (flags & 1u)
                return pfd_fail(h, PFD_ERR_INVALID_D8, "raster holds values outside the D8 code set {0,1,2,4,8,16,32,64,128,247,255}");
            if (np == 0) return pfd_fail(h, PFD_ERR_NO_PITS, "Invalid FlwdirRaster: no pits found");
            return pfd_fail(h, PFD_ERR_UNSUPPORTED, "tile solver: more than 2^31 pits");
        }
        PFD_TRY(pfd_reserve(h, h->pits, (size_t)h->n_pits * sizeof(cell_t)));
        PFD_TRY(pfd_reserve(h, h->pit_outlet, (size_t)h->n_pits));
        pit_scatter_kernel<<<(unsigned)nblk, 256, 0, h->stream>>>((const uint8_t*)h->dir.p, (const unsigned long long*)h->blk_offsets.p,
                                                                 (cell_t*)h->pits.p, (uint8_t*)h->pit_outlet.p);
        PFD_LAUNCH_CHECK(h);
        if (basin_dev) {
            stash_pit_ids_kernel<<<grid_for(h->n_pits, 256, 1, 148 * 16), 256, 0, h->stream>>>((const cell_t*)h->pits.p, h->n_pits, 0,
                                                                                              0ull, basin_dev);
            PFD_LAUNCH_CHECK(h);
        }
        slots_finalize_kernel<<<grid_for(T.nslots, 256, 2, 148 * 16), 256, 0, h->stream>>>(
            T.B[0], T.B[1], (const int*)(T.flag + 3), T.term, T.term_h, T.nslots, T.srank, T.sbasin, basin_dev, ncol, T.ntx);
        PFD_LAUNCH_CHECK(h);
    }
    PFD_TRY(tiles_phase_c(h, T, rank_dev, basin_dev, uparea_dev, idxs_dev, idx_dtype));
    return PFD_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Row-tiled multi-GPU solve (BASELINE config 4): every rank owns a block of rows (multiple of 64 except the last),
// parses it with one halo row of D8 codes from each neighbour, solves its tiles and its local reduced graph, and
// meets the other ranks in exactly two small exchanges: the pit counts (for global basin ids) and ONE all-reduce of
// the boundary tables (4 x 2(R-1) x ncol uint32). The boundary graph (entries on the block edges) is then solved
// redundantly by every rank with the same doubling kernels. Integer payloads only => bit-exact.
// The steps are exposed separately so the exchange can be NCCL (pfd_comm_*) or an in-process emulation (tests).
// ---------------------------------------------------------------------------------------------------------
static long long boundary_entries(const pfd_handle* h) { return 2ll * (h->mg_nranks - 1) * h->ncol; }

static void boundary_tables(pfd_handle* h, BoundaryTables& bt) {
    const long long nb = boundary_entries(h);
    uint32_t* p = (uint32_t*)h->btab.p;
    bt.h1 = p;
    bt.nxt = p + nb;
    bt.hop = p + 2 * nb;
    bt.bas = p + 3 * nb;
}

// Fused row-block path: phase A parses the block itself (the neighbours' edge rows of D8 codes are read for the
// forced-pit test only), so the separate parse pass over the block disappears, as on a single GPU. Queues: phase A, the pit
// count + scan, and the copy of {n_valid, n_pits, n_outlets, flags} to the page-locked mirror (event ev_copy).
static int tiled_fused_begin(pfd_handle* h, const uint8_t* d8_owned, int64_t nrow, int64_t ncol, int halo_top, int halo_bot) {
    const int64_t n = nrow * ncol;
    const int64_t npad = (n + PC_CHUNK - 1) / PC_CHUNK * PC_CHUNK;
    PFD_TRY(pfd_reserve(h, h->dir, (size_t)npad));
    PFD_TRY(pfd_reserve(h, h->counters, 8 * sizeof(unsigned long long)));
    if (!h->h_counters) PFD_CUDA(h, cudaHostAlloc((void**)&h->h_counters, 8 * sizeof(unsigned long long), cudaHostAllocDefault));
    PFD_CUDA(h, cudaMemsetAsync(h->counters.p, 0, 8 * sizeof(unsigned long long), h->stream));
    if (npad > n) PFD_CUDA(h, cudaMemsetAsync((uint8_t*)h->dir.p + n, 0xFF, (size_t)(npad - n), h->stream));
    unsigned int* flag = reinterpret_cast<unsigned int*>((unsigned long long*)h->counters.p + 3);
    h->nrow = nrow, h->ncol = ncol, h->n = n;
    h->dir_off = 0;
    h->tiled = true;
    h->parsed = true;
    h->have_upmask = false;
    h->mg_halo_top = halo_top, h->mg_halo_bot = halo_bot;
    TileCtx T;
    PFD_TRY(tiles_setup(h, T, true));
    {
        StageTimer t(h, PFD_STAGE_TILE_A);
        PFD_CUDA(h, cudaMemsetAsync(T.B[0].acc, 0, T.arr, h->stream));
        halo_slots_init_kernel<<<grid_for(2 * T.ntx * TL_RING, 256), 256, 0, h->stream>>>(T.B[0], T.term, T.term_h, T.ntx, T.nty);
        PFD_LAUNCH_CHECK(h);
        PhaseAArgs A{};
        A.nrow = nrow, A.ncol = ncol, A.ntx = T.ntx;
        A.loccnt = (uint2*)h->tile_loc.p, A.W = T.B[0].acc, A.s_nxt = T.B[0].nxt, A.s_rh = T.B[0].rh, A.s_ch = T.B[0].ch;
        A.s_term = T.term, A.s_term_h = T.term_h;
        A.d8 = d8_owned, A.dir_out = (uint8_t*)h->dir.p, A.invalid_flag = flag;
        A.al4 = (ncol % 4 == 0) && ((uintptr_t)d8_owned % 4 == 0);
        A.halo_top = halo_top, A.halo_bot = halo_bot;
        tl_launch_phase_a<true>(dim3((unsigned)T.ntx, (unsigned)T.nty), h->stream, A);
        PFD_LAUNCH_CHECK(h);
    }
    const int64_t nblk = npad / PC_CHUNK;
    {
        StageTimer t(h, PFD_STAGE_PITS);
        PFD_TRY(pfd_reserve(h, h->blk_counts, (size_t)nblk * sizeof(uint32_t)));
        PFD_TRY(pfd_reserve(h, h->blk_offsets, (size_t)(nblk + 1) * sizeof(unsigned long long)));
        pit_count_kernel<<<(unsigned)nblk, 256, 0, h->stream>>>((const uint8_t*)h->dir.p, (uint32_t*)h->blk_counts.p,
                                                               (unsigned long long*)h->counters.p);
        PFD_LAUNCH_CHECK(h);
        scan_counts_kernel<<<1, 1024, 0, h->stream>>>((const uint32_t*)h->blk_counts.p, nblk, (unsigned long long*)h->blk_offsets.p);
        PFD_LAUNCH_CHECK(h);
    }
    return PFD_OK;
}

// the pit list of the block once the host knows the count (h->n_pits etc. already set from the counters)
static int tiled_fused_pits(pfd_handle* h) {
    const int64_t npad = (h->n + PC_CHUNK - 1) / PC_CHUNK * PC_CHUNK;
    PFD_TRY(pfd_reserve(h, h->pits, (size_t)std::max<int64_t>(h->n_pits, 1) * sizeof(cell_t)));
    PFD_TRY(pfd_reserve(h, h->pit_outlet, (size_t)std::max<int64_t>(h->n_pits, 1)));
    if (h->n_pits > 0) {
        pit_scatter_kernel<<<(unsigned)(npad / PC_CHUNK), 256, 0, h->stream>>>((const uint8_t*)h->dir.p, (const unsigned long long*)h->blk_offsets.p,
                                                                             (cell_t*)h->pits.p, (uint8_t*)h->pit_outlet.p);
        PFD_LAUNCH_CHECK(h);
    }
    return PFD_OK;
}

// basin id (global pit ordinal + 1) of every pit of the block at the pit's own cell of the basin buffer
static int tiled_fused_stash(pfd_handle* h, uint32_t* basin_dev, int64_t pit_id_offset) {
    if (basin_dev && h->n_pits > 0) {
        stash_pit_ids_kernel<<<grid_for(h->n_pits, 256, 1, 148 * 16), 256, 0, h->stream>>>((const cell_t*)h->pits.p, h->n_pits, 0,
                                                                                          (unsigned long long)pit_id_offset, basin_dev);
        PFD_LAUNCH_CHECK(h);
    }
    return PFD_OK;
}

// keep the one-hop links of phase A (the remote inflow is added along them later); THE solve of the local reduced graph
static int tiled_local_solve1(pfd_handle* h, TileCtx& T) {
    PFD_CUDA(h, cudaMemcpyAsync(T.init.nxt, T.B[0].nxt, T.arr, cudaMemcpyDeviceToDevice, h->stream));
    StageTimer t(h, PFD_STAGE_TILE_B);
    return slots_solve(h, T.B, T.recv, T.nslots, T.flag, 1, tiles_protect(h, T, true));
}

// this rank's contribution to the boundary tables
static int tiled_local_fill(pfd_handle* h, TileCtx& T, int rank, const uint32_t* pit_stash) {
    const long long nb = boundary_entries(h);
    PFD_CUDA(h, cudaMemsetAsync(h->btab.p, 0, (size_t)(4 * nb) * sizeof(uint32_t), h->stream));
    BoundaryTables bt;
    boundary_tables(h, bt);
    boundary_fill_kernel<<<grid_for(2 * h->ncol, 256), 256, 0, h->stream>>>(
        T.B[0], T.B[1], (const int*)(T.flag + 3), T.term, T.term_h, h->nrow, h->ncol, T.ntx, T.nty, rank, h->mg_halo_top,
        h->mg_halo_bot, bt, pit_stash);
    PFD_LAUNCH_CHECK(h);
    return PFD_OK;
}

extern "C" int pfd_tiled_parse(pfd_handle* h, const uint8_t* d8_block, int64_t nrow_owned, int64_t ncol, int halo_top,
                               int halo_bot, int64_t glob_row0, void* idxs_ds_out, int idx_dtype, int64_t* n_valid,
                               int64_t* n_pits) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    PFD_TRY(check_shape(h, nrow_owned + 2, ncol, "pfd_tiled_parse"));
    if (!d8_block) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_tiled_parse: d8_block is null");
    if ((halo_top | halo_bot) & ~1) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_tiled_parse: halo flags must be 0 or 1");
    if (idxs_ds_out && idx_dtype != PFD_I32 && idx_dtype != PFD_U32 && idx_dtype != PFD_I64)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_tiled_parse: idx_dtype must be int32, uint32 or int64");
    invalidate(h);
    const int64_t next = (nrow_owned + halo_top + halo_bot) * ncol;
    const void* d8_dev = nullptr;
    PFD_TRY(pfd_stage_in(h, d8_block, (size_t)next, 0, &d8_dev));
    h->mg_fused = h->use_tiles && h->fuse_parse;
    if (h->mg_fused) {  // the block is parsed inside phase A; idxs_ds is written by pfd_tiled_finish
        h->mg_idxs_user = idxs_ds_out, h->mg_idx_dtype = idx_dtype, h->mg_glob_row0 = glob_row0;
        int rc = tiled_fused_begin(h, (const uint8_t*)d8_dev + (int64_t)halo_top * ncol, nrow_owned, ncol, halo_top, halo_bot);
        unsigned long long hc[4] = {0, 0, 0, 0};
        if (rc == PFD_OK && cudaMemcpyAsync(hc, h->counters.p, sizeof(hc), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) rc = PFD_ERR_CUDA;
        if (rc == PFD_OK && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = pfd_fail(h, PFD_ERR_CUDA, "pfd_tiled_parse: phase A failed");
        if (rc == PFD_OK && (hc[3] & 1ull))
            rc = pfd_fail(h, PFD_ERR_INVALID_D8, "raster holds values outside the D8 code set {0,1,2,4,8,16,32,64,128,247,255}");
        if (rc != PFD_OK) {
            invalidate(h);
            return rc;
        }
        h->n_valid = (int64_t)hc[0], h->n_pits = (int64_t)hc[1], h->n_outlets = (int64_t)hc[2];
        PFD_TRY(tiled_fused_pits(h));
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
        stage_collect(h);
        if (n_valid) *n_valid = h->n_valid;
        if (n_pits) *n_pits = h->n_pits;
        return PFD_OK;
    }
    void* idxs_dev = nullptr;
    const size_t ibytes = (size_t)(nrow_owned * ncol) * pfd_dtype_size(idx_dtype);
    if (idxs_ds_out) PFD_TRY(pfd_stage_out(h, idxs_ds_out, ibytes, 1, &idxs_dev));
    PFD_TRY(parse_device(h, (const uint8_t*)d8_dev, nrow_owned, ncol, idxs_dev, idx_dtype, halo_top, halo_bot, glob_row0));
    if (idxs_ds_out) PFD_TRY(pfd_finish_out(h, idxs_ds_out, idxs_dev, ibytes));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    h->mg_halo_top = halo_top;
    h->mg_halo_bot = halo_bot;
    if (n_valid) *n_valid = h->n_valid;
    if (n_pits) *n_pits = h->n_pits;
    return PFD_OK;
}

// Local part: tiles + local reduced graph; fills this rank's contribution to the boundary tables.
// table_dev / table_len: the device buffer to all-reduce (uint32 sum) across ranks before pfd_tiled_finish.
extern "C" int pfd_tiled_local(pfd_handle* h, int rank, int nranks, int64_t pit_id_offset, uint32_t* basins_out,
                               void** table_dev, int64_t* table_len) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    if (!h->parsed) return pfd_fail(h, PFD_ERR_STATE, "pfd_tiled_local: call pfd_tiled_parse first");
    if (nranks < 1 || rank < 0 || rank >= nranks) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_tiled_local: bad rank");
    if (nranks > 1 && rank < nranks - 1 && (h->nrow % TL_H) != 0)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_tiled_local: row blocks (except the last) must be multiples of 64 rows");
    if (basins_out && !pfd_is_device_ptr(basins_out))
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_tiled_local: basins_out must be a device buffer (it carries the pit ids)");
    h->mg_rank = rank;
    h->mg_nranks = nranks;
    h->mg_basins = basins_out;
    TileCtx T;
    PFD_TRY(tiles_setup(h, T, h->mg_fused || nranks > 1));
    if (h->mg_fused) {
        if (!basins_out) {  // the pit ids travel through a basin buffer: use the handle's own
            PFD_TRY(pfd_reserve(h, h->basins, (size_t)h->n * sizeof(uint32_t)));
            h->mg_basins = basins_out = (uint32_t*)h->basins.p;
        }
        PFD_TRY(tiled_fused_stash(h, basins_out, pit_id_offset));
    } else {
        PFD_TRY(tiles_phase_a(h, T, basins_out, (unsigned long long)pit_id_offset));
    }
    const long long nb = boundary_entries(h);
    PFD_TRY(pfd_reserve(h, h->btab, (size_t)std::max<long long>(4 * nb, 1) * sizeof(uint32_t)));
    if (nranks > 1) {
        PFD_TRY(tiled_local_solve1(h, T));
        PFD_TRY(tiled_local_fill(h, T, rank, h->mg_fused ? basins_out : nullptr));
    }
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    if (table_dev) *table_dev = h->btab.p;
    if (table_len) *table_len = 4 * nb;
    return PFD_OK;
}

// After the boundary tables were all-reduced: boundary graph, remote inflow along the local chains, per-tile finalisation.
static int tiled_finish_impl(pfd_handle* h, int32_t* rank_out, int32_t* uparea_out, uint32_t* basins_out);

extern "C" int pfd_tiled_finish(pfd_handle* h, int32_t* rank_out, int32_t* uparea_out, uint32_t* basins_out) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    return tiled_finish_impl(h, rank_out, uparea_out, basins_out);
}

static int tiled_finish_impl(pfd_handle* h, int32_t* rank_out, int32_t* uparea_out, uint32_t* basins_out) {
    if (!h->parsed || h->mg_nranks < 1) return pfd_fail(h, PFD_ERR_STATE, "pfd_tiled_finish: call pfd_tiled_local first");
    if (basins_out != h->mg_basins && !(h->mg_fused && !basins_out))
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_tiled_finish: basins_out differs from pfd_tiled_local");
    if (h->mg_fused && !basins_out) basins_out = h->mg_basins;  // internal buffer (results not requested)
    const size_t b4 = (size_t)h->n * 4;
    void *rk = nullptr, *up = nullptr, *ix = nullptr;
    if (rank_out) PFD_TRY(pfd_stage_out(h, rank_out, b4, 2, &rk));
    if (uparea_out) PFD_TRY(pfd_stage_out(h, uparea_out, b4, 3, &up));
    const size_t ibytes = (size_t)h->n * pfd_dtype_size(h->mg_idx_dtype);
    if (h->mg_fused && h->mg_idxs_user) PFD_TRY(pfd_stage_out(h, h->mg_idxs_user, ibytes, 1, &ix));
    TileCtx T;
    PFD_TRY(tiles_setup(h, T, h->mg_fused || h->mg_nranks > 1));
    if (h->mg_nranks > 1) {
        const long long nb = boundary_entries(h);
        // boundary graph: 12 node arrays + results, solved redundantly on every rank
        PFD_TRY(pfd_reserve(h, h->bgraph, (size_t)(14 * nb) * sizeof(uint32_t) + 256));
        uint32_t* base = (uint32_t*)h->bgraph.p;
        int a = 0;
        auto take = [&]() { return base + (size_t)nb * a++; };
        SlotBuf G[2];
        uint32_t* grecv[2];
        for (int side = 0; side < 2; ++side) {
            G[side].nxt = take();
            G[side].rh = take();
            G[side].ch = take();
            grecv[side] = take();
        }
        G[0].acc = G[1].acc = take();
        uint32_t *gterm = take(), *gterm_h = take(), *gbasin = take();
        int32_t* grank = (int32_t*)take();
        unsigned int* gflag = (unsigned int*)take();  // + 256 spare bytes behind it
        BoundaryTables bt;
        boundary_tables(h, bt);
        boundary_build_kernel<<<grid_for(nb, 256), 256, 0, h->stream>>>(bt, nb, G[0], gterm, gterm_h);
        PFD_LAUNCH_CHECK(h);
        SlotProtect noprot{1, 0, 0, 0};
        PFD_TRY(slots_solve(h, G, grecv, nb, gflag, 0, noprot));
        slots_finalize_kernel<<<grid_for(nb, 256), 256, 0, h->stream>>>(G[0], G[1], (const int*)(gflag + 3), gterm, gterm_h, nb,
                                                                        grank, gbasin, nullptr, 0, 0);
        PFD_LAUNCH_CHECK(h);
        // No second solve of the local reduced graph: the first one already left, for every ring node, its local
        // terminal, the hops to it and the local inflow. The halo slots now become pit-like terminals carrying (rank,
        // basin) of the neighbour's entry -- picked up by the finalisation below --, and the remote inflow of my own
        // boundary entries is added along their local chains (one-hop links saved before the first solve).
        boundary_writeback_kernel<<<grid_for(2 * h->ncol, 256), 256, 0, h->stream>>>(
            grank, gbasin, G[0].acc, h->nrow, h->ncol, T.ntx, h->mg_rank, h->mg_halo_top, h->mg_halo_bot, T.B[0].acc,
            T.term, T.term_h, T.init.nxt, T.B[0], T.B[1], (const int*)(T.flag + 3), T.nslots);
        PFD_LAUNCH_CHECK(h);
        {
            StageTimer t(h, PFD_STAGE_TILE_B);
            slots_finalize_kernel<<<grid_for(T.nslots, 256, 2, 148 * 16), 256, 0, h->stream>>>(
                T.B[0], T.B[1], (const int*)(T.flag + 3), T.term, T.term_h, T.nslots, T.srank, T.sbasin,
                h->mg_fused ? basins_out : nullptr, h->ncol, T.ntx, T.nty);
            PFD_LAUNCH_CHECK(h);
        }
    }
    if (h->mg_fused) {
        if (h->mg_nranks == 1) PFD_TRY(tiles_phase_b(h, T, basins_out));  // (several ranks: solved in pfd_tiled_local)
        PFD_TRY(tiles_phase_c(h, T, (int32_t*)rk, basins_out, (int32_t*)up, ix, h->mg_idx_dtype, (long long)h->mg_glob_row0 * h->ncol));
        if (ix) PFD_TRY(pfd_finish_out(h, h->mg_idxs_user, ix, ibytes));
    } else {
        if (h->mg_nranks == 1) PFD_TRY(tiles_phase_b(h, T));
        PFD_TRY(tiles_phase_c(h, T, (int32_t*)rk, basins_out, (int32_t*)up));
    }
    if (rank_out) PFD_TRY(pfd_finish_out(h, rank_out, rk, b4));
    if (uparea_out) PFD_TRY(pfd_finish_out(h, uparea_out, up, b4));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    return PFD_OK;
}


// ---------------------------------------------------------------------------------------------------------
// NCCL plumbing for the row-tiled solve (one process per GPU; the caller distributes the unique id)
// ---------------------------------------------------------------------------------------------------------
#define PFD_NCCL(h, call)                                                                                         \
    do {                                                                                                          \
        ncclResult_t r__ = (call);                                                                                \
        if (r__ != ncclSuccess)                                                                                   \
            return pfd_fail((h), PFD_ERR_NCCL, std::string(#call) + ": " + ncclGetErrorString(r__));              \
    } while (0)

extern "C" int pfd_comm_unique_id(void* out, int64_t capacity) {
    if (!out || capacity < (int64_t)sizeof(ncclUniqueId)) return pfd_fail(nullptr, PFD_ERR_INVALID_ARG, "pfd_comm_unique_id: need 128 bytes");
    ncclUniqueId id;
    PFD_NCCL(nullptr, ncclGetUniqueId(&id));
    memcpy(out, &id, sizeof(id));
    return PFD_OK;
}

extern "C" int pfd_comm_init(pfd_handle* h, int rank, int nranks, const void* unique_id) {
    PFD_TRY(check_handle(h));
    if (!unique_id || nranks < 1 || rank < 0 || rank >= nranks) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_comm_init: bad argument");
    if (h->nccl_comm) return pfd_fail(h, PFD_ERR_STATE, "pfd_comm_init: communicator already initialised");
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    ncclComm_t comm;
    PFD_NCCL(h, ncclCommInitRank(&comm, nranks, id, rank));
    h->nccl_comm = (void*)comm;
    h->mg_rank = rank;
    h->mg_nranks = nranks;
    // buffers of exchange #1 are allocated now, so that a rank that fails later can still tell the others
    PFD_TRY(pfd_reserve(h, h->counters, 8 * sizeof(unsigned long long)));
    PFD_TRY(pfd_reserve(h, h->mg_counts, (size_t)(4 * nranks + 4) * sizeof(unsigned long long)));
    if (h->h_gather && h->h_gather_ranks < nranks) {
        cudaFreeHost(h->h_gather);
        h->h_gather = nullptr;
    }
    if (!h->h_gather) {
        PFD_CUDA(h, cudaHostAlloc((void**)&h->h_gather, (size_t)(4 * nranks) * sizeof(unsigned long long), cudaHostAllocDefault));
        h->h_gather_ranks = nranks;
    }
    return PFD_OK;
}

extern "C" int pfd_comm_destroy(pfd_handle* h) {
    if (h && h->nccl_comm) {
        cudaSetDevice(h->device);
        ncclCommDestroy((ncclComm_t)h->nccl_comm);
        h->nccl_comm = nullptr;
    }
    return PFD_OK;
}

extern "C" int pfd_comm_barrier(pfd_handle* h) {
    PFD_TRY(check_handle(h));
    if (!h->nccl_comm) return pfd_fail(h, PFD_ERR_STATE, "pfd_comm_barrier: no communicator");
    PFD_TRY(pfd_reserve(h, h->counters, 8 * sizeof(unsigned long long)));
    unsigned int* p = (unsigned int*)((unsigned long long*)h->counters.p + 7);
    PFD_NCCL(h, ncclAllReduce(p, p, 1, ncclUint32, ncclSum, (ncclComm_t)h->nccl_comm, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    return PFD_OK;
}

// parse + order-free solve of one row block, exchanges over NCCL. Outputs (device or host, any may be NULL except
// that basins_out, when given, must be a device buffer) hold the nrow_owned rows of this rank.
static int flow_all_tiled_unfused(pfd_handle* h, const uint8_t* d8_block, int64_t nrow_owned, int64_t ncol, int halo_top,
                                     int halo_bot, int64_t glob_row0, void* idxs_ds_out, int idx_dtype, int32_t* rank_out,
                                     int32_t* uparea_out, uint32_t* basins_out, int64_t* n_valid, int64_t* n_pits_global) {
    PFD_TRY(check_handle(h));
    const int nranks = h->nccl_comm ? h->mg_nranks : 1, rank = h->nccl_comm ? h->mg_rank : 0;
    int64_t nv = 0, np = 0;
    cudaEventRecord(h->ev_total[0], h->stream);
    PFD_TRY(pfd_tiled_parse(h, d8_block, nrow_owned, ncol, halo_top, halo_bot, glob_row0, idxs_ds_out, idx_dtype, &nv, &np));
    // exchange #1: pit counts -> global basin id offset of this block (blocks are in linear-index order)
    long long offset = 0, total = np;
    if (nranks > 1) {
        PFD_TRY(pfd_reserve(h, h->mg_counts, (size_t)(nranks + 1) * sizeof(unsigned long long)));
        unsigned long long* d = (unsigned long long*)h->mg_counts.p;
        unsigned long long mine = (unsigned long long)np;
        PFD_CUDA(h, cudaMemcpyAsync(d + nranks, &mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
        PFD_NCCL(h, ncclAllGather(d + nranks, d, 1, ncclUint64, (ncclComm_t)h->nccl_comm, h->stream));
        std::vector<unsigned long long> all(nranks);
        PFD_CUDA(h, cudaMemcpyAsync(all.data(), d, (size_t)nranks * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
        total = 0;
        for (int g = 0; g < nranks; ++g) {
            if (g < rank) offset += (long long)all[g];
            total += (long long)all[g];
        }
    }
    if (total >= (1ll << 31)) return pfd_fail(h, PFD_ERR_UNSUPPORTED, "pfd_d8_flow_all_tiled: more than 2^31 pits");
    void* table = nullptr;
    int64_t table_len = 0;
    PFD_TRY(pfd_tiled_local(h, rank, nranks, offset, basins_out, &table, &table_len));
    // exchange #2: ONE all-reduce of the boundary tables
    if (nranks > 1 && table_len > 0) {
        PFD_NCCL(h, ncclAllReduce(table, table, (size_t)table_len, ncclUint32, ncclSum, (ncclComm_t)h->nccl_comm, h->stream));
    }
    PFD_TRY(pfd_tiled_finish(h, rank_out, uparea_out, basins_out));
    cudaEventRecord(h->ev_total[1], h->stream);
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev_total[0], h->ev_total[1]) == cudaSuccess) h->stage_ms[PFD_STAGE_TOTAL] = ms;
    if (n_valid) *n_valid = nv;
    if (n_pits_global) *n_pits_global = total;
    return PFD_OK;
}

// A rank that fails after exchange #1 tears the communicator down, so that its peers' collectives return an error
// instead of waiting forever.
static int tiled_abort(pfd_handle* h, int rc) {
    if (rc != PFD_OK && h->nccl_comm && h->mg_nranks > 1) {
        const std::string msg = h->err;
        ncclCommAbort((ncclComm_t)h->nccl_comm);
        h->nccl_comm = nullptr;
        h->err = msg + " (communicator aborted)";
    }
    return rc;
}

// parse + order-free solve of one row block, exchanges over NCCL. Outputs (device or host, any may be NULL) hold the
// nrow_owned rows of this rank. ONE program with the single-GPU step: the block is parsed inside phase A; the host
// waits once, for the pit counts, while the first local solve keeps the GPU busy.
//   exchange #1  all-gather of {n_valid, n_pits, n_outlets, flags} per rank -> global basin-id offset of the block, and
//                ERROR AGREEMENT: a rank whose block is invalid (or that failed locally) says so here, and every rank
//                returns an error instead of entering exchange #2;
//   exchange #2  ONE all-reduce (uint32 sum) of the boundary tables.
extern "C" int pfd_d8_flow_all_tiled(pfd_handle* h, const uint8_t* d8_block, int64_t nrow_owned, int64_t ncol, int halo_top,
                                     int halo_bot, int64_t glob_row0, void* idxs_ds_out, int idx_dtype, int32_t* rank_out,
                                     int32_t* uparea_out, uint32_t* basins_out, int64_t* n_valid, int64_t* n_pits_global) {
    PFD_TRY(check_handle(h));
    if (!(h->use_tiles && h->fuse_parse))
        return flow_all_tiled_unfused(h, d8_block, nrow_owned, ncol, halo_top, halo_bot, glob_row0, idxs_ds_out, idx_dtype, rank_out,
                                      uparea_out, basins_out, n_valid, n_pits_global);
    const int nranks = h->nccl_comm ? h->mg_nranks : 1, rank = h->nccl_comm ? h->mg_rank : 0;
    stage_reset(h);
    cudaEventRecord(h->ev_total[0], h->stream);
    // ---- local part 1 (any failure is carried into exchange #1, not returned early: the peers must not be left waiting)
    int rc = PFD_OK;
    const void* d8_dev = nullptr;
    if (!d8_block) rc = pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_d8_flow_all_tiled: d8_block is null");
    else if ((halo_top | halo_bot) & ~1) rc = pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_d8_flow_all_tiled: halo flags must be 0 or 1");
    else if (idxs_ds_out && idx_dtype != PFD_I32 && idx_dtype != PFD_U32 && idx_dtype != PFD_I64)
        rc = pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_d8_flow_all_tiled: idx_dtype must be int32, uint32 or int64");
    else if (nranks > 1 && rank < nranks - 1 && (nrow_owned % TL_H) != 0)
        rc = pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_d8_flow_all_tiled: row blocks (except the last) must be multiples of 64 rows");
    else if (basins_out && !pfd_is_device_ptr(basins_out))
        rc = pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_d8_flow_all_tiled: basins_out must be a device buffer (it carries the pit ids)");
    if (rc == PFD_OK) rc = check_shape(h, nrow_owned + 2, ncol, "pfd_d8_flow_all_tiled");
    invalidate(h);
    if (rc == PFD_OK) rc = pfd_stage_in(h, d8_block, (size_t)((nrow_owned + halo_top + halo_bot) * ncol), 0, &d8_dev);
    h->mg_fused = true;
    h->mg_idxs_user = idxs_ds_out, h->mg_idx_dtype = idx_dtype, h->mg_glob_row0 = glob_row0;
    h->mg_rank = rank, h->mg_nranks = nranks;
    if (rc == PFD_OK && !basins_out) {
        rc = pfd_reserve(h, h->basins, (size_t)(nrow_owned * ncol) * sizeof(uint32_t));
        basins_out = (uint32_t*)h->basins.p;
    }
    h->mg_basins = basins_out;
    if (rc == PFD_OK) rc = tiled_fused_begin(h, (const uint8_t*)d8_dev + (int64_t)halo_top * ncol, nrow_owned, ncol, halo_top, halo_bot);
    // ---- exchange #1
    unsigned long long mine[4] = {0, 0, 0, 0};
    unsigned long long* all = mine;
    if (nranks > 1) {
        unsigned long long* d = (unsigned long long*)h->mg_counts.p;  // [4 * nranks] gathered, [4] send slot behind
        unsigned long long* send = d + 4 * nranks;
        if (rc != PFD_OK) {  // failure marker in the flags word
            const unsigned long long fail[4] = {0, 0, 0, (1ull << 32) | (unsigned long long)rc};
            cudaMemcpyAsync(send, fail, sizeof(fail), cudaMemcpyHostToDevice, h->stream);
        } else {
            cudaMemcpyAsync(send, h->counters.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, h->stream);
        }
        if (ncclAllGather(send, d, 4, ncclUint64, (ncclComm_t)h->nccl_comm, h->stream) != ncclSuccess)
            return tiled_abort(h, pfd_fail(h, PFD_ERR_NCCL, "pfd_d8_flow_all_tiled: ncclAllGather failed"));
        cudaMemcpyAsync(h->h_gather, d, (size_t)(4 * nranks) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream);
        all = h->h_gather;
    } else if (rc == PFD_OK) {
        cudaMemcpyAsync(h->h_counters, h->counters.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream);
        all = h->h_counters;
    }
    cudaEventRecord(h->ev_copy, h->stream);
    // the first local solve does not need the pit numbering: it keeps the GPU busy while the host waits for the counts
    TileCtx T;
    int rc_solve = PFD_OK;
    if (rc == PFD_OK) {
        rc_solve = tiles_setup(h, T, true);
        if (rc_solve == PFD_OK) rc_solve = pfd_reserve(h, h->btab, (size_t)std::max<long long>(4 * boundary_entries(h), 1) * sizeof(uint32_t));
        if (rc_solve == PFD_OK && nranks > 1) rc_solve = tiled_local_solve1(h, T);
    }
    if (cudaEventSynchronize(h->ev_copy) != cudaSuccess) return tiled_abort(h, pfd_fail(h, PFD_ERR_CUDA, "pfd_d8_flow_all_tiled: exchange #1 failed"));
    // ---- agreement
    int bad_rank = -1;
    bool invalid_codes = false;
    long long offset = 0, total = 0;
    for (int g = 0; g < nranks; ++g) {
        const unsigned long long f = (nranks > 1 || rc == PFD_OK) ? all[4 * g + 3] : ((1ull << 32) | (unsigned long long)rc);
        if ((f >> 32) && bad_rank < 0) bad_rank = g;
        if (f & 1ull) {
            invalid_codes = true;
            if (bad_rank < 0) bad_rank = g;
        }
        if (g < rank) offset += (long long)all[4 * g + 1];
        total += (long long)all[4 * g + 1];
    }
    if (bad_rank >= 0 || total >= (1ll << 31) || total == 0) {
        cudaStreamSynchronize(h->stream);
        cudaGetLastError();
        const std::string own = h->err;
        invalidate(h);
        if (rc != PFD_OK) {
            h->err = own;
            return rc;
        }
        if (bad_rank == rank && invalid_codes)
            return pfd_fail(h, PFD_ERR_INVALID_D8, "raster holds values outside the D8 code set {0,1,2,4,8,16,32,64,128,247,255}");
        if (bad_rank >= 0 && invalid_codes)
            return pfd_fail(h, PFD_ERR_INVALID_D8, "the row block of rank " + std::to_string(bad_rank) + " holds values outside the D8 code set");
        if (bad_rank >= 0)
            return pfd_fail(h, PFD_ERR_NCCL, "rank " + std::to_string(bad_rank) + " failed before exchange #1 (status " +
                                                 std::to_string((int)(all[4 * bad_rank + 3] & 0xFFFFFFFFull)) + ")");
        if (total == 0) return pfd_fail(h, PFD_ERR_NO_PITS, "Invalid FlwdirRaster: no pits found");
        return pfd_fail(h, PFD_ERR_UNSUPPORTED, "pfd_d8_flow_all_tiled: more than 2^31 pits");
    }
    h->n_valid = (int64_t)all[4 * rank + 0], h->n_pits = (int64_t)all[4 * rank + 1], h->n_outlets = (int64_t)all[4 * rank + 2];
    // ---- local part 2, exchange #2, finish: a failure from here on aborts the communicator
    rc = rc_solve;
    if (rc == PFD_OK) rc = tiled_fused_pits(h);
    if (rc == PFD_OK) rc = tiled_fused_stash(h, basins_out, offset);
    if (rc == PFD_OK && nranks > 1) rc = tiled_local_fill(h, T, rank, basins_out);
    if (rc != PFD_OK) return tiled_abort(h, rc);
    if (nranks > 1) {
        const long long nb = boundary_entries(h);
        if (nb > 0 && ncclAllReduce(h->btab.p, h->btab.p, (size_t)(4 * nb), ncclUint32, ncclSum, (ncclComm_t)h->nccl_comm, h->stream) != ncclSuccess)
            return tiled_abort(h, pfd_fail(h, PFD_ERR_NCCL, "pfd_d8_flow_all_tiled: ncclAllReduce failed"));
    }
    rc = tiled_finish_impl(h, rank_out, uparea_out, basins_out == (uint32_t*)h->basins.p ? nullptr : basins_out);
    if (rc != PFD_OK) return tiled_abort(h, rc);
    cudaEventRecord(h->ev_total[1], h->stream);
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev_total[0], h->ev_total[1]) == cudaSuccess) h->stage_ms[PFD_STAGE_TOTAL] = ms;
    if (n_valid) *n_valid = h->n_valid;
    if (n_pits_global) *n_pits_global = total;
    return PFD_OK;
}

// make rank / basins / uparea available in the handle's own cache buffers via the tile solver
static int tiles_ensure(pfd_handle* h, bool want_rank, bool want_basins, bool want_uparea) {
    want_rank = want_rank && !h->have_rank;
    want_basins = want_basins && !h->have_basins;
    want_uparea = want_uparea && !h->have_uparea;
    if (!want_rank && !want_basins && !want_uparea) return PFD_OK;
    if (h->tiled) return pfd_fail(h, PFD_ERR_UNSUPPORTED, "this handle holds a row block: only the pfd_tiled_* entry points apply");
    const size_t b4 = (size_t)h->n * 4;
    if (want_rank) PFD_TRY(pfd_reserve(h, h->rank, b4));
    if (want_basins) PFD_TRY(pfd_reserve(h, h->basins, b4));
    if (want_uparea) PFD_TRY(pfd_reserve(h, h->uparea, b4));
    PFD_TRY(tiles_solve(h, want_rank ? (int32_t*)h->rank.p : nullptr, want_basins ? (uint32_t*)h->basins.p : nullptr,
                        want_uparea ? (int32_t*)h->uparea.p : nullptr));
    h->have_rank |= want_rank;
    h->have_basins |= want_basins;
    h->have_uparea |= want_uparea;
    return PFD_OK;
}

static int count_ranked(pfd_handle* h, const int32_t* rank_dev, int64_t* out) {
    unsigned long long* ctr = (unsigned long long*)h->counters.p + 6;
    PFD_CUDA(h, cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), h->stream));
    count_ranked_kernel<<<grid_for(h->n, 256, 8, 148 * 8), 256, 0, h->stream>>>(rank_dev, h->n, ctr);
    PFD_LAUNCH_CHECK(h);
    unsigned long long v = 0;
    PFD_CUDA(h, cudaMemcpyAsync(&v, ctr, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    *out = (int64_t)v;
    return PFD_OK;
}

extern "C" int pfd_set_option(pfd_handle* h, const char* name, int64_t value) {
    PFD_TRY(check_handle(h));
    if (name && strcmp(name, "tiles") == 0) {
        h->use_tiles = value ? 1 : 0;
        return PFD_OK;
    }
    if (name && strcmp(name, "fuse_parse") == 0) {
        h->fuse_parse = value ? 1 : 0;
        return PFD_OK;
    }
    if (name && strcmp(name, "hand_pathsum") == 0) {
        h->hand_pathsum = value ? 1 : 0;
        return PFD_OK;
    }
    if (name && strcmp(name, "tile_sweeps") == 0) {
        // 1 (default): tile-dataflow sweeps unless the BFS ordering of this raster is already cached (then its level
        // replay is the cheaper one: 32 vs 46 ms for Strahler at 32768^2; the ordering itself costs 40 ms);
        // 2: always tile-dataflow; 0: always level replays over the BFS order
        h->tile_sweeps = (int)value;
        return PFD_OK;
    }
    if (name && strcmp(name, "sweep_max_passes") == 0) {  // profiling only
        h->ts_max_passes = (int)value;
        return PFD_OK;
    }
    if (name && strcmp(name, "release_scratch") == 0) {  // give the staging buffers back (very large rasters)
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
        for (auto& b : h->scratch) pfd_release(b);
        for (auto& b : h->fill_bufs) pfd_release(b);
        pfd_release(h->hand_root);
        pfd_release(h->hand_sum);
        pfd_release(h->hand_slots);
        return PFD_OK;
    }
    return pfd_fail(h, PFD_ERR_INVALID_ARG, std::string("pfd_set_option: unknown option ") + (name ? name : "(null)"));
}

extern "C" int64_t pfd_get_info(const pfd_handle* h, const char* name) {
    if (!h || !name) return -1;
    if (strcmp(name, "tiles") == 0) return h->use_tiles;
    if (strcmp(name, "fuse_parse") == 0) return h->fuse_parse;
    if (strcmp(name, "have_upmask") == 0) return h->have_upmask ? 1 : 0;
    if (strcmp(name, "tile_rounds") == 0) return h->tile_rounds;
    if (strcmp(name, "tile_sweeps") == 0) return h->tile_sweeps;
    if (strcmp(name, "hand_pathsum") == 0) return h->hand_pathsum;
    if (strcmp(name, "hand_engine") == 0) return h->hand_engine;
    if (strcmp(name, "sweep_passes") == 0) return h->sweep_passes;
    if (strcmp(name, "sweep_visits") == 0) return h->sweep_visits;
    if (strcmp(name, "nlevels") == 0) return h->nlevels;
    if (strcmp(name, "nnodes") == 0) return h->nnodes;
    if (strcmp(name, "n_pits") == 0) return h->n_pits;
    if (strcmp(name, "n_outlets") == 0) return h->n_outlets;
    if (strcmp(name, "n_valid") == 0) return h->n_valid;
    if (strcmp(name, "num_sms") == 0) return h->num_sms;
    return -1;
}

// ---------------------------------------------------------------------------------------------------------
// order
// ---------------------------------------------------------------------------------------------------------
template <class K>
static int coop_grid(pfd_handle* h, K kernel, int threads, int64_t max_useful_blocks, int* grid) {
    int per_sm = 0;
    PFD_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0));
    if (per_sm < 1) return pfd_fail(h, PFD_ERR_CUDA, "kernel cannot be made resident");
    int64_t g = (int64_t)per_sm * h->num_sms;
    if (g > max_useful_blocks) g = std::max<int64_t>(1, max_useful_blocks);
    *grid = (int)g;
    return PFD_OK;
}

static int build_schedule(pfd_handle* h) {
    std::vector<SweepSeg> segs;
    const int nlev = (int)h->nlevels;
    int l = 0;
    while (l < nlev) {
        const int64_t size = h->h_level_off[l + 1] - h->h_level_off[l];
        if (size > SW_SOLO_MAX) {
            segs.push_back(SweepSeg{l, 1, 0, 0});
            ++l;
        } else {
            int l2 = l;
            while (l2 < nlev && h->h_level_off[l2 + 1] - h->h_level_off[l2] <= SW_SOLO_MAX) ++l2;
            segs.push_back(SweepSeg{l, l2 - l, 1, 0});
            l = l2;
        }
    }
    h->nsegs = (int)segs.size();
    h->max_level_size = 0;
    for (int i = 0; i < nlev; ++i) h->max_level_size = std::max<int64_t>(h->max_level_size, (int64_t)(h->h_level_off[i + 1] - h->h_level_off[i]));
    PFD_TRY(pfd_reserve(h, h->segs, std::max<size_t>(1, segs.size()) * sizeof(SweepSeg)));
    if (!segs.empty())
        PFD_CUDA(h, cudaMemcpyAsync(h->segs.p, segs.data(), segs.size() * sizeof(SweepSeg), cudaMemcpyHostToDevice, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));  // segs is a local vector
    return PFD_OK;
}

template <class Op, bool UP>
static int run_sweep(pfd_handle* h, Op op, int skip_level0) {
    if (h->nsegs == 0) return PFD_OK;
    SweepParams P;
    P.seq = (const cell_t*)h->seq.p;
    P.level_off = (const long long*)h->level_off.p;
    P.segs = (const SweepSeg*)h->segs.p;
    P.nsegs = h->nsegs;
    P.skip_level0 = skip_level0;
    int grid = 1;
    PFD_TRY(coop_grid(h, sweep_kernel<Op, UP>, SW_THREADS, (h->max_level_size + SW_THREADS - 1) / SW_THREADS, &grid));
    void* args[] = {(void*)&P, (void*)&op};
    StageTimer t(h, PFD_STAGE_SWEEP);
    PFD_CUDA(h, cudaLaunchCooperativeKernel((void*)sweep_kernel<Op, UP>, dim3(grid), dim3(SW_THREADS), args, 0, h->stream));
    h->launches++;
    return PFD_OK;
}

static int order_impl(pfd_handle* h, bool want_rank, bool want_basins) {
    if (!h->parsed) return pfd_fail(h, PFD_ERR_STATE, "pfd_order: no raster parsed on this handle");
    if (h->tiled) return pfd_fail(h, PFD_ERR_UNSUPPORTED, "this handle holds a row block: only the pfd_tiled_* entry points apply");
    const int64_t n = h->n;
    if (tiles_usable(h)) {  // scattered rank/basins writes of the BFS are replaced by the tile solver
        PFD_TRY(tiles_ensure(h, want_rank, want_basins, false));
        want_rank = want_basins = false;
    }
    if (!h->ordered) {
        StageTimer t(h, PFD_STAGE_ORDER);
        PFD_TRY(pfd_reserve(h, h->seq, (size_t)n * sizeof(cell_t)));
        if (want_basins) {
            PFD_TRY(pfd_reserve(h, h->bseq, (size_t)n * sizeof(uint32_t)));
            PFD_TRY(pfd_reserve(h, h->basins, (size_t)n * sizeof(uint32_t)));
            PFD_CUDA(h, cudaMemsetAsync(h->basins.p, 0, (size_t)n * sizeof(uint32_t), h->stream));
        }
        if (want_rank) {
            PFD_TRY(pfd_reserve(h, h->rank, (size_t)n * sizeof(int32_t)));
            order_init_rank_kernel<<<grid_for(n, 256, 4, 1ll << 30), 256, 0, h->stream>>>((const uint8_t*)h->dir.p, n, (int32_t*)h->rank.p);
            PFD_LAUNCH_CHECK(h);
        }
        if (h->level_cap == 0) h->level_cap = std::min<int64_t>(std::max<int64_t>(n, 1), 1 << 20);
        PFD_TRY(pfd_reserve(h, h->level_off, (size_t)(h->level_cap + 1) * sizeof(long long)));
        PFD_TRY(pfd_reserve(h, h->bfs_state, sizeof(BfsState)));
        const int64_t nstatus = n / BFS_CHUNK + 2;
        PFD_TRY(pfd_reserve(h, h->chunk_status, (size_t)nstatus * sizeof(unsigned long long)));
        PFD_CUDA(h, cudaMemsetAsync(h->chunk_status.p, 0, (size_t)nstatus * sizeof(unsigned long long), h->stream));
        BfsState st;
        memset(&st, 0, sizeof(st));
        st.slot_start[0] = 0;
        st.slot_end[0] = (unsigned long long)h->n_pits;
        PFD_CUDA(h, cudaMemcpyAsync(h->bfs_state.p, &st, sizeof(st), cudaMemcpyHostToDevice, h->stream));
        PFD_CUDA(h, cudaMemsetAsync(h->level_off.p, 0, sizeof(long long), h->stream));
        if (h->n_pits > 0) {
            order_init_pits_kernel<<<grid_for(h->n_pits, 256), 256, 0, h->stream>>>(
                (const cell_t*)h->pits.p, h->n_pits, (cell_t*)h->seq.p, want_basins ? (uint32_t*)h->bseq.p : nullptr,
                want_rank ? (int32_t*)h->rank.p : nullptr, want_basins ? (uint32_t*)h->basins.p : nullptr);
            PFD_LAUNCH_CHECK(h);
        }
        for (;;) {
            BfsParams P;
            PFD_TRY(ensure_upmask(h));
            P.upmask = (const uint8_t*)h->upmask.p;
            P.seq = (cell_t*)h->seq.p;
            P.bseq = (uint32_t*)h->bseq.p;
            P.rank = (int32_t*)h->rank.p;
            P.basins = (uint32_t*)h->basins.p;
            P.level_off = (long long*)h->level_off.p;
            P.level_cap = h->level_cap;
            P.status = (unsigned long long*)h->chunk_status.p;
            P.st = (BfsState*)h->bfs_state.p;
            P.ncol = h->ncol;
            void* args[] = {(void*)&P};
            void* kern = want_rank ? (want_basins ? (void*)bfs_kernel<true, true> : (void*)bfs_kernel<true, false>)
                                   : (want_basins ? (void*)bfs_kernel<false, true> : (void*)bfs_kernel<false, false>);
            int per_sm = 0;
            if (want_rank && want_basins) PFD_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bfs_kernel<true, true>, BFS_THREADS, 0));
            else if (want_rank) PFD_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bfs_kernel<true, false>, BFS_THREADS, 0));
            else if (want_basins) PFD_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bfs_kernel<false, true>, BFS_THREADS, 0));
            else PFD_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bfs_kernel<false, false>, BFS_THREADS, 0));
            if (per_sm < 1) return pfd_fail(h, PFD_ERR_CUDA, "bfs kernel cannot be made resident");
            int64_t grid = (int64_t)per_sm * h->num_sms;
            grid = std::max<int64_t>(1, std::min<int64_t>(grid, (n + BFS_CHUNK - 1) / BFS_CHUNK));
            {
                StageTimer tb(h, PFD_STAGE_BFS);
                PFD_CUDA(h, cudaLaunchCooperativeKernel(kern, dim3((unsigned)grid), dim3(BFS_THREADS), args, 0, h->stream));
            }
            h->launches++;
            PFD_CUDA(h, cudaMemcpyAsync(&st, h->bfs_state.p, sizeof(st), cudaMemcpyDeviceToHost, h->stream));
            PFD_CUDA(h, cudaStreamSynchronize(h->stream));
            if (st.stop == 1) break;
            if (st.stop != 2) return pfd_fail(h, PFD_ERR_CUDA, "bfs kernel ended in an unknown state");
            // grow the level table and continue from the level the kernel stopped at
            const int64_t new_cap = std::min<int64_t>(n + 1, h->level_cap * 4);
            if (new_cap <= h->level_cap) return pfd_fail(h, PFD_ERR_CUDA, "bfs level table cannot grow");
            DevBuf bigger;
            PFD_TRY(pfd_reserve(h, bigger, (size_t)(new_cap + 1) * sizeof(long long)));
            PFD_CUDA(h, cudaMemcpyAsync(bigger.p, h->level_off.p, (size_t)(h->level_cap + 1) * sizeof(long long), cudaMemcpyDeviceToDevice, h->stream));
            PFD_CUDA(h, cudaStreamSynchronize(h->stream));
            pfd_release(h->level_off);
            h->level_off = bigger;
            h->level_cap = new_cap;
            unsigned int zero = 0;
            PFD_CUDA(h, cudaMemcpyAsync(&((BfsState*)h->bfs_state.p)->stop, &zero, sizeof(zero), cudaMemcpyHostToDevice, h->stream));
        }
        h->nnodes = (int64_t)st.total;
        h->nlevels = (int64_t)st.cur_level;
        h->h_level_off.resize((size_t)h->nlevels + 1);
        PFD_CUDA(h, cudaMemcpyAsync(h->h_level_off.data(), h->level_off.p, (size_t)(h->nlevels + 1) * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
        PFD_TRY(build_schedule(h));
        h->ordered = true;
        h->have_rank |= want_rank;
        h->have_basins |= want_basins;
    }
    if (want_rank && !h->have_rank) {
        PFD_TRY(pfd_reserve(h, h->rank, (size_t)n * sizeof(int32_t)));
        order_init_rank_kernel<<<grid_for(n, 256, 4, 1ll << 30), 256, 0, h->stream>>>((const uint8_t*)h->dir.p, n, (int32_t*)h->rank.p);
        PFD_LAUNCH_CHECK(h);
        RankOp op{(int32_t*)h->rank.p};
        PFD_TRY((run_sweep<RankOp, false>(h, op, 0)));
        h->have_rank = true;
    }
    if (want_basins && !h->have_basins) {
        PFD_TRY(pfd_reserve(h, h->basins, (size_t)n * sizeof(uint32_t)));
        PFD_CUDA(h, cudaMemsetAsync(h->basins.p, 0, (size_t)n * sizeof(uint32_t), h->stream));
        if (h->n_pits > 0) {
            // seq[0:n_pits] already holds the pits; only the labels are needed
            order_init_pits_kernel<<<grid_for(h->n_pits, 256), 256, 0, h->stream>>>(
                (const cell_t*)h->pits.p, h->n_pits, (cell_t*)h->seq.p, nullptr, nullptr, (uint32_t*)h->basins.p);
            PFD_LAUNCH_CHECK(h);
        }
        FillUpOp<uint32_t> op{(const uint8_t*)h->dir.p, (uint32_t*)h->basins.p, h->ncol};
        PFD_TRY((run_sweep<FillUpOp<uint32_t>, false>(h, op, 1)));
        h->have_basins = true;
    }
    return PFD_OK;
}

extern "C" int pfd_order(pfd_handle* h, int64_t* nnodes, int64_t* nlevels) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    PFD_TRY(order_impl(h, true, false));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    if (nnodes) *nnodes = h->nnodes;
    if (nlevels) *nlevels = h->nlevels;
    return PFD_OK;
}

// ---------------------------------------------------------------------------------------------------------
// fetch
// ---------------------------------------------------------------------------------------------------------
static int copy_cells_out(pfd_handle* h, const cell_t* src, int64_t count, void* out, int idx_dtype) {
    if (count == 0) return PFD_OK;
    const size_t osz = pfd_dtype_size(idx_dtype);
    if (idx_dtype == PFD_I32 || idx_dtype == PFD_U32) {
        PFD_CUDA(h, cudaMemcpyAsync(out, src, (size_t)count * 4, cudaMemcpyDefault, h->stream));
        return PFD_OK;
    }
    if (idx_dtype != PFD_I64 && idx_dtype != PFD_U64) return pfd_fail(h, PFD_ERR_INVALID_ARG, "index dtype must be 32/64-bit integer");
    void* dev = nullptr;
    PFD_TRY(pfd_stage_out(h, out, (size_t)count * osz, 2, &dev));
    widen_cells_kernel<int64_t><<<grid_for(count, 256, 4), 256, 0, h->stream>>>(src, count, (int64_t*)dev);
    PFD_LAUNCH_CHECK(h);
    PFD_TRY(pfd_finish_out(h, out, dev, (size_t)count * osz));
    return PFD_OK;
}

extern "C" int pfd_fetch(pfd_handle* h, int which, void* out, int idx_dtype) {
    PFD_TRY(check_handle(h));
    if (!h->parsed) return pfd_fail(h, PFD_ERR_STATE, "pfd_fetch: no raster parsed on this handle");
    if (!out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_fetch: out is null");
    if (h->tiled && which != PFD_ARR_PITS && which != PFD_ARR_PIT_IS_OUTLET)
        return pfd_fail(h, PFD_ERR_UNSUPPORTED, "this handle holds a row block: only the pfd_tiled_* entry points apply");
    stage_reset(h);
    const int64_t n = h->n;
    switch (which) {
    case PFD_ARR_IDXS_DS: {
        const size_t osz = pfd_dtype_size(idx_dtype);
        if (idx_dtype != PFD_I32 && idx_dtype != PFD_U32 && idx_dtype != PFD_I64 && idx_dtype != PFD_U64)
            return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_fetch: bad index dtype");
        void* dev = nullptr;
        PFD_TRY(pfd_stage_out(h, out, (size_t)n * osz, 2, &dev));
        const int g = grid_for(n, 256, 4);
        if (osz == 4) dir_to_idxs_kernel<uint32_t><<<g, 256, 0, h->stream>>>((const uint8_t*)h->dir.p, n, h->ncol, (uint32_t*)dev);
        else dir_to_idxs_kernel<int64_t><<<g, 256, 0, h->stream>>>((const uint8_t*)h->dir.p, n, h->ncol, (int64_t*)dev);
        PFD_LAUNCH_CHECK(h);
        PFD_TRY(pfd_finish_out(h, out, dev, (size_t)n * osz));
        break;
    }
    case PFD_ARR_PITS:
        PFD_TRY(copy_cells_out(h, (const cell_t*)h->pits.p, h->n_pits, out, idx_dtype));
        break;
    case PFD_ARR_PIT_IS_OUTLET:
        if (h->n_pits) PFD_CUDA(h, cudaMemcpyAsync(out, h->pit_outlet.p, (size_t)h->n_pits, cudaMemcpyDefault, h->stream));
        break;
    case PFD_ARR_SEQ:
        PFD_TRY(order_impl(h, false, false));
        PFD_TRY(copy_cells_out(h, (const cell_t*)h->seq.p, h->nnodes, out, idx_dtype));
        break;
    case PFD_ARR_RANK:
        if (tiles_usable(h)) PFD_TRY(tiles_ensure(h, true, false, false));
        else PFD_TRY(order_impl(h, true, false));
        PFD_CUDA(h, cudaMemcpyAsync(out, h->rank.p, (size_t)n * sizeof(int32_t), cudaMemcpyDefault, h->stream));
        break;
    case PFD_ARR_N_UPSTREAM: {
        void* dev = nullptr;
        PFD_TRY(pfd_stage_out(h, out, (size_t)n, 2, &dev));
        PFD_TRY(ensure_upmask(h));
        upstream_count_kernel<<<grid_for(n, 256, 4), 256, 0, h->stream>>>((const uint8_t*)h->dir.p, (const uint8_t*)h->upmask.p, n, (int8_t*)dev);
        PFD_LAUNCH_CHECK(h);
        PFD_TRY(pfd_finish_out(h, out, dev, (size_t)n));
        break;
    }
    case PFD_ARR_D8:
    case PFD_ARR_LDD: {
        void* dev = nullptr;
        PFD_TRY(pfd_stage_out(h, out, (size_t)n, 2, &dev));
        if (which == PFD_ARR_D8) dir_to_codes_kernel<0><<<grid_for(n, 256, 4), 256, 0, h->stream>>>((const uint8_t*)h->dir.p, n, (uint8_t*)dev);
        else dir_to_codes_kernel<1><<<grid_for(n, 256, 4), 256, 0, h->stream>>>((const uint8_t*)h->dir.p, n, (uint8_t*)dev);
        PFD_LAUNCH_CHECK(h);
        PFD_TRY(pfd_finish_out(h, out, dev, (size_t)n));
        break;
    }
    case PFD_ARR_NEXTXY: {
        void* dev = nullptr;
        PFD_TRY(pfd_stage_out(h, out, (size_t)n * 8, 2, &dev));
        dir_to_nextxy_kernel<<<grid_for(n, 256, 4), 256, 0, h->stream>>>((const uint8_t*)h->dir.p, n, h->ncol, (int32_t*)dev, (int32_t*)dev + n);
        PFD_LAUNCH_CHECK(h);
        PFD_TRY(pfd_finish_out(h, out, dev, (size_t)n * 8));
        break;
    }
    case PFD_ARR_SUBBASIN_OUTLETS:
        PFD_TRY(copy_cells_out(h, (const cell_t*)h->sub_idxs.p, h->n_sub, out, idx_dtype));
        break;
    case PFD_ARR_STREAM_OFFSETS:
        if (h->n_streams < 0) return pfd_fail(h, PFD_ERR_STATE, "pfd_fetch: no pfd_streams result on this handle");
        PFD_CUDA(h, cudaMemcpyAsync(out, h->stream_off.p, (size_t)(h->n_streams + 1) * sizeof(long long), cudaMemcpyDefault, h->stream));
        break;
    case PFD_ARR_STREAM_CELLS:
        if (h->n_streams < 0) return pfd_fail(h, PFD_ERR_STATE, "pfd_fetch: no pfd_streams result on this handle");
        PFD_TRY(copy_cells_out(h, (const cell_t*)h->stream_cells.p, h->n_stream_cells, out, idx_dtype));
        break;
    case PFD_ARR_REGION_LABELS:
        if (!h->have_sub_labels) return pfd_fail(h, PFD_ERR_STATE, "pfd_fetch: no pfd_region_outlets / pfd_region_slices result on this handle");
        if (h->n_sub) PFD_CUDA(h, cudaMemcpyAsync(out, h->sub_labels.p, (size_t)h->n_sub * sizeof(int64_t), cudaMemcpyDefault, h->stream));
        break;
    case PFD_ARR_REGION_SLICES:
        if (!h->have_sub_slices) return pfd_fail(h, PFD_ERR_STATE, "pfd_fetch: no pfd_region_slices result on this handle");
        if (h->n_sub) PFD_CUDA(h, cudaMemcpyAsync(out, h->sub_slices.p, (size_t)h->n_sub * sizeof(int4), cudaMemcpyDefault, h->stream));
        break;
    case PFD_ARR_LEVEL_OFFSETS:
        PFD_TRY(order_impl(h, false, false));
        PFD_CUDA(h, cudaMemcpyAsync(out, h->level_off.p, (size_t)(h->nlevels + 1) * sizeof(long long), cudaMemcpyDefault, h->stream));
        break;
    default:
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_fetch: unknown array id");
    }
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    return PFD_OK;
}

// ---------------------------------------------------------------------------------------------------------
// tile-dataflow sweeps (pfd_tilesweep.cuh): no cell ordering needed
// ---------------------------------------------------------------------------------------------------------
#ifndef TS_NT_OVERRIDE
#define TS_NT_OVERRIDE 0  // -DTS_NT_OVERRIDE=32|64|128|256: threads per tile visit for every tile sweep (tuning)
#endif
static bool use_tile_sweeps(const pfd_handle* h) { return h->tile_sweeps == 2 || (h->tile_sweeps == 1 && !h->ordered); }

static int ts_prepare(pfd_handle* h, TsArgs& A, const char* who, std::initializer_list<const void*> arrays) {
    if (!h->parsed) return pfd_fail(h, PFD_ERR_STATE, std::string(who) + ": no raster parsed on this handle");
    if (h->tiled) return pfd_fail(h, PFD_ERR_UNSUPPORTED, "this handle holds a row block: only the pfd_tiled_* entry points apply");
    const long long ntx = (h->ncol + TS_T - 1) / TS_T, nty = (h->nrow + TS_T - 1) / TS_T;
    const long long ntiles = ntx * nty;
    if (ntiles >= (1ll << 31)) return pfd_fail(h, PFD_ERR_UNSUPPORTED, std::string(who) + ": too many tiles");
    PFD_TRY(pfd_reserve(h, h->ts_done, (size_t)ntiles * TS_BMW * sizeof(uint32_t)));
    PFD_TRY(pfd_reserve(h, h->ts_lists, (size_t)ntiles * 3 * sizeof(uint32_t) + sizeof(TsCtl)));
    // control block first (it holds a 64-bit counter), then the stamps, then the two work lists
    A.ctl = (TsCtl*)h->ts_lists.p;
    uint32_t* base = (uint32_t*)((char*)h->ts_lists.p + sizeof(TsCtl));
    A.dir = (const uint8_t*)h->dir.p + h->dir_off;
    A.nrow = h->nrow, A.ncol = h->ncol, A.ntx = (int)ntx, A.nty = (int)nty;
    A.own_lo = 0, A.own_hi = h->nrow, A.fdone = nullptr, A.first_pass = 1;
    A.max_passes = h->ts_max_passes;
    A.al16 = (h->ncol % 16 == 0) && ((uintptr_t)A.dir % 16 == 0);
    for (const void* q : arrays) A.al16 = A.al16 && ((uintptr_t)q % 16 == 0);
    A.done = (uint32_t*)h->ts_done.p;
    A.stamp = base, A.list[0] = base + ntiles, A.list[1] = base + 2 * ntiles;
    PFD_CUDA(h, cudaMemsetAsync(h->ts_lists.p, 0, sizeof(TsCtl) + (size_t)ntiles * sizeof(uint32_t), h->stream));
    return PFD_OK;
}

static int ts_launch(pfd_handle* h, void* kern, int nt, size_t smem, const TsArgs& A, void* op_ptr) {
    PFD_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    PFD_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nt, smem));
    if (per_sm < 1) return pfd_fail(h, PFD_ERR_CUDA, "tile sweep kernel cannot be made resident");
    const long long ntiles = (long long)A.ntx * A.nty;
    const long long grid = std::max<long long>(1, std::min<long long>((long long)per_sm * h->num_sms, ntiles));
    void* args[] = {(void*)&A, op_ptr};
    StageTimer t(h, PFD_STAGE_SWEEP);
    PFD_CUDA(h, cudaLaunchCooperativeKernel(kern, dim3((unsigned)grid), dim3((unsigned)nt), args, smem, h->stream));
    h->launches++;
    return PFD_OK;
}

static int ts_result(pfd_handle* h, const TsArgs& A, unsigned long long* resolved) {
    TsCtl c;
    PFD_CUDA(h, cudaMemcpyAsync(&c, A.ctl, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    h->sweep_passes = (int)c.passes;
    h->sweep_visits = (int64_t)c.visits;
    *resolved = c.resolved;
    return PFD_OK;
}


// Up-sweep over the whole raster; `op.out` receives the result. Cells that drain to no pit keep op.init().
template <class Op>
static int run_tile_up(pfd_handle* h, Op op, const char* who) {
    TsArgs A;
    PFD_TRY(ts_prepare(h, A, who, {op.out, op.init_src(), op.aux_src()}));
    // threads per tile visit (measured on B200 at 8192^2, profiles/r2_sweeps.md: 32 / 64 / 128 / 256 threads = 5.2 / 3.7 / 3.2 /
    // 4.0 ms for uint8 values, 7.5 / 5.1 / 4.3 / 4.2 ms for float32): warps in flight per SM matter more than tiles in flight
    constexpr int NT = TS_NT_OVERRIDE ? TS_NT_OVERRIDE : (sizeof(typename Op::V) == 1 ? 128 : 256);
    PFD_TRY(ts_launch(h, (void*)tile_up_sweep_kernel<NT, Op>, NT, sizeof(TsShared<typename Op::V, Op::AUX>), A, (void*)&op));
    unsigned long long resolved = 0;
    PFD_TRY(ts_result(h, A, &resolved));
    if ((int64_t)resolved != h->n_valid && h->ts_max_passes == 0) {
        // loops: the trees hanging on them were resolved by the dataflow but are outside the reference's `seq`
        PFD_TRY(tiles_usable(h) ? tiles_ensure(h, true, false, false) : order_impl(h, true, false));
        ts_reset_unranked_kernel<Op><<<grid_for(h->n, 256, 4, 148 * 32), 256, 0, h->stream>>>(
            (const uint8_t*)h->dir.p + h->dir_off, (const int32_t*)h->rank.p, h->n, op);
        PFD_LAUNCH_CHECK(h);
    }
    return PFD_OK;
}

template <class Op>
static int run_tile_down(pfd_handle* h, Op op, const char* who) {
    TsArgs A;
    PFD_TRY(ts_prepare(h, A, who, {op.out}));
    constexpr int NT = TS_NT_OVERRIDE ? TS_NT_OVERRIDE : 256;  // (HAND at 8192^2: 16.4 / 7.9 / 8.0 / 7.2 ms for 32 / 64 / 128 / 256)
    PFD_TRY(ts_launch(h, (void*)tile_down_sweep_kernel<NT, Op>, NT, sizeof(TsShared<typename Op::V, false>), A, (void*)&op));
    unsigned long long resolved = 0;
    return ts_result(h, A, &resolved);
}

// ---------------------------------------------------------------------------------------------------------
// Row-block (multi-GPU) tile sweeps: Strahler, accuflux (any dtype), HAND across the row blocks of ONE raster.
// These outputs are order-sensitive / not re-associable, so the single all-reduce of the integer path does not apply
// (SURVEY.md section 8e): every rank sweeps its block extended by the neighbours' edge rows ("foreign" rows: never
// resolved here), then the ranks swap the values + done flags of their edge rows and resume from the edge tiles, until
// a round resolves nothing anywhere. The number of rounds is the largest number of block-boundary crossings of a
// dependency chain + 1. Bit-identical to the single-GPU sweep: the same per-cell statement on the same final values.
// Step functions (the exchange is NCCL send/recv in pfd_sweep_tiled, or emulated by the caller in tests):
//   pfd_sweep_tiled_begin -> { pfd_sweep_tiled_round, pfd_sweep_tiled_edges (pack), swap, pfd_sweep_tiled_halo (unpack) }* -> pfd_sweep_tiled_end
// ---------------------------------------------------------------------------------------------------------
static size_t sw_edge_bytes(const pfd_handle* h) { return (size_t)h->ncol * (2 + h->sw.vsz + h->sw.asz); }

// edge record of one row: [dir ncol][done ncol][value ncol * vsz][aux ncol * asz]
__global__ void sw_pack_kernel(const uint8_t* __restrict__ dir, const uint32_t* __restrict__ done, const uint8_t* __restrict__ out,
                               const uint8_t* __restrict__ aux, long long row, long long ncol, int ntx, int vsz, int asz,
                               uint8_t* __restrict__ buf) {
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < ncol; c += (long long)gridDim.x * blockDim.x) {
        buf[c] = dir[row * ncol + c];
        const long long t = (row >> 6) * ntx + (c >> 6);
        buf[ncol + c] = (uint8_t)((done[t * TS_BMW + (row & 63) * 2 + ((c & 63) >> 5)] >> (c & 31)) & 1u);
        for (int b = 0; b < vsz; ++b) buf[2 * ncol + c * vsz + b] = out[(row * ncol + c) * vsz + b];
        for (int b = 0; b < asz; ++b) buf[(2 + vsz) * ncol + c * asz + b] = aux[(row * ncol + c) * asz + b];
    }
}

__global__ void sw_unpack_kernel(const uint8_t* __restrict__ buf, long long row, long long ncol, int vsz, int asz,
                                 uint8_t* __restrict__ dir, uint8_t* __restrict__ fdone, uint8_t* __restrict__ out,
                                 uint8_t* __restrict__ aux) {
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < ncol; c += (long long)gridDim.x * blockDim.x) {
        dir[row * ncol + c] = buf[c];
        fdone[c] = buf[ncol + c];
        for (int b = 0; b < vsz; ++b) out[(row * ncol + c) * vsz + b] = buf[2 * ncol + c * vsz + b];
        for (int b = 0; b < asz; ++b) aux[(row * ncol + c) * asz + b] = buf[(2 + vsz) * ncol + c * asz + b];
    }
}

static int sw_args(pfd_handle* h, TsArgs& A) {
    const long long nrow_ext = h->nrow + 2, ncol = h->ncol;
    const long long ntx = (ncol + TS_T - 1) / TS_T, nty = (nrow_ext + TS_T - 1) / TS_T;
    const long long ntiles = ntx * nty;
    if (ntiles >= (1ll << 31)) return pfd_fail(h, PFD_ERR_UNSUPPORTED, "pfd_sweep_tiled: too many tiles");
    PFD_TRY(pfd_reserve(h, h->ts_done, (size_t)ntiles * TS_BMW * sizeof(uint32_t)));
    PFD_TRY(pfd_reserve(h, h->ts_lists, (size_t)ntiles * 3 * sizeof(uint32_t) + sizeof(TsCtl)));
    A.ctl = (TsCtl*)h->ts_lists.p;
    uint32_t* base = (uint32_t*)((char*)h->ts_lists.p + sizeof(TsCtl));
    A.dir = (const uint8_t*)h->sw_dir.p;
    A.nrow = nrow_ext, A.ncol = ncol, A.ntx = (int)ntx, A.nty = (int)nty;
    A.own_lo = 1, A.own_hi = h->nrow + 1, A.fdone = (const uint8_t*)h->sw_fdone.p, A.first_pass = h->sw.next_pass;
    A.max_passes = 0;
    A.al16 = (ncol % 16 == 0) && ((uintptr_t)h->sw.data % 16 == 0) && ((uintptr_t)h->sw.drain % 16 == 0);
    A.done = (uint32_t*)h->ts_done.p;
    A.stamp = base, A.list[0] = base + ntiles, A.list[1] = base + 2 * ntiles;
    return PFD_OK;
}

extern "C" int pfd_sweep_tiled_begin(pfd_handle* h, int kind, const void* data, int dtype, const uint8_t* drain, double nodata_f,
                                     int64_t nodata_i, int nodata_is_int) {
    PFD_TRY(check_handle(h));
    if (!h->parsed) return pfd_fail(h, PFD_ERR_STATE, "pfd_sweep_tiled_begin: parse a row block first (pfd_tiled_parse / pfd_d8_flow_all_tiled)");
    if (kind < 0 || kind > 2) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_sweep_tiled_begin: kind must be 0 (strahler), 1 (accuflux), 2 (hand)");
    const size_t esz = pfd_dtype_size(dtype);
    if (kind == 1 && (!data || !esz)) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_sweep_tiled_begin: accuflux needs data of a known dtype");
    if (kind == 2 && (!data || !drain || (dtype != PFD_F32 && dtype != PFD_F64)))
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_sweep_tiled_begin: hand needs drain and float32 / float64 elevtn");
    auto& w = h->sw;
    w = pfd_handle::SweepShard();
    w.kind = kind, w.dtype = dtype;
    w.vsz = kind == 0 ? 1 : kind == 1 ? (int)esz : 8;
    w.asz = kind == 2 ? (int)esz : 0;
    w.nodata_f = nodata_f, w.nodata_i = nodata_i, w.nodata_is_int = nodata_is_int;
    const int64_t n = h->n, ncol = h->ncol, next = (h->nrow + 2) * ncol;
    PFD_TRY(pfd_reserve(h, h->sw_dir, (size_t)next));
    PFD_TRY(pfd_reserve(h, h->sw_out, (size_t)next * w.vsz));
    PFD_TRY(pfd_reserve(h, h->sw_aux, (size_t)std::max<int64_t>(next * w.asz, 16)));
    PFD_TRY(pfd_reserve(h, h->sw_fdone, (size_t)(2 * ncol)));
    for (int k = 0; k < 4; ++k) PFD_TRY(pfd_reserve(h, h->sw_edge[k], sw_edge_bytes(h)));
    // extended direction raster: the neighbours' edge rows start as nodata (no neighbour / nothing received yet)
    PFD_CUDA(h, cudaMemsetAsync(h->sw_dir.p, 0xFF, (size_t)next, h->stream));
    PFD_CUDA(h, cudaMemcpyAsync((uint8_t*)h->sw_dir.p + ncol, (const uint8_t*)h->dir.p + h->dir_off, (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
    PFD_CUDA(h, cudaMemsetAsync(h->sw_fdone.p, 0, (size_t)(2 * ncol), h->stream));
    // per-cell inputs of the own rows, on the device; addressed as rows 1 .. nrow of the extended raster
    const void* ddev = nullptr;
    if (kind == 1) {
        PFD_TRY(pfd_stage_in(h, data, (size_t)n * esz, 4, &ddev));
        w.data = (const uint8_t*)ddev - (size_t)ncol * esz;
    } else if (kind == 2) {
        PFD_TRY(pfd_stage_in(h, data, (size_t)n * esz, 4, &ddev));
        PFD_CUDA(h, cudaMemsetAsync(h->sw_aux.p, 0, (size_t)next * esz, h->stream));
        PFD_CUDA(h, cudaMemcpyAsync((uint8_t*)h->sw_aux.p + (size_t)ncol * esz, ddev, (size_t)n * esz, cudaMemcpyDeviceToDevice, h->stream));
        const void* drdev = nullptr;
        PFD_TRY(pfd_stage_in(h, drain, (size_t)n, 5, &drdev));
        w.drain = (const uint8_t*)drdev - ncol;
    }
    TsArgs A;
    PFD_TRY(sw_args(h, A));
    const long long ntiles = (long long)A.ntx * A.nty;
    PFD_CUDA(h, cudaMemsetAsync(h->ts_lists.p, 0, sizeof(TsCtl) + (size_t)ntiles * sizeof(uint32_t), h->stream));
    PFD_CUDA(h, cudaMemsetAsync(h->ts_done.p, 0, (size_t)ntiles * TS_BMW * sizeof(uint32_t), h->stream));  // (the first swap packs it)
    w.active = true;
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    return PFD_OK;
}

template <typename T>
static int sw_launch_accu(pfd_handle* h, const TsArgs& A) {
    AccuUpTileOp<T> op{(const T*)h->sw.data, (T*)h->sw_out.p, NoData{h->sw.nodata_f, h->sw.nodata_i, h->sw.nodata_is_int}};
    constexpr int NT = TS_NT_OVERRIDE ? TS_NT_OVERRIDE : (sizeof(T) == 1 ? 128 : 256);
    return ts_launch(h, (void*)tile_up_sweep_kernel<NT, AccuUpTileOp<T>>, NT, sizeof(TsShared<T, false>), A, (void*)&op);
}

// one round: every pass the block can make on its own (first round: every tile; later: from the tiles at the block edges)
extern "C" int pfd_sweep_tiled_round(pfd_handle* h, int64_t* newly_resolved) {
    PFD_TRY(check_handle(h));
    auto& w = h->sw;
    if (!w.active) return pfd_fail(h, PFD_ERR_STATE, "pfd_sweep_tiled_round: call pfd_sweep_tiled_begin first");
    TsArgs A;
    PFD_TRY(sw_args(h, A));
    if (w.next_pass > 1) {  // resume: the tile rows that see a foreign row
        std::vector<uint32_t> tiles;
        std::vector<long long> rows = {0, (long long)A.nty - 1, (long long)A.nty - 2};
        std::sort(rows.begin(), rows.end());
        rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
        for (long long ty : rows)
            if (ty >= 0)
                for (int tx = 0; tx < A.ntx; ++tx) tiles.push_back((uint32_t)(ty * A.ntx + tx));
        const unsigned int cnt[4] = {0, 0, 0, 0};
        PFD_CUDA(h, cudaMemcpyAsync(A.ctl->count, cnt, sizeof(cnt), cudaMemcpyHostToDevice, h->stream));
        const unsigned int c = (unsigned int)tiles.size();
        PFD_CUDA(h, cudaMemcpyAsync(&A.ctl->count[w.next_pass & 3], &c, sizeof(c), cudaMemcpyHostToDevice, h->stream));
        PFD_CUDA(h, cudaMemcpyAsync(A.list[w.next_pass & 1], tiles.data(), tiles.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));  // `tiles` is a local vector
    }
    int rc;
    if (w.kind == 0) {
        StrahlerTileOp<false> op{nullptr, (uint8_t*)h->sw_out.p};
        constexpr int NT = TS_NT_OVERRIDE ? TS_NT_OVERRIDE : 128;
        rc = ts_launch(h, (void*)tile_up_sweep_kernel<NT, StrahlerTileOp<false>>, NT, sizeof(TsShared<uint8_t, false>), A, (void*)&op);
    } else if (w.kind == 1) {
        switch (w.dtype) {
        case PFD_I8: rc = sw_launch_accu<int8_t>(h, A); break;
        case PFD_U8: rc = sw_launch_accu<uint8_t>(h, A); break;
        case PFD_I16: rc = sw_launch_accu<int16_t>(h, A); break;
        case PFD_U16: rc = sw_launch_accu<uint16_t>(h, A); break;
        case PFD_I32: rc = sw_launch_accu<int32_t>(h, A); break;
        case PFD_U32: rc = sw_launch_accu<uint32_t>(h, A); break;
        case PFD_I64: rc = sw_launch_accu<int64_t>(h, A); break;
        case PFD_U64: rc = sw_launch_accu<uint64_t>(h, A); break;
        case PFD_F32: rc = sw_launch_accu<float>(h, A); break;
        default: rc = sw_launch_accu<double>(h, A); break;
        }
    } else {
        constexpr int NT = TS_NT_OVERRIDE ? TS_NT_OVERRIDE : 256;
        if (w.dtype == PFD_F32) {
            HandTileOp<float> op{w.drain, (const float*)h->sw_aux.p, (double*)h->sw_out.p};
            rc = ts_launch(h, (void*)tile_down_sweep_kernel<NT, HandTileOp<float>>, NT, sizeof(TsShared<double, false>), A, (void*)&op);
        } else {
            HandTileOp<double> op{w.drain, (const double*)h->sw_aux.p, (double*)h->sw_out.p};
            rc = ts_launch(h, (void*)tile_down_sweep_kernel<NT, HandTileOp<double>>, NT, sizeof(TsShared<double, false>), A, (void*)&op);
        }
    }
    PFD_TRY(rc);
    TsCtl c;
    PFD_CUDA(h, cudaMemcpyAsync(&c, A.ctl, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    if (newly_resolved) *newly_resolved = (int64_t)(c.resolved - w.resolved);
    w.resolved = c.resolved;
    w.next_pass = std::max<int>((int)c.passes + 1, w.next_pass + 1);
    h->sweep_passes = (int)c.passes;
    return PFD_OK;
}

// which: 0 = my first row (for the rank above), 1 = my last row (for the rank below). Returns the packed record (device).
extern "C" int pfd_sweep_tiled_edges(pfd_handle* h, int which, void** buf_dev, int64_t* nbytes) {
    PFD_TRY(check_handle(h));
    auto& w = h->sw;
    if (!w.active || (which & ~1)) return pfd_fail(h, PFD_ERR_STATE, "pfd_sweep_tiled_edges: bad state / argument");
    TsArgs A;
    PFD_TRY(sw_args(h, A));
    const long long row = which == 0 ? 1 : h->nrow;
    sw_pack_kernel<<<grid_for(h->ncol, 256), 256, 0, h->stream>>>((const uint8_t*)h->sw_dir.p, (const uint32_t*)h->ts_done.p,
                                                                 (const uint8_t*)h->sw_out.p, (const uint8_t*)h->sw_aux.p, row, h->ncol,
                                                                 A.ntx, w.vsz, w.asz, (uint8_t*)h->sw_edge[which].p);
    PFD_LAUNCH_CHECK(h);
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    if (buf_dev) *buf_dev = h->sw_edge[which].p;
    if (nbytes) *nbytes = (int64_t)sw_edge_bytes(h);
    return PFD_OK;
}

// which: 0 = the row above my block (the last row of the rank above), 1 = the row below. record: device or host, as packed
// by the neighbour's pfd_sweep_tiled_edges; NULL = take what was received into the handle's own buffer (NCCL path).
extern "C" int pfd_sweep_tiled_halo(pfd_handle* h, int which, const void* record) {
    PFD_TRY(check_handle(h));
    auto& w = h->sw;
    if (!w.active || (which & ~1)) return pfd_fail(h, PFD_ERR_STATE, "pfd_sweep_tiled_halo: bad state / argument");
    if (record) PFD_CUDA(h, cudaMemcpyAsync(h->sw_edge[2 + which].p, record, sw_edge_bytes(h), cudaMemcpyDefault, h->stream));
    const long long row = which == 0 ? 0 : h->nrow + 1;
    sw_unpack_kernel<<<grid_for(h->ncol, 256), 256, 0, h->stream>>>((const uint8_t*)h->sw_edge[2 + which].p, row, h->ncol, w.vsz, w.asz,
                                                                   (uint8_t*)h->sw_dir.p, (uint8_t*)h->sw_fdone.p + (which ? h->ncol : 0),
                                                                   (uint8_t*)h->sw_out.p, (uint8_t*)h->sw_aux.p);
    PFD_LAUNCH_CHECK(h);
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    return PFD_OK;
}

// own rows of the result -> out (host or device); *resolved = cells of this block that were resolved
extern "C" int pfd_sweep_tiled_end(pfd_handle* h, void* out, int64_t* resolved) {
    PFD_TRY(check_handle(h));
    auto& w = h->sw;
    if (!w.active) return pfd_fail(h, PFD_ERR_STATE, "pfd_sweep_tiled_end: call pfd_sweep_tiled_begin first");
    if (out)
        PFD_CUDA(h, cudaMemcpyAsync(out, (const uint8_t*)h->sw_out.p + (size_t)h->ncol * w.vsz, (size_t)h->n * w.vsz, cudaMemcpyDefault, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    if (resolved) *resolved = (int64_t)w.resolved;
    w.active = false;
    return PFD_OK;
}

// NCCL composition of the step functions above: neighbour send/recv of the packed edge rows, one all-reduce of the
// progress counter per round. kind / data / dtype / drain / nodata as in pfd_sweep_tiled_begin; out = own rows.
extern "C" int pfd_sweep_tiled(pfd_handle* h, int kind, const void* data, int dtype, const uint8_t* drain, double nodata_f,
                               int64_t nodata_i, int nodata_is_int, void* out, int64_t* rounds_out) {
    PFD_TRY(check_handle(h));
    const int nranks = h->nccl_comm ? h->mg_nranks : 1, rank = h->nccl_comm ? h->mg_rank : 0;
    int rc = pfd_sweep_tiled_begin(h, kind, data, dtype, drain, nodata_f, nodata_i, nodata_is_int);
    // error agreement before the first exchange: a rank that could not set up says so
    PFD_TRY(pfd_reserve(h, h->mg_counts, (size_t)(4 * std::max(nranks, 1) + 4) * sizeof(unsigned long long)));
    unsigned long long* dcount = (unsigned long long*)h->mg_counts.p;
    auto allreduce3 = [&](unsigned long long a, unsigned long long b, unsigned long long c, unsigned long long* res) -> int {
        const unsigned long long v[3] = {a, b, c};
        PFD_CUDA(h, cudaMemcpyAsync(dcount, v, sizeof(v), cudaMemcpyHostToDevice, h->stream));
        if (nranks > 1) PFD_NCCL(h, ncclAllReduce(dcount, dcount, 3, ncclUint64, ncclSum, (ncclComm_t)h->nccl_comm, h->stream));
        PFD_CUDA(h, cudaMemcpyAsync(res, dcount, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
        return PFD_OK;
    };
    unsigned long long res[3];
    {
        const int rc2 = allreduce3(rc != PFD_OK ? 1ull : 0ull, 0, 0, res);
        if (rc != PFD_OK) return rc;
        PFD_TRY(rc2);
        if (res[0]) return pfd_fail(h, PFD_ERR_NCCL, "pfd_sweep_tiled: another rank failed to set up the sweep");
    }
    auto exchange = [&]() -> int {
        if (nranks == 1) return PFD_OK;
        const size_t nb = sw_edge_bytes(h);
        if (rank > 0) PFD_TRY(pfd_sweep_tiled_edges(h, 0, nullptr, nullptr));
        if (rank < nranks - 1) PFD_TRY(pfd_sweep_tiled_edges(h, 1, nullptr, nullptr));
        PFD_NCCL(h, ncclGroupStart());
        if (rank > 0) {
            PFD_NCCL(h, ncclSend(h->sw_edge[0].p, nb, ncclUint8, rank - 1, (ncclComm_t)h->nccl_comm, h->stream));
            PFD_NCCL(h, ncclRecv(h->sw_edge[2].p, nb, ncclUint8, rank - 1, (ncclComm_t)h->nccl_comm, h->stream));
        }
        if (rank < nranks - 1) {
            PFD_NCCL(h, ncclSend(h->sw_edge[1].p, nb, ncclUint8, rank + 1, (ncclComm_t)h->nccl_comm, h->stream));
            PFD_NCCL(h, ncclRecv(h->sw_edge[3].p, nb, ncclUint8, rank + 1, (ncclComm_t)h->nccl_comm, h->stream));
        }
        PFD_NCCL(h, ncclGroupEnd());
        if (rank > 0) PFD_TRY(pfd_sweep_tiled_halo(h, 0, nullptr));
        if (rank < nranks - 1) PFD_TRY(pfd_sweep_tiled_halo(h, 1, nullptr));
        return PFD_OK;
    };
    // the neighbours' edge rows of directions (and elevations) must be in place before the first round
    rc = exchange();
    if (rc != PFD_OK) return tiled_abort(h, rc);
    int rounds = 0;
    for (;;) {
        int64_t newly = 0;
        rc = pfd_sweep_tiled_round(h, &newly);
        if (rc == PFD_OK) rc = exchange();
        if (rc != PFD_OK) return tiled_abort(h, rc);
        ++rounds;
        rc = allreduce3((unsigned long long)newly, 0, 0, res);
        if (rc != PFD_OK) return tiled_abort(h, rc);
        if (res[0] == 0) break;  // a round that resolved nothing anywhere
    }
    int64_t resolved = 0;
    PFD_TRY(pfd_sweep_tiled_end(h, out, &resolved));
    PFD_TRY(allreduce3((unsigned long long)resolved, (unsigned long long)h->n_valid, 0, res));
    if (rounds_out) *rounds_out = rounds;
    if (kind != 2 && res[0] != res[1])
        return pfd_fail(h, PFD_ERR_UNSUPPORTED, "pfd_sweep_tiled: the raster has cells that drain to no pit (loops); the row-block "
                                               "up-sweeps do not reset the trees hanging on them -- use the single-GPU call");
    return PFD_OK;
}

// ---------------------------------------------------------------------------------------------------------
// sweeps
// ---------------------------------------------------------------------------------------------------------
template <typename T>
static int accuflux_typed(pfd_handle* h, const void* data_dev, void* out_dev, const NoData& nd, int direction) {
    const int64_t n = h->n;
    if (direction == 0 && use_tile_sweeps(h)) {
        AccuUpTileOp<T> op{(const T*)data_dev, (T*)out_dev, nd};
        return run_tile_up(h, op, "pfd_accuflux");
    }
    if (data_dev != out_dev)
        PFD_CUDA(h, cudaMemcpyAsync(out_dev, data_dev, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, h->stream));
    if (direction == 0) {
        PFD_TRY(ensure_upmask(h));
        AccuUpOp<T> op{(const uint8_t*)h->upmask.p, (T*)out_dev, h->ncol, nd};
        return run_sweep<AccuUpOp<T>, true>(h, op, 0);
    }
    AccuDownOp<T> op{(const uint8_t*)h->dir.p, (T*)out_dev, h->ncol, nd};
    return run_sweep<AccuDownOp<T>, false>(h, op, 1);
}

extern "C" int pfd_accuflux(pfd_handle* h, const void* data, int dtype, double nodata_f, int64_t nodata_i,
                            int nodata_is_int, int direction, void* out) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    if (!data || !out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_accuflux: null array");
    if (direction != 0 && direction != 1) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_accuflux: direction must be 0 (up) or 1 (down)");
    const size_t esz = pfd_dtype_size(dtype);
    if (!esz) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_accuflux: unknown dtype");
    if (!(direction == 0 && use_tile_sweeps(h))) PFD_TRY(order_impl(h, false, false));
    else if (!h->parsed) return pfd_fail(h, PFD_ERR_STATE, "pfd_accuflux: no raster parsed on this handle");
    const size_t bytes = (size_t)h->n * esz;
    void* out_dev = nullptr;
    PFD_TRY(pfd_stage_out(h, out, bytes, 3, &out_dev));
    const void* data_dev = nullptr;
    if (pfd_is_device_ptr(data)) {
        data_dev = data;
    } else if (direction == 0 && use_tile_sweeps(h)) {  // the tile sweep reads `data` and writes `out` (never in place: cells
        PFD_TRY(pfd_stage_in(h, data, bytes, 4, &data_dev));  // outside `seq` are reset from the data afterwards)
    } else {  // host data goes straight into the output buffer (accu = data.copy())
        PFD_CUDA(h, cudaMemcpyAsync(out_dev, data, bytes, cudaMemcpyHostToDevice, h->stream));
        data_dev = out_dev;
    }
    NoData nd{nodata_f, (long long)nodata_i, nodata_is_int};
    int rc;
    switch (dtype) {
    case PFD_I8: rc = accuflux_typed<int8_t>(h, data_dev, out_dev, nd, direction); break;
    case PFD_U8: rc = accuflux_typed<uint8_t>(h, data_dev, out_dev, nd, direction); break;
    case PFD_I16: rc = accuflux_typed<int16_t>(h, data_dev, out_dev, nd, direction); break;
    case PFD_U16: rc = accuflux_typed<uint16_t>(h, data_dev, out_dev, nd, direction); break;
    case PFD_I32: rc = accuflux_typed<int32_t>(h, data_dev, out_dev, nd, direction); break;
    case PFD_U32: rc = accuflux_typed<uint32_t>(h, data_dev, out_dev, nd, direction); break;
    case PFD_I64: rc = accuflux_typed<int64_t>(h, data_dev, out_dev, nd, direction); break;
    case PFD_U64: rc = accuflux_typed<uint64_t>(h, data_dev, out_dev, nd, direction); break;
    case PFD_F32: rc = accuflux_typed<float>(h, data_dev, out_dev, nd, direction); break;
    default: rc = accuflux_typed<double>(h, data_dev, out_dev, nd, direction); break;
    }
    PFD_TRY(rc);
    PFD_TRY(pfd_finish_out(h, out, out_dev, bytes));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    return PFD_OK;
}

static int uparea_cells_device(pfd_handle* h, int32_t* out_dev) {
    const int64_t n = h->n;
    uparea_init_kernel<<<grid_for(n, 256, 4, 1ll << 30), 256, 0, h->stream>>>((const uint8_t*)h->dir.p, n, out_dev);
    PFD_LAUNCH_CHECK(h);
    NoData nd{-9999.0, -9999, 1};
    PFD_TRY(ensure_upmask(h));
    AccuUpOp<int32_t> op{(const uint8_t*)h->upmask.p, out_dev, h->ncol, nd};
    return run_sweep<AccuUpOp<int32_t>, true>(h, op, 0);
}

extern "C" int pfd_upstream_area_cells(pfd_handle* h, int32_t* out) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    if (!out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_upstream_area_cells: out is null");
    const size_t bytes = (size_t)h->n * sizeof(int32_t);
    void* out_dev = nullptr;
    PFD_TRY(pfd_stage_out(h, out, bytes, 3, &out_dev));
    if (tiles_usable(h)) {
        PFD_TRY(tiles_solve(h, nullptr, nullptr, (int32_t*)out_dev));
    } else {
        PFD_TRY(order_impl(h, false, false));
        PFD_TRY(uparea_cells_device(h, (int32_t*)out_dev));
    }
    PFD_TRY(pfd_finish_out(h, out, out_dev, bytes));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    return PFD_OK;
}

#include "pfd_hand.cuh"  // path summaries by pointer doubling: HAND, and the label / value filling of the two functions below

template <typename IDX, typename U>
static int basins_custom(pfd_handle* h, const void* idx_dev, const void* ids_dev, int64_t k, void* out_dev) {
    unsigned int* flag = reinterpret_cast<unsigned int*>((unsigned long long*)h->counters.p + 5);
    PFD_CUDA(h, cudaMemsetAsync(flag, 0, sizeof(unsigned int), h->stream));
    PFD_CUDA(h, cudaMemsetAsync(out_dev, 0, (size_t)h->n * sizeof(U), h->stream));
    if (k > 0) {
        scatter_ids_kernel<IDX, U><<<grid_for(k, 256), 256, 0, h->stream>>>((const IDX*)idx_dev, (const U*)ids_dev, 0, k, h->n, (U*)out_dev, flag);
        PFD_LAUNCH_CHECK(h);
    }
    unsigned int hflag = 0;
    PFD_CUDA(h, cudaMemcpyAsync(&hflag, flag, sizeof(hflag), cudaMemcpyDeviceToHost, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    if (hflag & 4u) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_basins: outlet index out of bounds");
    if (h->hand_pathsum && !h->tiled) return fill_up_paths<U>(h, (U*)out_dev, HasNonZero<U>{});  // no cell ordering needed
    PFD_TRY(order_impl(h, false, false));
    FillUpOp<U> op{(const uint8_t*)h->dir.p, (U*)out_dev, h->ncol};
    return run_sweep<FillUpOp<U>, false>(h, op, 0);
}

extern "C" int pfd_basins(pfd_handle* h, const void* outlets, int64_t n_outlets, int idx_dtype, const void* ids,
                          int ids_dtype, void* out) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    if (!out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_basins: out is null");
    if (!outlets) {
        if (tiles_usable(h)) PFD_TRY(tiles_ensure(h, false, true, false));
        else PFD_TRY(order_impl(h, false, true));
        PFD_CUDA(h, cudaMemcpyAsync(out, h->basins.p, (size_t)h->n * sizeof(uint32_t), cudaMemcpyDefault, h->stream));
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
        stage_collect(h);
        return PFD_OK;
    }
    if (!ids || n_outlets < 0) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_basins: ids is null");
    const size_t isz = pfd_dtype_size(idx_dtype), usz = pfd_dtype_size(ids_dtype);
    if ((isz != 4 && isz != 8) || idx_dtype == PFD_F32 || idx_dtype == PFD_F64)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_basins: outlet indices must be 32/64-bit integers");
    if (!usz || ids_dtype == PFD_F32 || ids_dtype == PFD_F64)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_basins: ids must be an integer dtype");
    if (!h->parsed) return pfd_fail(h, PFD_ERR_STATE, "pfd_basins: no raster parsed on this handle");
    const size_t bytes = (size_t)h->n * usz;
    void* out_dev = nullptr;
    PFD_TRY(pfd_stage_out(h, out, bytes, 3, &out_dev));
    const void *idx_dev = nullptr, *ids_dev = nullptr;
    PFD_TRY(pfd_stage_in(h, outlets, (size_t)std::max<int64_t>(n_outlets, 1) * isz, 4, &idx_dev));
    PFD_TRY(pfd_stage_in(h, ids, (size_t)std::max<int64_t>(n_outlets, 1) * usz, 5, &ids_dev));
    int rc;
    const bool i64 = (isz == 8);
    const bool isigned = (idx_dtype == PFD_I32 || idx_dtype == PFD_I64);
#define BAS(IDX)                                                                                          \
    (usz == 1 ? basins_custom<IDX, uint8_t>(h, idx_dev, ids_dev, n_outlets, out_dev)                      \
              : usz == 2 ? basins_custom<IDX, uint16_t>(h, idx_dev, ids_dev, n_outlets, out_dev)          \
                         : usz == 4 ? basins_custom<IDX, uint32_t>(h, idx_dev, ids_dev, n_outlets, out_dev) \
                                    : basins_custom<IDX, uint64_t>(h, idx_dev, ids_dev, n_outlets, out_dev))
    if (i64) rc = isigned ? BAS(int64_t) : BAS(uint64_t);
    else rc = isigned ? BAS(int32_t) : BAS(uint32_t);
#undef BAS
    PFD_TRY(rc);
    PFD_TRY(pfd_finish_out(h, out, out_dev, bytes));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    return PFD_OK;
}

extern "C" int pfd_strahler(pfd_handle* h, const uint8_t* mask, uint8_t* out) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    if (!out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_strahler: out is null");
    if (!h->parsed) return pfd_fail(h, PFD_ERR_STATE, "pfd_strahler: no raster parsed on this handle");
    const bool tile = use_tile_sweeps(h);
    if (!tile) PFD_TRY(order_impl(h, false, false));
    const size_t bytes = (size_t)h->n;
    void* out_dev = nullptr;
    PFD_TRY(pfd_stage_out(h, out, bytes, 3, &out_dev));
    const void* mask_dev = nullptr;
    if (mask) PFD_TRY(pfd_stage_in(h, mask, bytes, 4, &mask_dev));
    if (tile) {
        if (mask) {
            StrahlerTileOp<true> op{(const uint8_t*)mask_dev, (uint8_t*)out_dev};
            PFD_TRY(run_tile_up(h, op, "pfd_strahler"));
        } else {
            StrahlerTileOp<false> op{nullptr, (uint8_t*)out_dev};
            PFD_TRY(run_tile_up(h, op, "pfd_strahler"));
        }
    } else {
        PFD_CUDA(h, cudaMemsetAsync(out_dev, 0, bytes, h->stream));
        PFD_TRY(ensure_upmask(h));
        StrahlerOp op{(const uint8_t*)h->upmask.p, (const uint8_t*)mask_dev, (uint8_t*)out_dev, h->ncol};
        PFD_TRY((run_sweep<StrahlerOp, true>(h, op, 0)));
    }
    PFD_TRY(pfd_finish_out(h, out, out_dev, bytes));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    return PFD_OK;
}

__global__ void hand_reset_unranked_kernel(const uint8_t* __restrict__ dir, const int32_t* __restrict__ rank, int64_t n, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (dir[i] != PFD_DIR_NODATA && rank[i] < 0) out[i] = -9999.0;
}

extern "C" int pfd_hand(pfd_handle* h, const uint8_t* drain, const void* elevtn, int elev_dtype, double* out) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    if (!drain || !elevtn || !out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_hand: null array");
    if (elev_dtype != PFD_F32 && elev_dtype != PFD_F64) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_hand: elevtn must be float32 or float64");
    if (!h->parsed) return pfd_fail(h, PFD_ERR_STATE, "pfd_hand: no raster parsed on this handle");
    const bool tile = use_tile_sweeps(h);
    if (!tile && !(h->hand_pathsum && !h->tiled)) PFD_TRY(order_impl(h, false, false));
    const int64_t n = h->n;
    const size_t bytes = (size_t)n * sizeof(double);
    void* out_dev = nullptr;
    PFD_TRY(pfd_stage_out(h, out, bytes, 3, &out_dev));
    const void *drain_dev = nullptr, *elev_dev = nullptr;
    PFD_TRY(pfd_stage_in(h, drain, (size_t)n, 4, &drain_dev));
    PFD_TRY(pfd_stage_in(h, elevtn, (size_t)n * pfd_dtype_size(elev_dtype), 5, &elev_dev));
    bool pathsum_ran = false;
    if (h->hand_pathsum && !h->tiled) {
        // re-associated path sums (pfd_hand.cuh), accepted only when every cell satisfies the reference's statement bit for bit
        unsigned long long bad = 0;
        pathsum_ran = true;
        if (elev_dtype == PFD_F32) PFD_TRY(hand_pathsum<float>(h, (const uint8_t*)drain_dev, (const float*)elev_dev, (double*)out_dev, &bad));
        else PFD_TRY(hand_pathsum<double>(h, (const uint8_t*)drain_dev, (const double*)elev_dev, (double*)out_dev, &bad));
        if (bad == 0) {
            h->hand_engine = 1;
            PFD_TRY(pfd_finish_out(h, out, out_dev, bytes));
            PFD_CUDA(h, cudaStreamSynchronize(h->stream));
            stage_collect(h);
            return PFD_OK;
        }
    }
    h->hand_engine = tile ? 2 : 3;
    if (!tile) PFD_TRY(order_impl(h, false, false));  // (no-op when the ordering is cached)
    if (tile) {
        if (elev_dtype == PFD_F32) {
            HandTileOp<float> op{(const uint8_t*)drain_dev, (const float*)elev_dev, (double*)out_dev};
            PFD_TRY(run_tile_down(h, op, "pfd_hand"));
        } else {
            HandTileOp<double> op{(const uint8_t*)drain_dev, (const double*)elev_dev, (double*)out_dev};
            PFD_TRY(run_tile_down(h, op, "pfd_hand"));
        }
        // the dataflow starts at every drain cell, also at one that drains to no pit (above a loop): the reference only knows the
        // cells of its sequence. Which cells those are: from the path structure of the rejected attempt, else from the rank.
        if (pathsum_ran) {
            PFD_TRY(hand_mask_unreached(h, (double*)out_dev));
        } else {
            PFD_TRY(tiles_usable(h) ? tiles_ensure(h, true, false, false) : order_impl(h, true, false));
            hand_reset_unranked_kernel<<<grid_for(n, 256, 4, 148 * 32), 256, 0, h->stream>>>((const uint8_t*)h->dir.p + h->dir_off, (const int32_t*)h->rank.p, n,
                                                                                             (double*)out_dev);
            PFD_LAUNCH_CHECK(h);
        }
        PFD_TRY(pfd_finish_out(h, out, out_dev, bytes));
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
        stage_collect(h);
        return PFD_OK;
    }
    fill_kernel<double><<<grid_for(n, 256, 4), 256, 0, h->stream>>>((double*)out_dev, n, -9999.0);
    PFD_LAUNCH_CHECK(h);
    if (elev_dtype == PFD_F32) {
        HandOp<float> op{(const uint8_t*)h->dir.p, (const uint8_t*)drain_dev, (const float*)elev_dev, (double*)out_dev, h->ncol};
        PFD_TRY((run_sweep<HandOp<float>, false>(h, op, 0)));
    } else {
        HandOp<double> op{(const uint8_t*)h->dir.p, (const uint8_t*)drain_dev, (const double*)elev_dev, (double*)out_dev, h->ncol};
        PFD_TRY((run_sweep<HandOp<double>, false>(h, op, 0)));
    }
    PFD_TRY(pfd_finish_out(h, out, out_dev, bytes));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    return PFD_OK;
}


// ---------------------------------------------------------------------------------------------------------
// widening: fillnodata, main_upstream, masked upstream_count, classic stream order
// ---------------------------------------------------------------------------------------------------------
template <typename T>
static int fillnodata_typed(pfd_handle* h, void* out_dev, const NoData& nd, int direction, int how) {
    if (direction == 0) {
        if (h->hand_pathsum && !h->tiled) return fill_up_paths<T>(h, (T*)out_dev, HasData<T>{nd});  // no cell ordering needed
        PFD_TRY(order_impl(h, false, false));
        FillUpGenericOp<T> op{(const uint8_t*)h->dir.p, (T*)out_dev, h->ncol, nd};
        return run_sweep<FillUpGenericOp<T>, false>(h, op, 1);
    }
    if (how == 0) {
        PFD_TRY(ensure_upmask(h));
        FillDownOp<T, 0> op{(const uint8_t*)h->upmask.p, (T*)out_dev, h->ncol, nd};
        return run_sweep<FillDownOp<T, 0>, true>(h, op, 0);
    }
    if (how == 1) {
        PFD_TRY(ensure_upmask(h));
        FillDownOp<T, 1> op{(const uint8_t*)h->upmask.p, (T*)out_dev, h->ncol, nd};
        return run_sweep<FillDownOp<T, 1>, true>(h, op, 0);
    }
    PFD_TRY(ensure_upmask(h));
    FillDownOp<T, 2> op{(const uint8_t*)h->upmask.p, (T*)out_dev, h->ncol, nd};
    return run_sweep<FillDownOp<T, 2>, true>(h, op, 0);
}

extern "C" int pfd_fillnodata(pfd_handle* h, const void* data, int dtype, double nodata_f, int64_t nodata_i,
                              int nodata_is_int, int direction, int how, void* out) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    if (!data || !out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_fillnodata: null array");
    if ((direction != 0 && direction != 1) || how < 0 || how > 2)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_fillnodata: direction must be 0/1 and how 0 (max) / 1 (min) / 2 (sum)");
    const size_t esz = pfd_dtype_size(dtype);
    if (!esz) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_fillnodata: unknown dtype");
    if (!h->parsed) return pfd_fail(h, PFD_ERR_STATE, "pfd_fillnodata: no raster parsed on this handle");
    if (direction != 0) PFD_TRY(order_impl(h, false, false));  // (direction 0 orders only when it has to fall back to the level replay)
    const size_t bytes = (size_t)h->n * esz;
    void* out_dev = nullptr;
    PFD_TRY(pfd_stage_out(h, out, bytes, 3, &out_dev));
    if (out_dev != data) PFD_CUDA(h, cudaMemcpyAsync(out_dev, data, bytes, cudaMemcpyDefault, h->stream));
    NoData nd{nodata_f, (long long)nodata_i, nodata_is_int};
    int rc;
    switch (dtype) {
    case PFD_I8: rc = fillnodata_typed<int8_t>(h, out_dev, nd, direction, how); break;
    case PFD_U8: rc = fillnodata_typed<uint8_t>(h, out_dev, nd, direction, how); break;
    case PFD_I16: rc = fillnodata_typed<int16_t>(h, out_dev, nd, direction, how); break;
    case PFD_U16: rc = fillnodata_typed<uint16_t>(h, out_dev, nd, direction, how); break;
    case PFD_I32: rc = fillnodata_typed<int32_t>(h, out_dev, nd, direction, how); break;
    case PFD_U32: rc = fillnodata_typed<uint32_t>(h, out_dev, nd, direction, how); break;
    case PFD_I64: rc = fillnodata_typed<int64_t>(h, out_dev, nd, direction, how); break;
    case PFD_U64: rc = fillnodata_typed<uint64_t>(h, out_dev, nd, direction, how); break;
    case PFD_F32: rc = fillnodata_typed<float>(h, out_dev, nd, direction, how); break;
    default: rc = fillnodata_typed<double>(h, out_dev, nd, direction, how); break;
    }
    PFD_TRY(rc);
    PFD_TRY(pfd_finish_out(h, out, out_dev, bytes));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    return PFD_OK;
}

template <typename T>
static int main_upstream_typed(pfd_handle* h, const void* up_dev, double upa_min, void* out_dev, int idx_dtype) {
    const int g = grid_for(h->n, 256, 4);
    const T mn = (T)upa_min;
    PFD_TRY(ensure_upmask(h));
    if (pfd_dtype_size(idx_dtype) == 4)
        main_upstream_kernel<T, uint32_t><<<g, 256, 0, h->stream>>>((const uint8_t*)h->upmask.p, (const T*)up_dev, h->n, h->ncol, mn, (uint32_t*)out_dev);
    else
        main_upstream_kernel<T, int64_t><<<g, 256, 0, h->stream>>>((const uint8_t*)h->upmask.p, (const T*)up_dev, h->n, h->ncol, mn, (int64_t*)out_dev);
    PFD_LAUNCH_CHECK(h);
    return PFD_OK;
}

extern "C" int pfd_main_upstream(pfd_handle* h, const void* uparea, int dtype, double upa_min, void* out, int idx_dtype) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    if (!h->parsed || h->tiled) return pfd_fail(h, PFD_ERR_STATE, "pfd_main_upstream: no (whole) raster parsed on this handle");
    if (!uparea || !out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_main_upstream: null array");
    const size_t isz = pfd_dtype_size(idx_dtype), esz = pfd_dtype_size(dtype);
    if ((isz != 4 && isz != 8) || idx_dtype == PFD_F32 || idx_dtype == PFD_F64)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_main_upstream: index dtype must be a 32/64-bit integer");
    const void* up_dev = nullptr;
    PFD_TRY(pfd_stage_in(h, uparea, (size_t)h->n * esz, 4, &up_dev));
    void* out_dev = nullptr;
    PFD_TRY(pfd_stage_out(h, out, (size_t)h->n * isz, 3, &out_dev));
    int rc;
    switch (dtype) {
    case PFD_I32: rc = main_upstream_typed<int32_t>(h, up_dev, upa_min, out_dev, idx_dtype); break;
    case PFD_U32: rc = main_upstream_typed<uint32_t>(h, up_dev, upa_min, out_dev, idx_dtype); break;
    case PFD_I64: rc = main_upstream_typed<int64_t>(h, up_dev, upa_min, out_dev, idx_dtype); break;
    case PFD_F32: rc = main_upstream_typed<float>(h, up_dev, upa_min, out_dev, idx_dtype); break;
    case PFD_F64: rc = main_upstream_typed<double>(h, up_dev, upa_min, out_dev, idx_dtype); break;
    default: return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_main_upstream: uparea must be int32/uint32/int64/float32/float64");
    }
    PFD_TRY(rc);
    PFD_TRY(pfd_finish_out(h, out, out_dev, (size_t)h->n * isz));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    return PFD_OK;
}

// core.upstream_matrix (core.py:67-84): row i = the upstream cells of i in ascending linear index, padded with mv
__global__ void max_indegree_kernel(const uint8_t* __restrict__ upmask, int64_t n, unsigned int* __restrict__ out) {
    unsigned int m = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = max(m, (unsigned int)__popc((unsigned int)upmask[i]));
    m = __reduce_max_sync(0xFFFFFFFFu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

template <typename IDX>
__global__ void upstream_matrix_kernel(const uint8_t* __restrict__ upmask, int64_t n, int64_t ncol, int d, IDX* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t m = upmask[i];
        for (int j = 0; j < d; ++j) {
            IDX v = (IDX)-1;
            if (m) {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                v = (IDX)(i + pfd_slot_off(k, ncol));
            }
            out[i * d + j] = v;
        }
    }
}

extern "C" int pfd_upstream_matrix(pfd_handle* h, void* out, int idx_dtype, int64_t d_capacity, int64_t* d_out) {
    PFD_TRY(check_handle(h));
    if (!h->parsed || h->tiled) return pfd_fail(h, PFD_ERR_STATE, "pfd_upstream_matrix: no (whole) raster parsed on this handle");
    if (!d_out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_upstream_matrix: d_out is null");
    PFD_TRY(ensure_upmask(h));
    PFD_TRY(pfd_reserve(h, h->counters, 8 * sizeof(unsigned long long)));
    unsigned int* ctr = reinterpret_cast<unsigned int*>((unsigned long long*)h->counters.p + 6);
    PFD_CUDA(h, cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), h->stream));
    max_indegree_kernel<<<grid_for(h->n, 256, 8, 148 * 8), 256, 0, h->stream>>>((const uint8_t*)h->upmask.p, h->n, ctr);
    PFD_LAUNCH_CHECK(h);
    unsigned int d = 0;
    PFD_CUDA(h, cudaMemcpyAsync(&d, ctr, sizeof(d), cudaMemcpyDeviceToHost, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    *d_out = (int64_t)d;
    if (!out) return PFD_OK;  // size query
    if (d_capacity < (int64_t)d) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_upstream_matrix: out holds fewer columns than the largest in-degree");
    const size_t isz = pfd_dtype_size(idx_dtype);
    if ((isz != 4 && isz != 8) || idx_dtype == PFD_F32 || idx_dtype == PFD_F64)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_upstream_matrix: index dtype must be a 32/64-bit integer");
    if (d == 0) return PFD_OK;
    const size_t bytes = (size_t)h->n * d * isz;
    void* dev = nullptr;
    PFD_TRY(pfd_stage_out(h, out, bytes, 3, &dev));
    const int g = grid_for(h->n, 256, 2);
    if (isz == 4) upstream_matrix_kernel<uint32_t><<<g, 256, 0, h->stream>>>((const uint8_t*)h->upmask.p, h->n, h->ncol, (int)d, (uint32_t*)dev);
    else upstream_matrix_kernel<int64_t><<<g, 256, 0, h->stream>>>((const uint8_t*)h->upmask.p, h->n, h->ncol, (int)d, (int64_t*)dev);
    PFD_LAUNCH_CHECK(h);
    PFD_TRY(pfd_finish_out(h, out, dev, bytes));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    return PFD_OK;
}

extern "C" int pfd_upstream_count(pfd_handle* h, const uint8_t* mask, int8_t* out) {
    PFD_TRY(check_handle(h));
    if (!h->parsed || h->tiled) return pfd_fail(h, PFD_ERR_STATE, "pfd_upstream_count: no (whole) raster parsed on this handle");
    if (!out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_upstream_count: out is null");
    const int64_t n = h->n;
    void* dev = nullptr;
    PFD_TRY(pfd_stage_out(h, out, (size_t)n, 3, &dev));
    if (mask) {
        const void* mdev = nullptr;
        PFD_TRY(pfd_stage_in(h, mask, (size_t)n, 4, &mdev));
        PFD_TRY(ensure_upmask(h));
        upstream_count_mask_kernel<<<grid_for(n, 256, 4), 256, 0, h->stream>>>((const uint8_t*)h->dir.p, (const uint8_t*)h->upmask.p,
                                                                              (const uint8_t*)mdev, n, h->ncol, (int8_t*)dev);
    } else {
        PFD_TRY(ensure_upmask(h));
        upstream_count_kernel<<<grid_for(n, 256, 4), 256, 0, h->stream>>>((const uint8_t*)h->dir.p, (const uint8_t*)h->upmask.p, n, (int8_t*)dev);
    }
    PFD_LAUNCH_CHECK(h);
    PFD_TRY(pfd_finish_out(h, out, dev, (size_t)n));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    return PFD_OK;
}

extern "C" int pfd_stream_order_classic(pfd_handle* h, const void* idxs_us_main, int idx_dtype, const uint8_t* mask, uint8_t* out) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    if (!idxs_us_main || !out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_stream_order_classic: null array");
    const size_t isz = pfd_dtype_size(idx_dtype);
    if ((isz != 4 && isz != 8) || idx_dtype == PFD_F32 || idx_dtype == PFD_F64)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_stream_order_classic: index dtype must be a 32/64-bit integer");
    PFD_TRY(order_impl(h, false, false));
    const size_t bytes = (size_t)h->n;
    void* out_dev = nullptr;
    PFD_TRY(pfd_stage_out(h, out, bytes, 3, &out_dev));
    const void *mask_dev = nullptr, *main_dev = nullptr;
    if (mask) PFD_TRY(pfd_stage_in(h, mask, bytes, 4, &mask_dev));
    PFD_TRY(pfd_stage_in(h, idxs_us_main, (size_t)h->n * isz, 5, &main_dev));
    PFD_CUDA(h, cudaMemsetAsync(out_dev, 0, bytes, h->stream));
    PFD_TRY(ensure_upmask(h));
    const uint8_t *dir = (const uint8_t*)h->dir.p, *upm = (const uint8_t*)h->upmask.p;
    int rc;
    if (idx_dtype == PFD_I32) {
        ClassicOrderOp<int32_t> op{dir, upm, (const uint8_t*)mask_dev, (const int32_t*)main_dev, (uint8_t*)out_dev, h->ncol};
        rc = run_sweep<ClassicOrderOp<int32_t>, false>(h, op, 0);
    } else if (idx_dtype == PFD_U32) {
        ClassicOrderOp<uint32_t> op{dir, upm, (const uint8_t*)mask_dev, (const uint32_t*)main_dev, (uint8_t*)out_dev, h->ncol};
        rc = run_sweep<ClassicOrderOp<uint32_t>, false>(h, op, 0);
    } else {
        ClassicOrderOp<int64_t> op{dir, upm, (const uint8_t*)mask_dev, (const int64_t*)main_dev, (uint8_t*)out_dev, h->ncol};
        rc = run_sweep<ClassicOrderOp<int64_t>, false>(h, op, 0);
    }
    PFD_TRY(rc);
    PFD_TRY(pfd_finish_out(h, out, out_dev, bytes));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    return PFD_OK;
}


// streams.stream_distance (pyflwdir/streams.py:272-315). hop_table: nrow*3*2 float32 (host or device) when
// real_length, else NULL; out: N float32 (real_length) or N int32, -9999 outside the sequence.
extern "C" int pfd_stream_distance(pfd_handle* h, const uint8_t* mask, int real_length, const float* hop_table, void* out) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    if (!out || (real_length && !hop_table)) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_stream_distance: null array");
    if (!h->parsed) return pfd_fail(h, PFD_ERR_STATE, "pfd_stream_distance: no raster parsed on this handle");
    const bool paths = !real_length && h->hand_pathsum && !h->tiled;  // integer hop counts: path summaries, no ordering
    if (!paths) PFD_TRY(order_impl(h, false, false));
    const int64_t n = h->n;
    const size_t bytes = (size_t)n * 4;
    void* out_dev = nullptr;
    PFD_TRY(pfd_stage_out(h, out, bytes, 3, &out_dev));
    const void *mask_dev = nullptr, *hop_dev = nullptr;
    if (mask) PFD_TRY(pfd_stage_in(h, mask, (size_t)n, 4, &mask_dev));
    if (paths) {
        PFD_TRY((hd_solve(h, HopSrc{(const uint8_t*)mask_dev}, HopOut{(int32_t*)out_dev})));
        PFD_TRY(pfd_finish_out(h, out, out_dev, bytes));
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
        stage_collect(h);
        return PFD_OK;
    }
    if (real_length) PFD_TRY(pfd_stage_in(h, hop_table, (size_t)h->nrow * 6 * sizeof(float), 5, &hop_dev));
    int rc;
    if (real_length) {
        fill_kernel<float><<<grid_for(n, 256, 4), 256, 0, h->stream>>>((float*)out_dev, n, -9999.0f);
        PFD_LAUNCH_CHECK(h);
        StreamDistOp<true> op{(const uint8_t*)h->dir.p, (const uint8_t*)mask_dev, (const float*)hop_dev, out_dev, h->ncol};
        rc = run_sweep<StreamDistOp<true>, false>(h, op, 0);
    } else {
        fill_kernel<int32_t><<<grid_for(n, 256, 4), 256, 0, h->stream>>>((int32_t*)out_dev, n, -9999);
        PFD_LAUNCH_CHECK(h);
        StreamDistOp<false> op{(const uint8_t*)h->dir.p, (const uint8_t*)mask_dev, nullptr, out_dev, h->ncol};
        rc = run_sweep<StreamDistOp<false>, false>(h, op, 0);
    }
    PFD_TRY(rc);
    PFD_TRY(pfd_finish_out(h, out, out_dev, bytes));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    return PFD_OK;
}


// dem.floodplains (pyflwdir/dem.py:333-379). drainh_init: N float32 = float32(uparea ** b) at drain cells, -9999
// elsewhere (host-evaluated); elevtn: N float32 / float64; out: N int8 (-1 outside the sequence, else 0 / 1).
extern "C" int pfd_floodplains(pfd_handle* h, const float* drainh_init, const void* elevtn, int elev_dtype, int8_t* out) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    if (!drainh_init || !elevtn || !out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_floodplains: null array");
    if (elev_dtype != PFD_F32 && elev_dtype != PFD_F64) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_floodplains: elevtn must be float32 or float64");
    PFD_TRY(order_impl(h, false, false));
    const int64_t n = h->n;
    void* out_dev = nullptr;
    PFD_TRY(pfd_stage_out(h, out, (size_t)n, 3, &out_dev));
    const void* elev_dev = nullptr;
    PFD_TRY(pfd_stage_in(h, elevtn, (size_t)n * pfd_dtype_size(elev_dtype), 5, &elev_dev));
    // working copies of drainh (modified in place) and drainz
    PFD_TRY(pfd_reserve(h, h->scratch[4], (size_t)n * sizeof(float)));
    PFD_TRY(pfd_reserve(h, h->scratch[2], (size_t)n * sizeof(float)));
    float* drainh = (float*)h->scratch[4].p;
    float* drainz = (float*)h->scratch[2].p;
    PFD_CUDA(h, cudaMemcpyAsync(drainh, drainh_init, (size_t)n * sizeof(float), cudaMemcpyDefault, h->stream));
    PFD_CUDA(h, cudaMemsetAsync(out_dev, 0xFF, (size_t)n, h->stream));  // -1
    int rc;
    if (elev_dtype == PFD_F32) {
        FloodplainOp<float> op{(const uint8_t*)h->dir.p, (const float*)elev_dev, drainh, drainz, (int8_t*)out_dev, h->ncol};
        rc = run_sweep<FloodplainOp<float>, false>(h, op, 0);
    } else {
        FloodplainOp<double> op{(const uint8_t*)h->dir.p, (const double*)elev_dev, drainh, drainz, (int8_t*)out_dev, h->ncol};
        rc = run_sweep<FloodplainOp<double>, false>(h, op, 0);
    }
    PFD_TRY(rc);
    PFD_TRY(pfd_finish_out(h, out, out_dev, (size_t)n));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    return PFD_OK;
}

// ---------------------------------------------------------------------------------------------------------
// fused headline pass
// ---------------------------------------------------------------------------------------------------------
static int flow_all_impl(pfd_handle* h, const uint8_t* d8, int64_t nrow, int64_t ncol, void* idxs_ds_out,
                         int idx_dtype, int32_t* rank_out, int32_t* uparea_out, uint32_t* basins_out,
                         int64_t* n_valid, int64_t* n_pits, int64_t* nnodes) {
    stage_reset(h);
    cudaEventRecord(h->ev_start[PFD_STAGE_TOTAL], h->stream);
    h->stage_used[PFD_STAGE_TOTAL] = true;
    // Fused path: everything device-resident (the D2H copy of a host idxs_ds is better overlapped with the solve,
    // which needs the separate parse pass) and the tile solver enabled.
    const bool all_dev = d8 && pfd_is_device_ptr(d8) && (!idxs_ds_out || pfd_is_device_ptr(idxs_ds_out)) &&
                         (!rank_out || pfd_is_device_ptr(rank_out)) && (!uparea_out || pfd_is_device_ptr(uparea_out)) &&
                         (!basins_out || pfd_is_device_ptr(basins_out));
    if (h->use_tiles && h->fuse_parse && all_dev) {
        PFD_TRY(check_shape(h, nrow, ncol, "pfd_d8_flow_all"));
        if (idxs_ds_out && idx_dtype != PFD_I32 && idx_dtype != PFD_U32 && idx_dtype != PFD_I64)
            return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_d8_flow_all: idx_dtype must be int32, uint32 or int64");
        if (idxs_ds_out && idx_dtype == PFD_I32 && nrow * ncol >= 2147483647ll)
            return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_d8_flow_all: int32 indices cannot address this raster");
        invalidate(h);
        void* rk = rank_out;
        bool tmp_rank = false;
        if (!rk && nnodes) {  // nnodes needs the rank
            PFD_TRY(pfd_reserve(h, h->rank, (size_t)nrow * ncol * 4));
            rk = h->rank.p;
            tmp_rank = true;
        }
        PFD_TRY(flow_all_fused(h, d8, nrow, ncol, idxs_ds_out, idx_dtype, (int32_t*)rk, basins_out, uparea_out));
        if (nnodes) PFD_TRY(count_ranked(h, (const int32_t*)rk, &h->nnodes));
        if (tmp_rank) h->have_rank = true;
        cudaEventRecord(h->ev_stop[PFD_STAGE_TOTAL], h->stream);
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
        stage_collect(h);
        if (n_valid) *n_valid = h->n_valid;
        if (n_pits) *n_pits = h->n_pits;
        if (nnodes) *nnodes = h->nnodes;
        return PFD_OK;
    }
    PFD_TRY(parse_impl(h, d8, nrow, ncol, idxs_ds_out, idx_dtype, /*overlap_idxs_copy=*/true));
    if (h->n_pits == 0) return pfd_fail(h, PFD_ERR_NO_PITS, "Invalid FlwdirRaster: no pits found");
    const size_t b4 = (size_t)h->n * 4;
    if (tiles_usable(h)) {
        // write straight into the caller's buffers (device) or into staging buffers (host)
        void *rk = nullptr, *bs = nullptr, *up = nullptr;
        if (rank_out) PFD_TRY(pfd_stage_out(h, rank_out, b4, 2, &rk));
        if (basins_out) PFD_TRY(pfd_stage_out(h, basins_out, b4, 4, &bs));
        if (uparea_out) PFD_TRY(pfd_stage_out(h, uparea_out, b4, 3, &up));
        bool tmp_rank = false;
        if (!rk && nnodes) {  // nnodes needs the rank
            PFD_TRY(pfd_reserve(h, h->rank, b4));
            rk = h->rank.p;
            tmp_rank = true;
        }
        PFD_TRY(tiles_solve(h, (int32_t*)rk, (uint32_t*)bs, (int32_t*)up));
        if (rank_out) PFD_TRY(pfd_finish_out(h, rank_out, rk, b4));
        if (basins_out) PFD_TRY(pfd_finish_out(h, basins_out, bs, b4));
        if (uparea_out) PFD_TRY(pfd_finish_out(h, uparea_out, up, b4));
        if (nnodes) PFD_TRY(count_ranked(h, (const int32_t*)rk, &h->nnodes));
        if (tmp_rank) h->have_rank = true;
    } else {
        PFD_TRY(order_impl(h, rank_out != nullptr, basins_out != nullptr));
        if (rank_out) PFD_CUDA(h, cudaMemcpyAsync(rank_out, h->rank.p, b4, cudaMemcpyDefault, h->stream));
        if (basins_out) PFD_CUDA(h, cudaMemcpyAsync(basins_out, h->basins.p, b4, cudaMemcpyDefault, h->stream));
        if (uparea_out) {
            void* out_dev = nullptr;
            PFD_TRY(pfd_stage_out(h, uparea_out, b4, 3, &out_dev));
            PFD_TRY(uparea_cells_device(h, (int32_t*)out_dev));
            PFD_TRY(pfd_finish_out(h, uparea_out, out_dev, b4));
        }
    }
    if (h->copy_pending) {  // join the overlapped idxs_ds copy: the timed region ends when BOTH streams are done
        PFD_CUDA(h, cudaEventRecord(h->ev_copy, h->copy_stream));
        PFD_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_copy, 0));
        h->copy_pending = false;
    }
    cudaEventRecord(h->ev_stop[PFD_STAGE_TOTAL], h->stream);
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    if (n_valid) *n_valid = h->n_valid;
    if (n_pits) *n_pits = h->n_pits;
    if (nnodes) *nnodes = h->nnodes;
    return PFD_OK;
}

extern "C" int pfd_d8_flow_all(pfd_handle* h, const uint8_t* d8, int64_t nrow, int64_t ncol, void* idxs_ds_out,
                               int idx_dtype, int32_t* rank_out, int32_t* uparea_out, uint32_t* basins_out,
                               int64_t* n_valid, int64_t* n_pits, int64_t* nnodes) {
    PFD_TRY(check_handle(h));
    const int rc = flow_all_impl(h, d8, nrow, ncol, idxs_ds_out, idx_dtype, rank_out, uparea_out, basins_out, n_valid, n_pits, nnodes);
    if (rc != PFD_OK) {
        // every error exit: nothing may still be in flight into the caller's buffers, and the handle holds no half-built graph
        const std::string msg = h->err;
        if (h->copy_pending) {
            cudaStreamSynchronize(h->copy_stream);
            h->copy_pending = false;
        }
        cudaStreamSynchronize(h->stream);
        cudaGetLastError();
        invalidate(h);
        h->err = msg;
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------------------
// verification plumbing (pfd_verify.cuh): finished outputs against their defining recurrences, any raster size
// ---------------------------------------------------------------------------------------------------------
static int verify_begin(pfd_handle* h, const char* who, VerifyCounts** vc) {
    PFD_TRY(check_handle(h));
    if (!h->parsed) return pfd_fail(h, PFD_ERR_STATE, std::string(who) + ": no raster parsed on this handle");
    if (h->tiled) return pfd_fail(h, PFD_ERR_UNSUPPORTED, std::string(who) + ": this handle holds a row block");
    PFD_TRY(ensure_upmask(h));
    PFD_TRY(pfd_reserve(h, h->verify, sizeof(VerifyCounts)));
    PFD_CUDA(h, cudaMemsetAsync(h->verify.p, 0, sizeof(VerifyCounts), h->stream));
    *vc = (VerifyCounts*)h->verify.p;
    return PFD_OK;
}

static int verify_end(pfd_handle* h, int64_t* n_bad, int first, int count) {
    VerifyCounts host;
    PFD_CUDA(h, cudaMemcpyAsync(&host, h->verify.p, sizeof(host), cudaMemcpyDeviceToHost, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int k = 0; k < count; ++k) n_bad[k] = (int64_t)host.bad[first + k];
    return PFD_OK;
}

extern "C" int pfd_verify_flow(pfd_handle* h, const void* idxs_ds, int idx_dtype, const int32_t* rank, const int32_t* uparea,
                               const uint32_t* basins, int64_t* n_bad) {
    VerifyCounts* vc = nullptr;
    PFD_TRY(verify_begin(h, "pfd_verify_flow", &vc));
    if (!n_bad) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_verify_flow: n_bad is null");
    if (idxs_ds && idx_dtype != PFD_I32 && idx_dtype != PFD_U32 && idx_dtype != PFD_I64 && idx_dtype != PFD_U64)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_verify_flow: bad index dtype");
    const int64_t n = h->n;
    const void *di = nullptr, *dr = nullptr, *du = nullptr, *db = nullptr;
    if (idxs_ds) PFD_TRY(pfd_stage_in(h, idxs_ds, (size_t)n * pfd_dtype_size(idx_dtype), 2, &di));
    if (rank) PFD_TRY(pfd_stage_in(h, rank, (size_t)n * 4, 3, &dr));
    if (uparea) PFD_TRY(pfd_stage_in(h, uparea, (size_t)n * 4, 4, &du));
    if (basins) PFD_TRY(pfd_stage_in(h, basins, (size_t)n * 4, 5, &db));
    const int g = grid_for(n, 256, 4, 148 * 32);
    const uint8_t *dir = (const uint8_t*)h->dir.p, *um = (const uint8_t*)h->upmask.p;
    if (idx_dtype == PFD_I64 || idx_dtype == PFD_U64)
        verify_flow_kernel<int64_t><<<g, 256, 0, h->stream>>>(dir, um, n, h->ncol, (const int64_t*)di, (const int32_t*)dr,
                                                              (const int32_t*)du, (const uint32_t*)db, vc);
    else
        verify_flow_kernel<uint32_t><<<g, 256, 0, h->stream>>>(dir, um, n, h->ncol, (const uint32_t*)di, (const int32_t*)dr,
                                                               (const int32_t*)du, (const uint32_t*)db, vc);
    PFD_LAUNCH_CHECK(h);
    if (basins && h->n_pits > 0) {
        verify_pit_ids_kernel<<<grid_for(h->n_pits, 256), 256, 0, h->stream>>>((const cell_t*)h->pits.p, h->n_pits, (const uint32_t*)db, vc);
        PFD_LAUNCH_CHECK(h);
    }
    return verify_end(h, n_bad, 0, 5);
}

extern "C" int pfd_verify_strahler(pfd_handle* h, const uint8_t* mask, const uint8_t* strord, int64_t* n_bad) {
    VerifyCounts* vc = nullptr;
    PFD_TRY(verify_begin(h, "pfd_verify_strahler", &vc));
    if (!strord || !n_bad) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_verify_strahler: null argument");
    PFD_TRY(tiles_usable(h) ? tiles_ensure(h, true, false, false) : order_impl(h, true, false));
    const int64_t n = h->n;
    const void *dm = nullptr, *ds = nullptr;
    if (mask) PFD_TRY(pfd_stage_in(h, mask, (size_t)n, 4, &dm));
    PFD_TRY(pfd_stage_in(h, strord, (size_t)n, 5, &ds));
    verify_strahler_kernel<<<grid_for(n, 256, 4, 148 * 32), 256, 0, h->stream>>>(
        (const uint8_t*)h->dir.p, (const uint8_t*)h->upmask.p, (const int32_t*)h->rank.p, (const uint8_t*)dm, n, h->ncol,
        (const uint8_t*)ds, vc);
    PFD_LAUNCH_CHECK(h);
    return verify_end(h, n_bad, 5, 1);
}

extern "C" int pfd_verify_hand(pfd_handle* h, const uint8_t* drain, const void* elevtn, int elev_dtype, const double* hand,
                               int64_t* n_bad) {
    VerifyCounts* vc = nullptr;
    PFD_TRY(verify_begin(h, "pfd_verify_hand", &vc));
    if (!drain || !elevtn || !hand || !n_bad) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_verify_hand: null argument");
    if (elev_dtype != PFD_F32 && elev_dtype != PFD_F64) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_verify_hand: elevtn must be float32 or float64");
    PFD_TRY(tiles_usable(h) ? tiles_ensure(h, true, false, false) : order_impl(h, true, false));
    const int64_t n = h->n;
    const void *dd = nullptr, *de = nullptr, *dh = nullptr;
    PFD_TRY(pfd_stage_in(h, drain, (size_t)n, 2, &dd));
    PFD_TRY(pfd_stage_in(h, elevtn, (size_t)n * pfd_dtype_size(elev_dtype), 4, &de));
    PFD_TRY(pfd_stage_in(h, hand, (size_t)n * 8, 5, &dh));
    const int g = grid_for(n, 256, 4, 148 * 32);
    if (elev_dtype == PFD_F32)
        verify_hand_kernel<float><<<g, 256, 0, h->stream>>>((const uint8_t*)h->dir.p, (const int32_t*)h->rank.p, (const uint8_t*)dd,
                                                            (const float*)de, n, h->ncol, (const double*)dh, vc);
    else
        verify_hand_kernel<double><<<g, 256, 0, h->stream>>>((const uint8_t*)h->dir.p, (const int32_t*)h->rank.p, (const uint8_t*)dd,
                                                             (const double*)de, n, h->ncol, (const double*)dh, vc);
    PFD_LAUNCH_CHECK(h);
    return verify_end(h, n_bad, 5, 1);
}

template <typename T>
static void verify_accuflux_launch(pfd_handle* h, const void* data, const NoData& nd, const void* accu, VerifyCounts* vc) {
    verify_accuflux_kernel<T><<<grid_for(h->n, 256, 4, 148 * 32), 256, 0, h->stream>>>(
        (const uint8_t*)h->dir.p, (const uint8_t*)h->upmask.p, (const int32_t*)h->rank.p, (const T*)data, nd, h->n, h->ncol,
        (const T*)accu, vc);
}

extern "C" int pfd_verify_accuflux(pfd_handle* h, const void* data, int dtype, double nodata_f, int64_t nodata_i, int nodata_is_int,
                                   const void* accu, int64_t* n_bad) {
    VerifyCounts* vc = nullptr;
    PFD_TRY(verify_begin(h, "pfd_verify_accuflux", &vc));
    const size_t esz = pfd_dtype_size(dtype);
    if (!data || !accu || !n_bad || !esz) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_verify_accuflux: bad argument");
    PFD_TRY(tiles_usable(h) ? tiles_ensure(h, true, false, false) : order_impl(h, true, false));
    const void *dd = nullptr, *da = nullptr;
    PFD_TRY(pfd_stage_in(h, data, (size_t)h->n * esz, 4, &dd));
    PFD_TRY(pfd_stage_in(h, accu, (size_t)h->n * esz, 5, &da));
    NoData nd{nodata_f, (long long)nodata_i, nodata_is_int};
    switch (dtype) {
    case PFD_I8: verify_accuflux_launch<int8_t>(h, dd, nd, da, vc); break;
    case PFD_U8: verify_accuflux_launch<uint8_t>(h, dd, nd, da, vc); break;
    case PFD_I16: verify_accuflux_launch<int16_t>(h, dd, nd, da, vc); break;
    case PFD_U16: verify_accuflux_launch<uint16_t>(h, dd, nd, da, vc); break;
    case PFD_I32: verify_accuflux_launch<int32_t>(h, dd, nd, da, vc); break;
    case PFD_U32: verify_accuflux_launch<uint32_t>(h, dd, nd, da, vc); break;
    case PFD_I64: verify_accuflux_launch<int64_t>(h, dd, nd, da, vc); break;
    case PFD_U64: verify_accuflux_launch<uint64_t>(h, dd, nd, da, vc); break;
    case PFD_F32: verify_accuflux_launch<float>(h, dd, nd, da, vc); break;
    default: verify_accuflux_launch<double>(h, dd, nd, da, vc); break;
    }
    PFD_LAUNCH_CHECK(h);
    return verify_end(h, n_bad, 5, 1);
}

// ---------------------------------------------------------------------------------------------------------
// further rows of SURVEY.md §8f: arithmetics.upstream_sum, basins.subbasins_streamorder
// ---------------------------------------------------------------------------------------------------------
// ordered numbering of the cells selected by `pred` (positions of the sequence, optionally reversed) into
// h->sub_idxs / labels; leaves the count in h->n_sub
// ordered compaction of the positions [0, m) whose cell (seq[q], seq[m-1-q] or q itself: REVERSED = 0 / 1 / 2) satisfies
// pred -> h->sub_idxs / h->n_sub; label_dev (may be null) receives the 1-based ordinal per selected cell
template <class Pred, int REVERSED, typename LABEL>
static int compact_cells(pfd_handle* h, const cell_t* seq, int64_t m, Pred pred, LABEL* label_dev) {
    h->have_sub_labels = h->have_sub_slices = false;
    const int64_t nblk = std::max<int64_t>(1, (m + CP_CHUNK - 1) / CP_CHUNK);
    PFD_TRY(pfd_reserve(h, h->blk_counts, (size_t)nblk * sizeof(uint32_t)));
    PFD_TRY(pfd_reserve(h, h->blk_offsets, (size_t)(nblk + 1) * sizeof(unsigned long long)));
    compact_count_kernel<Pred, REVERSED><<<(unsigned)nblk, CP_THREADS, 0, h->stream>>>(seq, m, pred, (uint32_t*)h->blk_counts.p);
    PFD_LAUNCH_CHECK(h);
    scan_counts_kernel<<<1, 1024, 0, h->stream>>>((const uint32_t*)h->blk_counts.p, nblk, (unsigned long long*)h->blk_offsets.p);
    PFD_LAUNCH_CHECK(h);
    unsigned long long total = 0;
    PFD_CUDA(h, cudaMemcpyAsync(&total, (unsigned long long*)h->blk_offsets.p + nblk, sizeof(total), cudaMemcpyDeviceToHost, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    h->n_sub = (int64_t)total;
    PFD_TRY(pfd_reserve(h, h->sub_idxs, (size_t)std::max<int64_t>(h->n_sub, 1) * sizeof(cell_t)));
    if (h->n_sub > 0) {
        compact_scatter_kernel<Pred, REVERSED, LABEL><<<(unsigned)nblk, CP_THREADS, 0, h->stream>>>(
            seq, m, pred, (const unsigned long long*)h->blk_offsets.p, (cell_t*)h->sub_idxs.p, label_dev);
        PFD_LAUNCH_CHECK(h);
    }
    return PFD_OK;
}
template <class Pred, int REVERSED, typename LABEL>
static int number_outlets(pfd_handle* h, Pred pred, LABEL* label_dev) {
    return compact_cells<Pred, REVERSED, LABEL>(h, (const cell_t*)h->seq.p, h->nnodes, pred, label_dev);
}

template <typename T>
static int upstream_sum_typed(pfd_handle* h, const void* data_dev, void* out_dev, const NoData& nd) {
    // arr_sum[idx0] = nodata casts the nodata value to the data dtype
    const T ndv = nd.is_int ? (T)nd.i : (T)nd.f;
    upstream_sum_kernel<T><<<grid_for(h->n, 256, 4), 256, 0, h->stream>>>((const uint8_t*)h->dir.p, (const uint8_t*)h->upmask.p,
                                                                       (const T*)data_dev, h->n, h->ncol, nd, ndv, (T*)out_dev);
    PFD_LAUNCH_CHECK(h);
    return PFD_OK;
}

extern "C" int pfd_upstream_sum(pfd_handle* h, const void* data, int dtype, double nodata_f, int64_t nodata_i, int nodata_is_int,
                                void* out) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    if (!h->parsed) return pfd_fail(h, PFD_ERR_STATE, "pfd_upstream_sum: no raster parsed on this handle");
    if (h->tiled) return pfd_fail(h, PFD_ERR_UNSUPPORTED, "this handle holds a row block: only the pfd_tiled_* entry points apply");
    if (!data || !out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_upstream_sum: null array");
    const size_t esz = pfd_dtype_size(dtype);
    if (!esz) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_upstream_sum: unknown dtype");
    const size_t bytes = (size_t)h->n * esz;
    void* out_dev = nullptr;
    const void* data_dev = nullptr;
    PFD_TRY(pfd_stage_out(h, out, bytes, 3, &out_dev));
    PFD_TRY(pfd_stage_in(h, data, bytes, 5, &data_dev));
    if (data_dev == out_dev) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_upstream_sum: data and out must not alias");
    PFD_TRY(ensure_upmask(h));
    NoData nd{nodata_f, (long long)nodata_i, nodata_is_int};
    int rc;
    switch (dtype) {
    case PFD_I8: rc = upstream_sum_typed<int8_t>(h, data_dev, out_dev, nd); break;
    case PFD_U8: rc = upstream_sum_typed<uint8_t>(h, data_dev, out_dev, nd); break;
    case PFD_I16: rc = upstream_sum_typed<int16_t>(h, data_dev, out_dev, nd); break;
    case PFD_U16: rc = upstream_sum_typed<uint16_t>(h, data_dev, out_dev, nd); break;
    case PFD_I32: rc = upstream_sum_typed<int32_t>(h, data_dev, out_dev, nd); break;
    case PFD_U32: rc = upstream_sum_typed<uint32_t>(h, data_dev, out_dev, nd); break;
    case PFD_I64: rc = upstream_sum_typed<int64_t>(h, data_dev, out_dev, nd); break;
    case PFD_U64: rc = upstream_sum_typed<uint64_t>(h, data_dev, out_dev, nd); break;
    case PFD_F32: rc = upstream_sum_typed<float>(h, data_dev, out_dev, nd); break;
    default: rc = upstream_sum_typed<double>(h, data_dev, out_dev, nd); break;
    }
    PFD_TRY(rc);
    PFD_TRY(pfd_finish_out(h, out, out_dev, bytes));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    return PFD_OK;
}

extern "C" int pfd_subbasins_streamorder(pfd_handle* h, const uint8_t* strord, const uint8_t* mask, int64_t min_sto,
                                         int32_t* subbas_out, int64_t* n_outlets) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    if (!strord || !subbas_out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_subbasins_streamorder: null array");
    PFD_TRY(order_impl(h, false, false));
    const int64_t n = h->n;
    const void *so_dev = nullptr, *mask_dev = nullptr;
    PFD_TRY(pfd_stage_in(h, strord, (size_t)n, 5, &so_dev));
    if (mask) PFD_TRY(pfd_stage_in(h, mask, (size_t)n, 4, &mask_dev));
    void* out_dev = nullptr;
    PFD_TRY(pfd_stage_out(h, subbas_out, (size_t)n * sizeof(int32_t), 3, &out_dev));
    if (min_sto < 0) {  // relative to the global maximum (basins.py:88-89)
        unsigned int* mx = reinterpret_cast<unsigned int*>((unsigned long long*)h->counters.p + 6);
        PFD_CUDA(h, cudaMemsetAsync(mx, 0, sizeof(unsigned int), h->stream));
        max_u8_kernel<<<grid_for(n, 256, 16, 148 * 8), 256, 0, h->stream>>>((const uint8_t*)so_dev, n, mx);
        PFD_LAUNCH_CHECK(h);
        unsigned int hmx = 0;
        PFD_CUDA(h, cudaMemcpyAsync(&hmx, mx, sizeof(hmx), cudaMemcpyDeviceToHost, h->stream));
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
        min_sto += (int64_t)hmx;
    }
    const int min_sto_i = (int)std::max<int64_t>(std::min<int64_t>(min_sto, 1 << 20), -(1 << 20));
    SubbasinOutletPred pred{(const uint8_t*)h->dir.p, (const uint8_t*)so_dev, (const uint8_t*)mask_dev, min_sto_i, h->ncol};
    PFD_CUDA(h, cudaMemsetAsync(out_dev, 0, (size_t)n * sizeof(int32_t), h->stream));
    // outlets are numbered while walking the sequence from up- to downstream (seq[::-1], basins.py:92)
    PFD_TRY((number_outlets<SubbasinOutletPred, true, int32_t>(h, pred, (int32_t*)out_dev)));
    if (h->n_sub > 0) {
        FillUpOp<int32_t> op{(const uint8_t*)h->dir.p, (int32_t*)out_dev, h->ncol};  // core.fillnodata_upstream(.., 0)
        PFD_TRY((run_sweep<FillUpOp<int32_t>, false>(h, op, 0)));
    }
    PFD_TRY(pfd_finish_out(h, subbas_out, out_dev, (size_t)n * sizeof(int32_t)));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    if (n_outlets) *n_outlets = h->n_sub;
    return PFD_OK;
}

template <typename T, typename IDX>
static int subbasins_area_typed(pfd_handle* h, const void* main_dev, const void* upa_dev, double area_min, uint32_t* out_dev) {
    const int64_t n = h->n;
    PFD_TRY(pfd_reserve(h, h->scratch[2], (size_t)n * sizeof(T)));
    PFD_TRY(pfd_reserve(h, h->scratch[1], (size_t)n));
    T* upa_out = (T*)h->scratch[2].p;
    uint8_t* flag = (uint8_t*)h->scratch[1].p;
    PFD_CUDA(h, cudaMemcpyAsync(upa_out, upa_dev, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, h->stream));
    PFD_CUDA(h, cudaMemsetAsync(flag, 0, (size_t)n, h->stream));
    PFD_CUDA(h, cudaMemsetAsync(out_dev, 0, (size_t)n * sizeof(uint32_t), h->stream));
    SubbasinsAreaOp<T, IDX> op{(const uint8_t*)h->dir.p, (const uint8_t*)h->upmask.p, (const IDX*)main_dev, (const T*)upa_dev,
                               upa_out, flag, area_min, h->ncol};
    PFD_TRY((run_sweep<SubbasinsAreaOp<T, IDX>, false>(h, op, 0)));
    FlagPred pred{flag};
    PFD_TRY((number_outlets<FlagPred, false, uint32_t>(h, pred, out_dev)));  // numbered along seq (basins.py:208-223)
    if (h->n_sub > 0) {
        FillUpOp<uint32_t> fill{(const uint8_t*)h->dir.p, out_dev, h->ncol};  // core.fillnodata_upstream(.., 0)
        PFD_TRY((run_sweep<FillUpOp<uint32_t>, false>(h, fill, 0)));
    }
    return PFD_OK;
}

extern "C" int pfd_subbasins_area(pfd_handle* h, const void* idxs_us_main, int idx_dtype, const void* uparea, int dtype,
                                  double area_min, uint32_t* subbas_out, int64_t* n_outlets) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    if (!idxs_us_main || !uparea || !subbas_out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_subbasins_area: null array");
    const size_t isz = pfd_dtype_size(idx_dtype), esz = pfd_dtype_size(dtype);
    if ((isz != 4 && isz != 8) || idx_dtype == PFD_F32 || idx_dtype == PFD_F64)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_subbasins_area: index dtype must be a 32/64-bit integer");
    if (dtype != PFD_I32 && dtype != PFD_I64 && dtype != PFD_F32 && dtype != PFD_F64)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_subbasins_area: uparea must be int32, int64, float32 or float64");
    PFD_TRY(order_impl(h, false, false));
    PFD_TRY(ensure_upmask(h));
    const int64_t n = h->n;
    void* out_dev = nullptr;
    PFD_TRY(pfd_stage_out(h, subbas_out, (size_t)n * sizeof(uint32_t), 3, &out_dev));
    const void *main_dev = nullptr, *upa_dev = nullptr;
    PFD_TRY(pfd_stage_in(h, idxs_us_main, (size_t)n * isz, 5, &main_dev));
    PFD_TRY(pfd_stage_in(h, uparea, (size_t)n * esz, 4, &upa_dev));
    int rc;
#define SUBAREA(T)                                                                                                  \
    (isz == 4 ? subbasins_area_typed<T, uint32_t>(h, main_dev, upa_dev, area_min, (uint32_t*)out_dev)               \
              : subbasins_area_typed<T, int64_t>(h, main_dev, upa_dev, area_min, (uint32_t*)out_dev))
    switch (dtype) {
    case PFD_I32: rc = SUBAREA(int32_t); break;
    case PFD_I64: rc = SUBAREA(int64_t); break;
    case PFD_F32: rc = SUBAREA(float); break;
    default: rc = SUBAREA(double); break;
    }
#undef SUBAREA
    PFD_TRY(rc);
    PFD_TRY(pfd_finish_out(h, subbas_out, out_dev, (size_t)n * sizeof(uint32_t)));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    if (n_outlets) *n_outlets = h->n_sub;
    return PFD_OK;
}

template <typename T, bool MEDIAN>
static int moving_window_typed(pfd_handle* h, const void* data_dev, const void* w_dev, int nwin, const void* main_dev, size_t isz,
                               const void* so_dev, double nodata, void* out_dev, unsigned int* flag) {
    const int g = grid_for(h->n, 128, 1, 148 * 64);
    if (isz == 4)
        moving_window_kernel<T, uint32_t, MEDIAN><<<g, 128, 0, h->stream>>>(
            (const uint8_t*)h->dir.p, (const uint32_t*)main_dev, (const uint8_t*)so_dev, (const T*)data_dev, (const double*)w_dev, h->n,
            h->ncol, nwin, nodata, (T*)out_dev, flag);
    else
        moving_window_kernel<T, int64_t, MEDIAN><<<g, 128, 0, h->stream>>>(
            (const uint8_t*)h->dir.p, (const int64_t*)main_dev, (const uint8_t*)so_dev, (const T*)data_dev, (const double*)w_dev, h->n,
            h->ncol, nwin, nodata, (T*)out_dev, flag);
    PFD_LAUNCH_CHECK(h);
    return PFD_OK;
}

// shared driver of pfd_moving_average (median = 0) and pfd_moving_median (median = 1)
static int moving_window(pfd_handle* h, const char* who, int median, const void* data, int dtype, const void* weights, int wdtype,
                         int nwin, const void* idxs_us_main, int idx_dtype, const uint8_t* strord, double nodata, void* out) {
    PFD_TRY(check_handle(h));
    stage_reset(h);
    if (!h->parsed) return pfd_fail(h, PFD_ERR_STATE, std::string(who) + ": no raster parsed on this handle");
    if (h->tiled) return pfd_fail(h, PFD_ERR_UNSUPPORTED, "this handle holds a row block: only the pfd_tiled_* entry points apply");
    if (!data || !idxs_us_main || !out) return pfd_fail(h, PFD_ERR_INVALID_ARG, std::string(who) + ": null array");
    if (dtype != PFD_F32 && dtype != PFD_F64) return pfd_fail(h, PFD_ERR_INVALID_ARG, std::string(who) + ": data must be float32 or float64");
    if (weights && wdtype != PFD_F64)  // arithmetics.py:101 only types with float64 weights
        return pfd_fail(h, PFD_ERR_INVALID_ARG, std::string(who) + ": weights must be float64");
    if (nwin < 0 || nwin > MW_NMAX) return pfd_fail(h, PFD_ERR_UNSUPPORTED, std::string(who) + ": n must be in 0..64");
    const size_t isz = pfd_dtype_size(idx_dtype), esz = pfd_dtype_size(dtype);
    if ((isz != 4 && isz != 8) || idx_dtype == PFD_F32 || idx_dtype == PFD_F64)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, std::string(who) + ": index dtype must be a 32/64-bit integer");
    const int64_t n = h->n;
    void* out_dev = nullptr;
    const void *data_dev = nullptr, *w_dev = nullptr, *main_dev = nullptr, *so_dev = nullptr;
    PFD_TRY(pfd_stage_out(h, out, (size_t)n * esz, 3, &out_dev));
    PFD_TRY(pfd_stage_in(h, data, (size_t)n * esz, 5, &data_dev));
    PFD_TRY(pfd_stage_in(h, idxs_us_main, (size_t)n * isz, 4, &main_dev));
    if (weights) PFD_TRY(pfd_stage_in(h, weights, (size_t)n * pfd_dtype_size(wdtype), 2, &w_dev));
    if (strord) PFD_TRY(pfd_stage_in(h, strord, (size_t)n, 1, &so_dev));
    if (data_dev == out_dev) return pfd_fail(h, PFD_ERR_INVALID_ARG, std::string(who) + ": data and out must not alias");
    PFD_TRY(pfd_reserve(h, h->counters, 8 * sizeof(unsigned long long)));
    unsigned int* flag = reinterpret_cast<unsigned int*>((unsigned long long*)h->counters.p + 5);
    PFD_CUDA(h, cudaMemsetAsync(flag, 0, sizeof(unsigned int), h->stream));
    int rc;
    if (median) {
        rc = dtype == PFD_F32 ? moving_window_typed<float, true>(h, data_dev, nullptr, nwin, main_dev, isz, so_dev, nodata, out_dev, flag)
                              : moving_window_typed<double, true>(h, data_dev, nullptr, nwin, main_dev, isz, so_dev, nodata, out_dev, flag);
    } else {
        rc = dtype == PFD_F32 ? moving_window_typed<float, false>(h, data_dev, w_dev, nwin, main_dev, isz, so_dev, nodata, out_dev, flag)
                              : moving_window_typed<double, false>(h, data_dev, w_dev, nwin, main_dev, isz, so_dev, nodata, out_dev, flag);
    }
    PFD_TRY(rc);
    unsigned int hflag = 0;
    PFD_CUDA(h, cudaMemcpyAsync(&hflag, flag, sizeof(hflag), cudaMemcpyDeviceToHost, h->stream));
    PFD_TRY(pfd_finish_out(h, out, out_dev, (size_t)n * esz));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    if (hflag & 8u) return pfd_fail(h, PFD_ERR_INVALID_ARG, std::string(who) + ": idxs_us_main holds an index outside the raster");
    return PFD_OK;
}

extern "C" int pfd_moving_average(pfd_handle* h, const void* data, int dtype, const void* weights, int wdtype, int n,
                                  const void* idxs_us_main, int idx_dtype, const uint8_t* strord, double nodata, void* out) {
    return moving_window(h, "pfd_moving_average", 0, data, dtype, weights, wdtype, n, idxs_us_main, idx_dtype, strord, nodata, out);
}

extern "C" int pfd_moving_median(pfd_handle* h, const void* data, int dtype, int n, const void* idxs_us_main, int idx_dtype,
                                 const uint8_t* strord, double nodata, void* out) {
    return moving_window(h, "pfd_moving_median", 1, data, dtype, nullptr, 0, n, idxs_us_main, idx_dtype, strord, nodata, out);
}

// ---------------------------------------------------------------------------------------------------------
// local traces and region post-processing (pfd_local.cuh)
// ---------------------------------------------------------------------------------------------------------
static int require_raster(pfd_handle* h, const char* who) {
    PFD_TRY(check_handle(h));
    if (!h->parsed) return pfd_fail(h, PFD_ERR_STATE, std::string(who) + ": no raster parsed on this handle");
    if (h->tiled) return pfd_fail(h, PFD_ERR_UNSUPPORTED, "this handle holds a row block: only the pfd_tiled_* entry points apply");
    return PFD_OK;
}

extern "C" int pfd_downstream(pfd_handle* h, const void* data, int dtype, void* out) {
    PFD_TRY(require_raster(h, "pfd_downstream"));
    stage_reset(h);
    if (!data || !out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_downstream: null array");
    const size_t esz = pfd_dtype_size(dtype);
    if (!esz) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_downstream: unknown dtype");
    const int64_t n = h->n;
    void* out_dev = nullptr;
    const void* data_dev = nullptr;
    PFD_TRY(pfd_stage_out(h, out, (size_t)n * esz, 3, &out_dev));
    PFD_TRY(pfd_stage_in(h, data, (size_t)n * esz, 5, &data_dev));
    if (data_dev == out_dev) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_downstream: data and out must not alias");
    const int g = grid_for(n, 256, 4);
    const uint8_t* dir = (const uint8_t*)h->dir.p;
    switch (esz) {
    case 1: downstream_kernel<uint8_t><<<g, 256, 0, h->stream>>>(dir, (const uint8_t*)data_dev, n, h->ncol, (uint8_t*)out_dev); break;
    case 2: downstream_kernel<uint16_t><<<g, 256, 0, h->stream>>>(dir, (const uint16_t*)data_dev, n, h->ncol, (uint16_t*)out_dev); break;
    case 4: downstream_kernel<uint32_t><<<g, 256, 0, h->stream>>>(dir, (const uint32_t*)data_dev, n, h->ncol, (uint32_t*)out_dev); break;
    default: downstream_kernel<uint64_t><<<g, 256, 0, h->stream>>>(dir, (const uint64_t*)data_dev, n, h->ncol, (uint64_t*)out_dev); break;
    }
    PFD_LAUNCH_CHECK(h);
    PFD_TRY(pfd_finish_out(h, out, out_dev, (size_t)n * esz));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    return PFD_OK;
}

template <typename IDX, typename OUT>
static void launch_trace(pfd_handle* h, int up, const void* main_dev, const uint8_t* mask_dev, const int64_t* starts_dev, int64_t n0,
                         int has_max, double max_length, const double* hop_dev, const long long* offsets, int64_t* counts,
                         int64_t* ends, double* dists, void* paths, unsigned int* flag) {
    const int g = grid_for(n0, 64, 1, 148 * 32);
    if (up)
        trace_kernel<IDX, OUT, true><<<g, 64, 0, h->stream>>>((const uint8_t*)h->dir.p, (const IDX*)main_dev, mask_dev, starts_dev, n0, h->n,
                                                             h->ncol, has_max, max_length, hop_dev, offsets, counts, ends, dists,
                                                             (OUT*)paths, flag);
    else
        trace_kernel<IDX, OUT, false><<<g, 64, 0, h->stream>>>((const uint8_t*)h->dir.p, (const IDX*)main_dev, mask_dev, starts_dev, n0, h->n,
                                                              h->ncol, has_max, max_length, hop_dev, offsets, counts, ends, dists,
                                                              (OUT*)paths, flag);
}

extern "C" int pfd_trace(pfd_handle* h, const int64_t* starts, int64_t n0, int direction, const void* idxs_us_main, int idx_dtype,
                         const uint8_t* mask, int has_max_length, double max_length, const double* hop_table, int64_t* counts_out,
                         int64_t* ends_out, double* dists_out, void* paths_out, int path_dtype, int64_t paths_capacity) {
    PFD_TRY(require_raster(h, "pfd_trace"));
    stage_reset(h);
    if (n0 < 0 || (n0 > 0 && !starts)) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_trace: null start indices");
    if (direction != 0 && direction != 1) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_trace: direction must be 0 (down) or 1 (up)");
    if (direction == 1 && !idxs_us_main) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_trace: upstream traces need idxs_us_main");
    const size_t isz = direction == 1 ? pfd_dtype_size(idx_dtype) : 4;
    if ((isz != 4 && isz != 8) || (direction == 1 && (idx_dtype == PFD_F32 || idx_dtype == PFD_F64)))
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_trace: index dtype must be a 32/64-bit integer");
    const size_t psz = paths_out ? pfd_dtype_size(path_dtype) : 4;
    if ((psz != 4 && psz != 8) || (paths_out && (path_dtype == PFD_F32 || path_dtype == PFD_F64)))
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_trace: path dtype must be a 32/64-bit integer");
    if (n0 == 0) return PFD_OK;
    const int64_t n = h->n;
    const void *starts_dev = nullptr, *mask_dev = nullptr, *main_dev = nullptr, *hop_dev = nullptr;
    PFD_TRY(pfd_stage_in(h, starts, (size_t)n0 * sizeof(int64_t), 1, &starts_dev));
    if (mask) PFD_TRY(pfd_stage_in(h, mask, (size_t)n, 4, &mask_dev));
    if (direction == 1) PFD_TRY(pfd_stage_in(h, idxs_us_main, (size_t)n * isz, 5, &main_dev));
    if (hop_table) PFD_TRY(pfd_stage_in(h, hop_table, (size_t)h->nrow * 6 * sizeof(double), 2, &hop_dev));
    // per-start results: counts | ends | dists | offsets
    PFD_TRY(pfd_reserve(h, h->scratch[0], (size_t)(4 * n0 + 1) * 8));
    int64_t* counts = (int64_t*)h->scratch[0].p;
    int64_t* ends = counts + n0;
    double* dists = (double*)(ends + n0);
    long long* offsets = (long long*)(dists + n0);
    PFD_TRY(pfd_reserve(h, h->counters, 8 * sizeof(unsigned long long)));
    unsigned int* flag = reinterpret_cast<unsigned int*>((unsigned long long*)h->counters.p + 5);
    PFD_CUDA(h, cudaMemsetAsync(flag, 0, sizeof(unsigned int), h->stream));
#define TRACE(PATHS, OFFS)                                                                                                   \
    do {                                                                                                                     \
        if (isz == 4 && psz == 4)                                                                                            \
            launch_trace<uint32_t, uint32_t>(h, direction, main_dev, (const uint8_t*)mask_dev, (const int64_t*)starts_dev, n0, \
                                             has_max_length, max_length, (const double*)hop_dev, OFFS, counts, ends, dists, PATHS, flag); \
        else if (isz == 4)                                                                                                   \
            launch_trace<uint32_t, int64_t>(h, direction, main_dev, (const uint8_t*)mask_dev, (const int64_t*)starts_dev, n0, \
                                            has_max_length, max_length, (const double*)hop_dev, OFFS, counts, ends, dists, PATHS, flag); \
        else if (psz == 4)                                                                                                   \
            launch_trace<int64_t, uint32_t>(h, direction, main_dev, (const uint8_t*)mask_dev, (const int64_t*)starts_dev, n0, \
                                            has_max_length, max_length, (const double*)hop_dev, OFFS, counts, ends, dists, PATHS, flag); \
        else                                                                                                                 \
            launch_trace<int64_t, int64_t>(h, direction, main_dev, (const uint8_t*)mask_dev, (const int64_t*)starts_dev, n0,  \
                                           has_max_length, max_length, (const double*)hop_dev, OFFS, counts, ends, dists, PATHS, flag); \
        PFD_LAUNCH_CHECK(h);                                                                                                 \
    } while (0)
    TRACE(nullptr, nullptr);
    unsigned int hflag = 0;
    if (paths_out) {
        trace_offsets_kernel<<<1, 32, 0, h->stream>>>(counts, n0, offsets);
        PFD_LAUNCH_CHECK(h);
        long long total = 0;
        PFD_CUDA(h, cudaMemcpyAsync(&total, offsets + n0, sizeof(total), cudaMemcpyDeviceToHost, h->stream));
        PFD_CUDA(h, cudaMemcpyAsync(&hflag, flag, sizeof(hflag), cudaMemcpyDeviceToHost, h->stream));
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
        if (!hflag) {
            if (total > paths_capacity) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_trace: paths_out is too small for the traces");
            void* paths_dev = nullptr;
            PFD_TRY(pfd_stage_out(h, paths_out, (size_t)std::max<long long>(total, 1) * psz, 3, &paths_dev));
            TRACE(paths_dev, offsets);
            PFD_TRY(pfd_finish_out(h, paths_out, paths_dev, (size_t)total * psz));
        }
    }
#undef TRACE
    if (counts_out) PFD_CUDA(h, cudaMemcpyAsync(counts_out, counts, (size_t)n0 * 8, cudaMemcpyDefault, h->stream));
    if (ends_out) PFD_CUDA(h, cudaMemcpyAsync(ends_out, ends, (size_t)n0 * 8, cudaMemcpyDefault, h->stream));
    if (dists_out) PFD_CUDA(h, cudaMemcpyAsync(dists_out, dists, (size_t)n0 * 8, cudaMemcpyDefault, h->stream));
    PFD_CUDA(h, cudaMemcpyAsync(&hflag, flag, sizeof(hflag), cudaMemcpyDeviceToHost, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    if (hflag & 8u) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_trace: index outside the raster (start cell or idxs_us_main)");
    if (hflag & 16u) return pfd_fail(h, PFD_ERR_UNSUPPORTED, "pfd_trace: a trace does not end (loop in the flow directions and no stop condition)");
    return PFD_OK;
}

// shared driver of pfd_inflow_idxs (inflow = 1) and pfd_outflow_idxs (inflow = 0)
static int inout_idxs(pfd_handle* h, const char* who, int inflow, const uint8_t* region, int64_t* n_out) {
    PFD_TRY(require_raster(h, who));
    stage_reset(h);
    if (!region) return pfd_fail(h, PFD_ERR_INVALID_ARG, std::string(who) + ": null array");
    PFD_TRY(order_impl(h, false, false));
    const int64_t n = h->n;
    const void* region_dev = nullptr;
    PFD_TRY(pfd_stage_in(h, region, (size_t)n, 4, &region_dev));
    PFD_TRY(pfd_reserve(h, h->scratch[1], (size_t)n));
    uint8_t* st = (uint8_t*)h->scratch[1].p;
    PFD_CUDA(h, cudaMemsetAsync(st, 1, (size_t)n, h->stream));
    h->have_sub_labels = h->have_sub_slices = false;
    StateBit2Pred pred{st};
    if (inflow) {
        PFD_TRY(ensure_upmask(h));
        InflowOp op{(const uint8_t*)h->upmask.p, (const uint8_t*)region_dev, st, h->ncol};
        PFD_TRY((run_sweep<InflowOp, true>(h, op, 0)));
        PFD_TRY((number_outlets<StateBit2Pred, 1, uint32_t>(h, pred, nullptr)));  // appended while walking seq[::-1]
    } else {
        OutflowOp op{(const uint8_t*)h->dir.p, (const uint8_t*)region_dev, st, h->ncol};
        PFD_TRY((run_sweep<OutflowOp, false>(h, op, 0)));
        PFD_TRY((number_outlets<StateBit2Pred, 0, uint32_t>(h, pred, nullptr)));
    }
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    if (n_out) *n_out = h->n_sub;
    return PFD_OK;
}
extern "C" int pfd_inflow_idxs(pfd_handle* h, const uint8_t* region, int64_t* n_out) {
    return inout_idxs(h, "pfd_inflow_idxs", 1, region, n_out);
}
extern "C" int pfd_outflow_idxs(pfd_handle* h, const uint8_t* region, int64_t* n_out) {
    return inout_idxs(h, "pfd_outflow_idxs", 0, region, n_out);
}

extern "C" int pfd_interbasin_mask(pfd_handle* h, const uint8_t* region, const uint8_t* stream, uint8_t* out) {
    PFD_TRY(require_raster(h, "pfd_interbasin_mask"));
    stage_reset(h);
    if (!region || !out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_interbasin_mask: null array");
    PFD_TRY(order_impl(h, false, false));
    const int64_t n = h->n;
    const void* region_dev = nullptr;
    void* out_dev = nullptr;
    PFD_TRY(pfd_stage_in(h, region, (size_t)n, 4, &region_dev));
    PFD_TRY(pfd_stage_out(h, out, (size_t)n, 3, &out_dev));
    PFD_TRY(pfd_reserve(h, h->scratch[1], (size_t)n));
    uint8_t* st = (uint8_t*)h->scratch[1].p;
    if (stream) {
        PFD_CUDA(h, cudaMemcpyAsync(st, stream, (size_t)n, cudaMemcpyDefault, h->stream));
        PFD_TRY(ensure_upmask(h));
        AnyUpstreamOp any{(const uint8_t*)h->upmask.p, st, h->ncol};
        PFD_TRY((run_sweep<AnyUpstreamOp, true>(h, any, 0)));
    } else {
        PFD_CUDA(h, cudaMemsetAsync(st, 1, (size_t)n, h->stream));
    }
    InterbasinOp op{(const uint8_t*)h->dir.p, (const uint8_t*)region_dev, st, h->ncol};
    PFD_TRY((run_sweep<InterbasinOp, false>(h, op, 1)));
    and_mask_kernel<<<grid_for(n, 256, 4), 256, 0, h->stream>>>(st, (const uint8_t*)region_dev, n, (uint8_t*)out_dev);
    PFD_LAUNCH_CHECK(h);
    PFD_TRY(pfd_finish_out(h, out, out_dev, (size_t)n));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    return PFD_OK;
}

// np.argsort as numba compiles it (numba/misc/quicksort.py: median-of-three partition, explicit stack, insertion sort
// below 15 elements) -- regions.region_outlets sorts its outlets with it (regions.py:162), and the order of equal
// labels depends on the exact algorithm. Host side: the list holds one entry per outlet.
static void numba_argsort_i64(const std::vector<int64_t>& A, std::vector<int64_t>& R) {
    const int64_t n = (int64_t)A.size();
    R.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) R[(size_t)i] = i;
    if (n < 2) return;
    auto key = [&](int64_t pos) { return A[(size_t)R[(size_t)pos]]; };
    auto partition = [&](int64_t low, int64_t high) {
        const int64_t mid = (low + high) >> 1;
        if (key(mid) < key(low)) std::swap(R[(size_t)low], R[(size_t)mid]);
        if (key(high) < key(mid)) std::swap(R[(size_t)high], R[(size_t)mid]);
        if (key(mid) < key(low)) std::swap(R[(size_t)low], R[(size_t)mid]);
        const int64_t pivot = key(mid);
        std::swap(R[(size_t)high], R[(size_t)mid]);
        int64_t i = low, j = high - 1;
        while (true) {
            while (i < high && key(i) < pivot) ++i;
            while (j >= low && pivot < key(j)) --j;
            if (i >= j) break;
            std::swap(R[(size_t)i], R[(size_t)j]);
            ++i;
            --j;
        }
        std::swap(R[(size_t)i], R[(size_t)high]);
        return i;
    };
    std::vector<std::pair<int64_t, int64_t>> stack;
    stack.emplace_back(0, n - 1);
    while (!stack.empty()) {
        int64_t low = stack.back().first, high = stack.back().second;
        stack.pop_back();
        while (high - low >= 15) {
            const int64_t i = partition(low, high);
            if (high - i > i - low) {
                if (high > i) stack.emplace_back(i + 1, high);
                high = i - 1;
            } else {
                if (i > low) stack.emplace_back(low, i - 1);
                low = i + 1;
            }
        }
        for (int64_t i = low + 1; i <= high; ++i) {  // insertion sort
            const int64_t k = R[(size_t)i];
            const int64_t v = A[(size_t)k];
            int64_t j = i;
            while (j > low && v < A[(size_t)R[(size_t)(j - 1)]]) {
                R[(size_t)j] = R[(size_t)(j - 1)];
                --j;
            }
            R[(size_t)j] = k;
        }
    }
}

template <typename T>
static int region_outlets_typed(pfd_handle* h, const void* reg_dev) {
    RegionOutletPred<T> pred{(const uint8_t*)h->dir.p, (const T*)reg_dev, h->ncol};
    PFD_TRY((number_outlets<RegionOutletPred<T>, 1, uint32_t>(h, pred, nullptr)));  // found while walking seq[::-1]
    PFD_TRY(pfd_reserve(h, h->sub_labels, (size_t)std::max<int64_t>(h->n_sub, 1) * sizeof(int64_t)));
    if (h->n_sub > 0) {
        gather_labels_kernel<T><<<grid_for(h->n_sub, 256), 256, 0, h->stream>>>((const cell_t*)h->sub_idxs.p, h->n_sub, (const T*)reg_dev,
                                                                              (int64_t*)h->sub_labels.p);
        PFD_LAUNCH_CHECK(h);
    }
    return PFD_OK;
}

static int region_dtype_ok(int dtype) { return dtype == PFD_I32 || dtype == PFD_U32 || dtype == PFD_I64 || dtype == PFD_U64; }

extern "C" int pfd_region_outlets(pfd_handle* h, const void* regions, int dtype, int64_t* n_out) {
    PFD_TRY(require_raster(h, "pfd_region_outlets"));
    stage_reset(h);
    if (!regions) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_region_outlets: null array");
    if (!region_dtype_ok(dtype)) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_region_outlets: regions must be a 32/64-bit integer array");
    PFD_TRY(order_impl(h, false, false));
    const void* reg_dev = nullptr;
    PFD_TRY(pfd_stage_in(h, regions, (size_t)h->n * pfd_dtype_size(dtype), 5, &reg_dev));
    h->have_sub_labels = h->have_sub_slices = false;
    int rc;
    switch (dtype) {
    case PFD_I32: rc = region_outlets_typed<int32_t>(h, reg_dev); break;
    case PFD_U32: rc = region_outlets_typed<uint32_t>(h, reg_dev); break;
    case PFD_I64: rc = region_outlets_typed<int64_t>(h, reg_dev); break;
    default: rc = region_outlets_typed<uint64_t>(h, reg_dev); break;
    }
    PFD_TRY(rc);
    const int64_t m = h->n_sub;
    if (m > 1) {  // sort = np.argsort(lbs); lbs[sort], idxs_out[sort] (regions.py:162-163)
        std::vector<int64_t> lbs((size_t)m), order;
        std::vector<cell_t> cells((size_t)m), cells2((size_t)m);
        PFD_CUDA(h, cudaMemcpyAsync(lbs.data(), h->sub_labels.p, (size_t)m * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
        PFD_CUDA(h, cudaMemcpyAsync(cells.data(), h->sub_idxs.p, (size_t)m * sizeof(cell_t), cudaMemcpyDeviceToHost, h->stream));
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
        numba_argsort_i64(lbs, order);
        std::vector<int64_t> lbs2((size_t)m);
        for (int64_t i = 0; i < m; ++i) {
            lbs2[(size_t)i] = lbs[(size_t)order[(size_t)i]];
            cells2[(size_t)i] = cells[(size_t)order[(size_t)i]];
        }
        PFD_CUDA(h, cudaMemcpyAsync(h->sub_labels.p, lbs2.data(), (size_t)m * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
        PFD_CUDA(h, cudaMemcpyAsync(h->sub_idxs.p, cells2.data(), (size_t)m * sizeof(cell_t), cudaMemcpyHostToDevice, h->stream));
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    h->have_sub_labels = true;
    stage_collect(h);
    if (n_out) *n_out = m;
    return PFD_OK;
}

#define PFD_MAX_REGION_LABEL (1ll << 28)

template <typename T>
static int region_slices_typed(pfd_handle* h, const void* reg_dev) {
    const int64_t n = h->n;
    unsigned long long* mx = (unsigned long long*)h->counters.p + 6;
    PFD_CUDA(h, cudaMemsetAsync(mx, 0, sizeof(unsigned long long), h->stream));
    max_label_kernel<T><<<grid_for(n, 256, 8, 148 * 8), 256, 0, h->stream>>>((const T*)reg_dev, n, mx);
    PFD_LAUNCH_CHECK(h);
    unsigned long long hmx = 0;
    PFD_CUDA(h, cudaMemcpyAsync(&hmx, mx, sizeof(hmx), cudaMemcpyDeviceToHost, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    h->n_sub = 0;
    if (hmx == 0) return PFD_OK;
    if (hmx > (unsigned long long)PFD_MAX_REGION_LABEL)
        return pfd_fail(h, PFD_ERR_UNSUPPORTED, "pfd_region_slices: labels above 2^28 are not supported");
    const int64_t nlab = (int64_t)hmx;
    PFD_TRY(pfd_reserve(h, h->scratch[2], (size_t)nlab * sizeof(int4)));
    int4* box = (int4*)h->scratch[2].p;
    region_box_init_kernel<<<grid_for(nlab, 256), 256, 0, h->stream>>>(box, nlab);
    PFD_LAUNCH_CHECK(h);
    region_box_kernel<T><<<grid_for(n, 256, 4), 256, 0, h->stream>>>((const T*)reg_dev, h->nrow, h->ncol, box);
    PFD_LAUNCH_CHECK(h);
    BoxPresentPred pred{box};
    PFD_TRY((compact_cells<BoxPresentPred, 2, uint32_t>(h, nullptr, nlab, pred, nullptr)));  // present labels, ascending
    PFD_TRY(pfd_reserve(h, h->sub_labels, (size_t)std::max<int64_t>(h->n_sub, 1) * sizeof(int64_t)));
    PFD_TRY(pfd_reserve(h, h->sub_slices, (size_t)std::max<int64_t>(h->n_sub, 1) * sizeof(int4)));
    if (h->n_sub > 0) {
        region_slices_kernel<<<grid_for(h->n_sub, 256), 256, 0, h->stream>>>((const cell_t*)h->sub_idxs.p, h->n_sub, box,
                                                                           (int64_t*)h->sub_labels.p, (int4*)h->sub_slices.p);
        PFD_LAUNCH_CHECK(h);
    }
    return PFD_OK;
}

extern "C" int pfd_region_slices(pfd_handle* h, const void* regions, int dtype, int64_t* n_labels) {
    PFD_TRY(require_raster(h, "pfd_region_slices"));
    stage_reset(h);
    if (!regions) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_region_slices: null array");
    if (!region_dtype_ok(dtype)) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_region_slices: regions must be a 32/64-bit integer array");
    if (h->nrow >= (1ll << 31) || h->ncol >= (1ll << 31)) return pfd_fail(h, PFD_ERR_UNSUPPORTED, "pfd_region_slices: raster side above 2^31");
    const void* reg_dev = nullptr;
    PFD_TRY(pfd_stage_in(h, regions, (size_t)h->n * pfd_dtype_size(dtype), 5, &reg_dev));
    PFD_TRY(pfd_reserve(h, h->counters, 8 * sizeof(unsigned long long)));
    h->have_sub_labels = h->have_sub_slices = false;
    int rc;
    switch (dtype) {
    case PFD_I32: rc = region_slices_typed<int32_t>(h, reg_dev); break;
    case PFD_U32: rc = region_slices_typed<uint32_t>(h, reg_dev); break;
    case PFD_I64: rc = region_slices_typed<int64_t>(h, reg_dev); break;
    default: rc = region_slices_typed<uint64_t>(h, reg_dev); break;
    }
    PFD_TRY(rc);
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    h->have_sub_labels = h->have_sub_slices = true;
    stage_collect(h);
    if (n_labels) *n_labels = h->n_sub;
    return PFD_OK;
}

// streams.streams (pyflwdir/streams.py:131-188)
extern "C" int pfd_streams(pfd_handle* h, const uint8_t* mask, int64_t max_len, int64_t* n_streams, int64_t* n_cells) {
    PFD_TRY(require_raster(h, "pfd_streams"));
    stage_reset(h);
    PFD_TRY(order_impl(h, false, false));
    PFD_TRY(ensure_upmask(h));
    const int64_t n = h->n;
    h->n_streams = -1;
    const void* mask_dev = nullptr;
    if (mask) PFD_TRY(pfd_stage_in(h, mask, (size_t)n, 4, &mask_dev));
    PFD_TRY(pfd_reserve(h, h->scratch[1], (size_t)n));
    uint8_t* st = (uint8_t*)h->scratch[1].p;
    PFD_CUDA(h, cudaMemsetAsync(st, 0, (size_t)n, h->stream));
    StreamStartOp op{(const uint8_t*)h->upmask.p, (const uint8_t*)mask_dev, st, h->ncol};
    PFD_TRY((run_sweep<StreamStartOp, true>(h, op, 0)));
    StateBit2Pred pred{st};
    PFD_TRY((number_outlets<StateBit2Pred, 1, uint32_t>(h, pred, nullptr)));  // segments start while walking seq[::-1]
    const int64_t ns = h->n_sub;
    int64_t npieces = 0, ncells = 0;
    PFD_TRY(pfd_reserve(h, h->stream_off, sizeof(long long)));
    if (ns > 0) {
        // per start: len | npiece | ncell (uint32 each), then the two exclusive scans (uint64 [ns + 1] each)
        PFD_TRY(pfd_reserve(h, h->scratch[0], (size_t)ns * 3 * sizeof(uint32_t)));
        PFD_TRY(pfd_reserve(h, h->scratch[2], (size_t)(ns + 1) * 2 * sizeof(unsigned long long)));
        uint32_t* len = (uint32_t*)h->scratch[0].p;
        uint32_t* npiece = len + ns;
        uint32_t* ncell = npiece + ns;
        unsigned long long* piece_off = (unsigned long long*)h->scratch[2].p;
        unsigned long long* cell_off = piece_off + (ns + 1);
        const int g = grid_for(ns, 128, 1, 148 * 16);
        stream_count_kernel<<<g, 128, 0, h->stream>>>((const uint8_t*)h->dir.p, (const uint8_t*)h->upmask.p, (const uint8_t*)mask_dev,
                                                     (const cell_t*)h->sub_idxs.p, ns, h->ncol, max_len, len, npiece, ncell);
        PFD_LAUNCH_CHECK(h);
        scan_counts_kernel<<<1, 1024, 0, h->stream>>>(npiece, ns, piece_off);
        PFD_LAUNCH_CHECK(h);
        scan_counts_kernel<<<1, 1024, 0, h->stream>>>(ncell, ns, cell_off);
        PFD_LAUNCH_CHECK(h);
        unsigned long long tot[2] = {0, 0};
        PFD_CUDA(h, cudaMemcpyAsync(&tot[0], piece_off + ns, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
        PFD_CUDA(h, cudaMemcpyAsync(&tot[1], cell_off + ns, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
        npieces = (int64_t)tot[0];
        ncells = (int64_t)tot[1];
        PFD_TRY(pfd_reserve(h, h->stream_off, (size_t)(npieces + 1) * sizeof(long long)));
        PFD_TRY(pfd_reserve(h, h->stream_cells, (size_t)std::max<int64_t>(ncells, 1) * sizeof(cell_t)));
        stream_write_kernel<<<g, 128, 0, h->stream>>>((const uint8_t*)h->dir.p, (const cell_t*)h->sub_idxs.p, ns, h->ncol, max_len, len,
                                                     piece_off, cell_off, (long long*)h->stream_off.p, (cell_t*)h->stream_cells.p);
        PFD_LAUNCH_CHECK(h);
    }
    const long long last = (long long)ncells;
    PFD_CUDA(h, cudaMemcpyAsync((long long*)h->stream_off.p + npieces, &last, sizeof(last), cudaMemcpyHostToDevice, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    h->n_streams = npieces;
    h->n_stream_cells = ncells;
    stage_collect(h);
    if (n_streams) *n_streams = npieces;
    if (n_cells) *n_cells = ncells;
    return PFD_OK;
}

// basins.subbasins_pfafstetter (pyflwdir/basins.py:106-191)
struct TmpBufs {  // device scratch that lives for one call
    std::vector<DevBuf*> all;
    ~TmpBufs() {
        for (DevBuf* b : all) {
            pfd_release(*b);
            delete b;
        }
    }
    DevBuf* get() {
        all.push_back(new DevBuf());
        return all.back();
    }
};

extern "C" int pfd_subbasins_pfafstetter(pfd_handle* h, const void* idxs_us_main, int idx_dtype, const void* uparea, int dtype,
                                         const uint8_t* mask, int depth, int64_t* pfafbas_out, int64_t* n_outlets) {
    PFD_TRY(require_raster(h, "pfd_subbasins_pfafstetter"));
    stage_reset(h);
    if (!idxs_us_main || !uparea || !pfafbas_out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_subbasins_pfafstetter: null array");
    const size_t isz = pfd_dtype_size(idx_dtype), esz = pfd_dtype_size(dtype);
    if ((isz != 4 && isz != 8) || idx_dtype == PFD_F32 || idx_dtype == PFD_F64)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_subbasins_pfafstetter: index dtype must be a 32/64-bit integer");
    if (dtype != PFD_I32 && dtype != PFD_I64 && dtype != PFD_F32 && dtype != PFD_F64)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_subbasins_pfafstetter: uparea must be int32, int64, float32 or float64");
    if (depth < 1 || depth > 8) return pfd_fail(h, PFD_ERR_UNSUPPORTED, "pfd_subbasins_pfafstetter: depth must be in 1..8");
    PFD_TRY(order_impl(h, false, false));
    PFD_TRY(ensure_upmask(h));
    const int64_t n = h->n;
    void* out_dev = nullptr;
    PFD_TRY(pfd_stage_out(h, pfafbas_out, (size_t)n * sizeof(int64_t), 3, &out_dev));
    const void *main_dev = nullptr, *upa_dev = nullptr, *mask_dev = nullptr;
    PFD_TRY(pfd_stage_in(h, idxs_us_main, (size_t)n * isz, 5, &main_dev));
    PFD_TRY(pfd_stage_in(h, uparea, (size_t)n * esz, 4, &upa_dev));
    if (mask) PFD_TRY(pfd_stage_in(h, mask, (size_t)n, 2, &mask_dev));
    PFD_TRY(pfd_reserve(h, h->counters, 8 * sizeof(unsigned long long)));
    unsigned int* flag = reinterpret_cast<unsigned int*>((unsigned long long*)h->counters.p + 5);
    PFD_CUDA(h, cudaMemsetAsync(flag, 0, sizeof(unsigned int), h->stream));
    TmpBufs tmp;
    DevBuf *b_so = tmp.get(), *b_pb = tmp.get(), *b_in = tmp.get(), *b_pos = tmp.get(), *b_upa = tmp.get(), *b_main = tmp.get();
    PFD_TRY(pfd_reserve(h, *b_so, (size_t)n));
    PFD_TRY(pfd_reserve(h, *b_pb, (size_t)n * sizeof(int32_t)));
    PFD_TRY(pfd_reserve(h, *b_in, (size_t)n));
    PFD_TRY(pfd_reserve(h, *b_pos, (size_t)n * sizeof(uint32_t)));
    PFD_TRY(pfd_reserve(h, *b_upa, (size_t)n * sizeof(double)));
    const int g = grid_for(n, 256, 4);
    // classic stream order of the masked network, orders above depth + 1 dropped (basins.py:120-121)
    uint8_t* so = (uint8_t*)b_so->p;
    PFD_CUDA(h, cudaMemsetAsync(so, 0, (size_t)n, h->stream));
    {
        const uint8_t *dir = (const uint8_t*)h->dir.p, *upm = (const uint8_t*)h->upmask.p;
        int rc;
        if (idx_dtype == PFD_I32) {
            ClassicOrderOp<int32_t> op{dir, upm, (const uint8_t*)mask_dev, (const int32_t*)main_dev, so, h->ncol};
            rc = run_sweep<ClassicOrderOp<int32_t>, false>(h, op, 0);
        } else if (idx_dtype == PFD_U32) {
            ClassicOrderOp<uint32_t> op{dir, upm, (const uint8_t*)mask_dev, (const uint32_t*)main_dev, so, h->ncol};
            rc = run_sweep<ClassicOrderOp<uint32_t>, false>(h, op, 0);
        } else {
            ClassicOrderOp<int64_t> op{dir, upm, (const uint8_t*)mask_dev, (const int64_t*)main_dev, so, h->ncol};
            rc = run_sweep<ClassicOrderOp<int64_t>, false>(h, op, 0);
        }
        PFD_TRY(rc);
    }
    pf_prepare_kernel<<<g, 256, 0, h->stream>>>(so, n, depth);
    PFD_LAUNCH_CHECK(h);
    PFD_CUDA(h, cudaMemsetAsync(b_pos->p, 0, (size_t)n * sizeof(uint32_t), h->stream));
    pf_pos_kernel<<<grid_for(h->nnodes, 256, 4), 256, 0, h->stream>>>((const cell_t*)h->seq.p, h->nnodes, (uint32_t*)b_pos->p);
    PFD_LAUNCH_CHECK(h);
    const uint32_t* main32 = (const uint32_t*)main_dev;  // int32 -1 and uint32 mv are both 0xFFFFFFFF
    if (isz == 8) {
        PFD_TRY(pfd_reserve(h, *b_main, (size_t)n * sizeof(uint32_t)));
        pf_main32_kernel<int64_t><<<g, 256, 0, h->stream>>>((const int64_t*)main_dev, n, (uint32_t*)b_main->p, flag);
        PFD_LAUNCH_CHECK(h);
        main32 = (const uint32_t*)b_main->p;
    }
    switch (dtype) {
    case PFD_I32: pf_to_double_kernel<int32_t><<<g, 256, 0, h->stream>>>((const int32_t*)upa_dev, n, (double*)b_upa->p); break;
    case PFD_I64: pf_to_double_kernel<int64_t><<<g, 256, 0, h->stream>>>((const int64_t*)upa_dev, n, (double*)b_upa->p); break;
    case PFD_F32: pf_to_double_kernel<float><<<g, 256, 0, h->stream>>>((const float*)upa_dev, n, (double*)b_upa->p); break;
    default: pf_to_double_kernel<double><<<g, 256, 0, h->stream>>>((const double*)upa_dev, n, (double*)b_upa->p); break;
    }
    PFD_LAUNCH_CHECK(h);
    PFD_CUDA(h, cudaMemsetAsync(b_pb->p, 0, (size_t)n * sizeof(int32_t), h->stream));
    PFD_CUDA(h, cudaMemsetAsync(b_in->p, 0, (size_t)n, h->stream));
    PfGraph G{(const uint8_t*)h->dir.p, (const uint8_t*)h->upmask.p, so, main32, (const uint32_t*)b_pos->p, (const double*)b_upa->p,
              (int32_t*)b_pb->p, (uint8_t*)b_in->p, h->ncol, depth};

    // round 0: one label per pit (basins.py:131-141)
    int64_t nlab = h->n_pits, nout_total = h->n_pits;
    DevBuf *b_lab = tmp.get(), *b_labo = tmp.get(), *b_outl = tmp.get();
    PFD_TRY(pfd_reserve(h, *b_lab, (size_t)std::max<int64_t>(nlab, 1) * sizeof(long long)));
    PFD_TRY(pfd_reserve(h, *b_labo, (size_t)std::max<int64_t>(nlab, 1) * sizeof(uint32_t)));
    PFD_TRY(pfd_reserve(h, *b_outl, (size_t)std::max<int64_t>(nout_total, 1) * sizeof(cell_t)));
    if (nlab > 0) {
        pf_init_pits_kernel<<<grid_for(nlab, 128), 128, 0, h->stream>>>(G, (const cell_t*)h->pits.p, nlab, (long long*)b_lab->p,
                                                                       (uint32_t*)b_labo->p, (cell_t*)b_outl->p);
        PFD_LAUNCH_CHECK(h);
    }
    for (int d0 = 1; d0 <= depth && nlab > 0; ++d0) {
        const int gl = grid_for(nlab, 64, 1, 148 * 32);
        DevBuf *b_cnt = tmp.get(), *b_off = tmp.get();
        PFD_TRY(pfd_reserve(h, *b_cnt, (size_t)nlab * sizeof(uint32_t)));
        PFD_TRY(pfd_reserve(h, *b_off, (size_t)(nlab + 1) * sizeof(unsigned long long)));
        pf_count_kernel<<<gl, 64, 0, h->stream>>>(G, (const long long*)b_lab->p, (const uint32_t*)b_labo->p, nlab, (uint32_t*)b_cnt->p);
        PFD_LAUNCH_CHECK(h);
        scan_counts_kernel<<<1, 1024, 0, h->stream>>>((const uint32_t*)b_cnt->p, nlab, (unsigned long long*)b_off->p);
        PFD_LAUNCH_CHECK(h);
        unsigned long long ncand = 0;
        PFD_CUDA(h, cudaMemcpyAsync(&ncand, (unsigned long long*)b_off->p + nlab, sizeof(ncand), cudaMemcpyDeviceToHost, h->stream));
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
        if (ncand == 0) break;
        DevBuf *b_cand = tmp.get(), *b_key = tmp.get(), *b_R = tmp.get(), *b_cl = tmp.get(), *b_co = tmp.get(), *b_nc = tmp.get(),
               *b_oc = tmp.get(), *b_no = tmp.get(), *b_coff = tmp.get(), *b_ooff = tmp.get();
        PFD_TRY(pfd_reserve(h, *b_cand, (size_t)ncand * sizeof(uint32_t)));
        PFD_TRY(pfd_reserve(h, *b_key, (size_t)ncand * sizeof(double)));
        PFD_TRY(pfd_reserve(h, *b_R, (size_t)ncand * sizeof(uint32_t)));
        PFD_TRY(pfd_reserve(h, *b_cl, (size_t)nlab * 8 * sizeof(long long)));
        PFD_TRY(pfd_reserve(h, *b_co, (size_t)nlab * 8 * sizeof(uint32_t)));
        PFD_TRY(pfd_reserve(h, *b_nc, (size_t)nlab * sizeof(uint32_t)));
        PFD_TRY(pfd_reserve(h, *b_oc, (size_t)nlab * 8 * sizeof(uint32_t)));
        PFD_TRY(pfd_reserve(h, *b_no, (size_t)nlab * sizeof(uint32_t)));
        PFD_TRY(pfd_reserve(h, *b_coff, (size_t)(nlab + 1) * sizeof(unsigned long long)));
        PFD_TRY(pfd_reserve(h, *b_ooff, (size_t)(nlab + 1) * sizeof(unsigned long long)));
        pf_process_kernel<<<gl, 64, 0, h->stream>>>(G, (const long long*)b_lab->p, (const uint32_t*)b_labo->p, nlab, d0,
                                                   (const unsigned long long*)b_off->p, (uint32_t*)b_cand->p, (double*)b_key->p,
                                                   (uint32_t*)b_R->p, (long long*)b_cl->p, (uint32_t*)b_co->p, (uint32_t*)b_nc->p,
                                                   (uint32_t*)b_oc->p, (uint32_t*)b_no->p, flag);
        PFD_LAUNCH_CHECK(h);
        scan_counts_kernel<<<1, 1024, 0, h->stream>>>((const uint32_t*)b_nc->p, nlab, (unsigned long long*)b_coff->p);
        PFD_LAUNCH_CHECK(h);
        scan_counts_kernel<<<1, 1024, 0, h->stream>>>((const uint32_t*)b_no->p, nlab, (unsigned long long*)b_ooff->p);
        PFD_LAUNCH_CHECK(h);
        unsigned long long tot[2] = {0, 0};
        PFD_CUDA(h, cudaMemcpyAsync(&tot[0], (unsigned long long*)b_coff->p + nlab, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
        PFD_CUDA(h, cudaMemcpyAsync(&tot[1], (unsigned long long*)b_ooff->p + nlab, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
        const int64_t nnext = (int64_t)tot[0], nadd = (int64_t)tot[1];
        DevBuf *b_nlab = tmp.get(), *b_nlabo = tmp.get(), *b_noutl = tmp.get();
        PFD_TRY(pfd_reserve(h, *b_nlab, (size_t)std::max<int64_t>(nnext, 1) * sizeof(long long)));
        PFD_TRY(pfd_reserve(h, *b_nlabo, (size_t)std::max<int64_t>(nnext, 1) * sizeof(uint32_t)));
        PFD_TRY(pfd_reserve(h, *b_noutl, (size_t)(nout_total + nadd + 1) * sizeof(cell_t)));
        PFD_CUDA(h, cudaMemcpyAsync(b_noutl->p, b_outl->p, (size_t)nout_total * sizeof(cell_t), cudaMemcpyDeviceToDevice, h->stream));
        pf_gather_kernel<<<gl, 64, 0, h->stream>>>(nlab, (const long long*)b_cl->p, (const uint32_t*)b_co->p, (const uint32_t*)b_nc->p,
                                                  (const unsigned long long*)b_coff->p, (const uint32_t*)b_oc->p, (const uint32_t*)b_no->p,
                                                  (const unsigned long long*)b_ooff->p, (long long*)b_nlab->p, (uint32_t*)b_nlabo->p,
                                                  (cell_t*)b_noutl->p + nout_total);
        PFD_LAUNCH_CHECK(h);
        b_lab = b_nlab;
        b_labo = b_nlabo;
        b_outl = b_noutl;
        nlab = nnext;
        nout_total += nadd;
    }
    // pfafbas = core.fillnodata_upstream(idxs_ds, seq, pfaf_branch, 0) % 10**depth (basins.py:190)
    {
        FillUpOp<int32_t> fill{(const uint8_t*)h->dir.p, (int32_t*)b_pb->p, h->ncol};
        PFD_TRY((run_sweep<FillUpOp<int32_t>, false>(h, fill, 0)));
    }
    long long mod = 1;
    for (int i = 0; i < depth; ++i) mod *= 10;
    pf_mod_kernel<<<g, 256, 0, h->stream>>>((const int32_t*)b_pb->p, n, mod, (int64_t*)out_dev);
    PFD_LAUNCH_CHECK(h);
    h->have_sub_labels = h->have_sub_slices = false;
    h->n_sub = nout_total;
    PFD_TRY(pfd_reserve(h, h->sub_idxs, (size_t)std::max<int64_t>(nout_total, 1) * sizeof(cell_t)));
    PFD_CUDA(h, cudaMemcpyAsync(h->sub_idxs.p, b_outl->p, (size_t)nout_total * sizeof(cell_t), cudaMemcpyDeviceToDevice, h->stream));
    unsigned int hflag = 0;
    PFD_CUDA(h, cudaMemcpyAsync(&hflag, flag, sizeof(hflag), cudaMemcpyDeviceToHost, h->stream));
    PFD_TRY(pfd_finish_out(h, pfafbas_out, out_dev, (size_t)n * sizeof(int64_t)));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    if (hflag & 8u) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_subbasins_pfafstetter: idxs_us_main holds an index outside the raster");
    if (hflag & 32u)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_subbasins_pfafstetter: idxs_us_main misses the main upstream cell of a confluence");
    if (n_outlets) *n_outlets = nout_total;
    return PFD_OK;
}

// rivers.classify_estuary (pyflwdir/rivers.py:11-53). est_init: N int8 with 1 at the estuary outlets (pits with
// elevtn <= max_elevtn; selected by the caller), 0 elsewhere.
template <typename TD, typename TW>
static int estuary_typed(pfd_handle* h, const void* dst_dev, const void* wth_dev, double min_convergence, int8_t* est) {
    EstuaryOp<TD, TW> op{(const uint8_t*)h->dir.p, (const TD*)dst_dev, (const TW*)wth_dev, est, min_convergence, h->ncol};
    return run_sweep<EstuaryOp<TD, TW>, false>(h, op, 1);
}
extern "C" int pfd_classify_estuary(pfd_handle* h, const int8_t* est_init, const void* rivdst, int dst_dtype, const void* rivwth,
                                    int wth_dtype, double min_convergence, int8_t* out) {
    PFD_TRY(require_raster(h, "pfd_classify_estuary"));
    stage_reset(h);
    if (!est_init || !rivdst || !rivwth || !out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_classify_estuary: null array");
    if ((dst_dtype != PFD_F32 && dst_dtype != PFD_F64) || (wth_dtype != PFD_F32 && wth_dtype != PFD_F64))
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_classify_estuary: rivdst / rivwth must be float32 or float64");
    PFD_TRY(order_impl(h, false, false));
    const int64_t n = h->n;
    void* out_dev = nullptr;
    PFD_TRY(pfd_stage_out(h, out, (size_t)n, 3, &out_dev));
    const void *dst_dev = nullptr, *wth_dev = nullptr;
    PFD_TRY(pfd_stage_in(h, rivdst, (size_t)n * pfd_dtype_size(dst_dtype), 5, &dst_dev));
    PFD_TRY(pfd_stage_in(h, rivwth, (size_t)n * pfd_dtype_size(wth_dtype), 4, &wth_dev));
    if (out_dev != (const void*)est_init) PFD_CUDA(h, cudaMemcpyAsync(out_dev, est_init, (size_t)n, cudaMemcpyDefault, h->stream));
    int rc;
    if (dst_dtype == PFD_F32)
        rc = wth_dtype == PFD_F32 ? estuary_typed<float, float>(h, dst_dev, wth_dev, min_convergence, (int8_t*)out_dev)
                                  : estuary_typed<float, double>(h, dst_dev, wth_dev, min_convergence, (int8_t*)out_dev);
    else
        rc = wth_dtype == PFD_F32 ? estuary_typed<double, float>(h, dst_dev, wth_dev, min_convergence, (int8_t*)out_dev)
                                  : estuary_typed<double, double>(h, dst_dev, wth_dev, min_convergence, (int8_t*)out_dev);
    PFD_TRY(rc);
    PFD_TRY(pfd_finish_out(h, out, out_dev, (size_t)n));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    stage_collect(h);
    return PFD_OK;
}

// ---------------------------------------------------------------------------------------------------------
// synthetic input
// ---------------------------------------------------------------------------------------------------------
__global__ void synth_elevation_kernel(int64_t nrow, int64_t ncol, int64_t nref, int octaves, uint32_t seed, float* __restrict__ z,
                                       int64_t row_off = 0) {
    const int64_t n = nrow * ncol;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        z[i] = pfd_synth_z(i / ncol + row_off, i % ncol, nref, octaves, seed);
}

__global__ void synth_d8_kernel(const float* __restrict__ z, int64_t nrow, int64_t ncol, float sea_level, uint8_t* __restrict__ d8) {
    const int64_t n = nrow * ncol;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / ncol, c = i % ncol;
        float w[9];
        int valid[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int64_t rr = r + k / 3 - 1, cc = c + k % 3 - 1;
            valid[k] = (rr >= 0 && rr < nrow && cc >= 0 && cc < ncol);
            w[k] = valid[k] ? __ldg(z + rr * ncol + cc) : 0.0f;
        }
        d8[i] = pfd_synth_d8_from_window(w, valid, sea_level);
    }
}

extern "C" int pfd_synth_elevation(pfd_handle* h, int64_t nrow, int64_t ncol, int64_t nref, int octaves, uint32_t seed, float* z_out) {
    PFD_TRY(check_handle(h));
    PFD_TRY(check_shape(h, nrow, ncol, "pfd_synth_elevation"));
    if (!z_out || nref < 8 || octaves < 1) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_synth_elevation: bad argument");
    const int64_t n = nrow * ncol;
    void* dev = nullptr;
    PFD_TRY(pfd_stage_out(h, z_out, (size_t)n * 4, 3, &dev));
    synth_elevation_kernel<<<grid_for(n, 256, 2), 256, 0, h->stream>>>(nrow, ncol, nref, octaves, seed, (float*)dev);
    PFD_LAUNCH_CHECK(h);
    PFD_TRY(pfd_finish_out(h, z_out, dev, (size_t)n * 4));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    return PFD_OK;
}

extern "C" int pfd_synth_d8(pfd_handle* h, const float* z, int64_t nrow, int64_t ncol, float sea_level, uint8_t* d8_out) {
    PFD_TRY(check_handle(h));
    PFD_TRY(check_shape(h, nrow, ncol, "pfd_synth_d8"));
    if (!z || !d8_out) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_synth_d8: null array");
    const int64_t n = nrow * ncol;
    const void* zdev = nullptr;
    PFD_TRY(pfd_stage_in(h, z, (size_t)n * 4, 4, &zdev));
    void* dev = nullptr;
    PFD_TRY(pfd_stage_out(h, d8_out, (size_t)n, 3, &dev));
    synth_d8_kernel<<<grid_for(n, 256, 2), 256, 0, h->stream>>>((const float*)zdev, nrow, ncol, sea_level, (uint8_t*)dev);
    PFD_LAUNCH_CHECK(h);
    PFD_TRY(pfd_finish_out(h, d8_out, dev, (size_t)n));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    return PFD_OK;
}

// D8 codes of rows [row0, row0 + nrow) of a synthetic raster with nrow_global rows (row blocks of the multi-GPU
// bench are generated in place, rank by rank; identical to the same rows of the single-piece generator)
extern "C" int pfd_synth_d8_block(pfd_handle* h, int64_t row0, int64_t nrow, int64_t ncol, int64_t nrow_global, int64_t nref,
                                  int octaves, uint32_t seed, float sea_level, uint8_t* d8_out) {
    PFD_TRY(check_handle(h));
    PFD_TRY(check_shape(h, nrow, ncol, "pfd_synth_d8_block"));
    if (!d8_out || row0 < 0 || row0 + nrow > nrow_global || nref < 8 || octaves < 1)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_synth_d8_block: bad argument");
    const int64_t top = row0 > 0 ? 1 : 0, bot = (row0 + nrow < nrow_global) ? 1 : 0;
    const int64_t zrows = nrow + top + bot;
    PFD_TRY(pfd_reserve(h, h->scratch[4], (size_t)(zrows * ncol) * sizeof(float)));
    PFD_TRY(pfd_reserve(h, h->scratch[5], (size_t)(zrows * ncol)));
    float* z = (float*)h->scratch[4].p;
    uint8_t* tmp = (uint8_t*)h->scratch[5].p;
    synth_elevation_kernel<<<grid_for(zrows * ncol, 256, 2), 256, 0, h->stream>>>(zrows, ncol, nref, octaves, seed, z, row0 - top);
    PFD_LAUNCH_CHECK(h);
    // steepest descent on the extended block: its first / last rows are only correct when they are true raster
    // edges, which is exactly when they are owned rows
    synth_d8_kernel<<<grid_for(zrows * ncol, 256, 2), 256, 0, h->stream>>>(z, zrows, ncol, sea_level, tmp);
    PFD_LAUNCH_CHECK(h);
    PFD_CUDA(h, cudaMemcpyAsync(d8_out, tmp + top * ncol, (size_t)(nrow * ncol), cudaMemcpyDefault, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    return PFD_OK;
}

// ---------------------------------------------------------------------------------------------------------
// dem.fill_depressions (pyflwdir/dem.py:17-143) -- SURVEY.md §8f-3; kernels in pfd_fill.cuh
// ---------------------------------------------------------------------------------------------------------
#include "pfd_fill.cuh"

extern "C" int pfd_fill_depressions(pfd_handle* h, const void* elevtn, int elev_dtype, int64_t nrow, int64_t ncol, int outlets_mode,
                                    const int64_t* idxs_pit, int64_t n_pit, double nodata, double max_depth, int has_elv_max,
                                    double elv_max, int connectivity, int int_delv, void* elevtn_out, uint8_t* d8_out, int64_t* stats) {
    PFD_TRY(check_handle(h));
    if (!elevtn || !elevtn_out || !d8_out || nrow < 0 || ncol < 0) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_fill_depressions: bad argument");
    if (elev_dtype != PFD_F32 && elev_dtype != PFD_F64)
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_fill_depressions: elevtn must be float32 or float64 (integers: pass float64 with int_delv)");
    if (connectivity != 4 && connectivity != 8) return pfd_fail(h, PFD_ERR_INVALID_ARG, "\"connectivity\" should either be 4 or 8");
    if (outlets_mode < 0 || outlets_mode > 2 || (outlets_mode == 2 && n_pit > 0 && !idxs_pit))
        return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_fill_depressions: outlets_mode is 0 (edge), 1 (min) or 2 (idxs_pit)");
    if (max_depth >= 0)
        return pfd_fail(h, PFD_ERR_UNSUPPORTED,
                        "pfd_fill_depressions: max_depth >= 0 (re-opening of visited cells, dem.py:121-132) is not implemented on the device");
    const int64_t n = nrow * ncol;
    if (n >= (1ll << 31)) return pfd_fail(h, PFD_ERR_UNSUPPORTED, "pfd_fill_depressions: rasters of 2^31 cells or more are not supported");
    if (n == 0) return PFD_OK;
    const size_t esz = pfd_dtype_size(elev_dtype);
    const void* elev_dev = nullptr;
    void* out_dev = nullptr;
    void* d8_dev = nullptr;
    const void* pit_dev = nullptr;
    PFD_TRY(pfd_stage_in(h, elevtn, (size_t)n * esz, 0, &elev_dev));
    PFD_TRY(pfd_stage_out(h, elevtn_out, (size_t)n * esz, 1, &out_dev));
    PFD_TRY(pfd_stage_out(h, d8_out, (size_t)n, 2, &d8_dev));
    if (outlets_mode == 2 && n_pit > 0) PFD_TRY(pfd_stage_in(h, idxs_pit, (size_t)n_pit * sizeof(int64_t), 3, &pit_dev));
    if (elev_dtype == PFD_F32)
        PFD_TRY((fd_fill_impl<float, float>(h, (const float*)elev_dev, nrow, ncol, outlets_mode, (const int64_t*)pit_dev, n_pit, nodata, has_elv_max,
                                            elv_max, connectivity, int_delv, (float*)out_dev, (uint8_t*)d8_dev, stats)));
    else
        PFD_TRY((fd_fill_impl<double, double>(h, (const double*)elev_dev, nrow, ncol, outlets_mode, (const int64_t*)pit_dev, n_pit, nodata,
                                              has_elv_max, elv_max, connectivity, int_delv, (double*)out_dev, (uint8_t*)d8_dev, stats)));
    PFD_TRY(pfd_finish_out(h, elevtn_out, out_dev, (size_t)n * esz));
    PFD_TRY(pfd_finish_out(h, d8_out, d8_dev, (size_t)n));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    return PFD_OK;
}
