// pfd_sweeps.cuh -- level-ordered sweeps over the BFS sequence.
// Replaces streams.accuflux / accuflux_ds (pyflwdir/streams.py:15-41,44-70), core.fillnodata_upstream as used
// by basins.basins (pyflwdir/core.py:120-146, basins.py:12-18), streams.strahler_order (streams.py:228-269)
// and dem.height_above_nearest_drain (pyflwdir/dem.py:299-330).
//
// The reference walks `seq` (down->up) or `seq[::-1]` (up->down) serially. All cells of one rank level are
// independent, so a sweep = for level in order: all cells of the level in parallel. Up-sweeps are PULLS: the
// downstream cell gathers its (complete) upstream neighbours in DESCENDING linear index, which is exactly the
// order in which the reference's `seq[::-1]` loop adds them -> floats are bit-exact and no atomics are needed.
// One persistent cooperative kernel per sweep follows a host-built schedule: big levels use the whole grid +
// grid.sync(), runs of small levels are walked by CTA 0 alone.
#pragma once
#include "pfd_common.cuh"

#ifndef SW_THREADS
#define SW_THREADS 1024
#endif
#ifndef SW_SOLO_MAX
#define SW_SOLO_MAX 4096
#endif

struct SweepSeg {
    int first;   // first level of the segment
    int count;   // number of levels (1 for a big level)
    int solo;    // 1: CTA 0 walks `count` small levels alone
    int pad;
};

struct SweepParams {
    const cell_t* seq;
    const long long* level_off;
    const SweepSeg* segs;
    int nsegs;
    int skip_level0;  // down-sweeps whose level-0 (pit) step is a no-op
};

template <class Op, bool UP>
__global__ void __launch_bounds__(SW_THREADS) sweep_kernel(SweepParams P, Op op) {
    cg::grid_group grid = cg::this_grid();
    for (int si = 0; si < P.nsegs; ++si) {
        const SweepSeg sg = P.segs[UP ? (P.nsegs - 1 - si) : si];
        if (sg.solo) {
            if (blockIdx.x == 0) {
                for (int li = 0; li < sg.count; ++li) {
                    const int lev = UP ? (sg.first + sg.count - 1 - li) : (sg.first + li);
                    if (!(P.skip_level0 && lev == 0)) {
                        const long long s = P.level_off[lev], e = P.level_off[lev + 1];
                        for (long long p = s + threadIdx.x; p < e; p += SW_THREADS) op(__ldg(P.seq + p), p, lev);
                    }
                    __syncthreads();
                }
            }
        } else {
            const int lev = sg.first;
            if (!(P.skip_level0 && lev == 0)) {
                const long long s = P.level_off[lev], e = P.level_off[lev + 1];
                const long long stride = (long long)gridDim.x * SW_THREADS;
                for (long long p = s + (long long)blockIdx.x * SW_THREADS + threadIdx.x; p < e; p += stride)
                    op(__ldg(P.seq + p), p, lev);
            }
        }
        if (si + 1 < P.nsegs) grid.sync();
    }
}

// ---- typed helpers -------------------------------------------------------------------------------------
template <typename T> struct AccT { typedef T U; };
template <> struct AccT<int8_t> { typedef uint8_t U; };
template <> struct AccT<int16_t> { typedef uint16_t U; };
template <> struct AccT<int32_t> { typedef uint32_t U; };
template <> struct AccT<int64_t> { typedef uint64_t U; };

template <typename T>
__device__ __forceinline__ T acc_add(T a, T b) {  // integer: wrap-around like numba
    typedef typename AccT<T>::U U;
    return (T)((U)a + (U)b);
}
template <>
__device__ __forceinline__ float acc_add<float>(float a, float b) { return __fadd_rn(a, b); }
template <>
__device__ __forceinline__ double acc_add<double>(double a, double b) { return __dadd_rn(a, b); }

struct NoData {
    double f;
    long long i;
    int is_int;
};

// numba's promotion for `x != nodata`: int vs int -> int64 compare, otherwise float64 compare
template <typename T>
__device__ __forceinline__ bool not_nodata(T x, const NoData& nd) {
    return nd.is_int ? ((long long)x != nd.i) : ((double)x != nd.f);
}
template <>
__device__ __forceinline__ bool not_nodata<float>(float x, const NoData& nd) { return (double)x != nd.f; }
template <>
__device__ __forceinline__ bool not_nodata<double>(double x, const NoData& nd) { return x != nd.f; }

__device__ __forceinline__ long long ds_of(cell_t c, uint32_t d, long long ncol) {
    return (d < 8u) ? (long long)c + pfd_slot_off((int)d, ncol) : (long long)c;
}

// ---- streams.accuflux (up): out pre-initialised with data (accu = data.copy(), streams.py:36) -------------
template <typename T>
struct AccuUpOp {
    const uint8_t* upmask;
    T* out;
    long long ncol;
    NoData nd;
    __device__ __forceinline__ void operator()(cell_t c, long long, int) const {
        uint32_t m = __ldg(upmask + c);
        if (!m) return;
        T acc = ld_cg(out + c);
        while (m) {  // descending slot = descending linear index of the upstream neighbour
            const int k = 31 - __clz(m);
            m ^= 1u << k;
            const T a = ld_cg(out + ((long long)c + pfd_slot_off(k, ncol)));
            if (not_nodata(acc, nd) && not_nodata(a, nd)) acc = acc_add(acc, a);
        }
        out[c] = acc;
    }
};

// ---- streams.accuflux_ds (down) ------------------------------------------------------------------------
template <typename T>
struct AccuDownOp {
    const uint8_t* dir;
    T* out;
    long long ncol;
    NoData nd;
    __device__ __forceinline__ void operator()(cell_t c, long long, int) const {
        const uint32_t d = __ldg(dir + c);
        if (d >= 8u) return;  // pit: idx0 == idx_ds
        const T a = ld_cg(out + ((long long)c + pfd_slot_off((int)d, ncol)));
        const T v = ld_cg(out + c);
        if (not_nodata(a, nd) && not_nodata(v, nd)) out[c] = acc_add(v, a);
    }
};

// ---- core.fillnodata_upstream with nodata = 0 (basins) --------------------------------------------------
template <typename U>
struct FillUpOp {
    const uint8_t* dir;
    U* out;
    long long ncol;
    __device__ __forceinline__ void operator()(cell_t c, long long, int) const {
        if (ld_cg(out + c) != (U)0) return;
        const uint32_t d = __ldg(dir + c);
        if (d >= 8u) return;
        const U a = ld_cg(out + ((long long)c + pfd_slot_off((int)d, ncol)));
        if (a != (U)0) out[c] = a;
    }
};

// ---- core.fillnodata_upstream, any dtype / nodata (core.py:120-146): down-sweep ---------------------------
template <typename T>
struct FillUpGenericOp {
    const uint8_t* dir;
    T* out;
    long long ncol;
    NoData nd;
    __device__ __forceinline__ void operator()(cell_t c, long long, int) const {
        const uint32_t d = __ldg(dir + c);
        if (d >= 8u) return;  // pit: idx_ds == idx0, nothing to copy
        if (not_nodata(ld_cg(out + c), nd)) return;
        const T a = ld_cg(out + ((long long)c + pfd_slot_off((int)d, ncol)));
        if (not_nodata(a, nd)) out[c] = a;
    }
};

// ---- core.fillnodata_downstream (core.py:149-188): up-sweep, pull in descending upstream index ------------
// HOW: 0 = max, 1 = min, 2 = sum. A cell is filled only if its ORIGINAL value is nodata; it is written by its own
// step only, so out[c] still holds the original value when its turn comes.
template <typename T, int HOW>
struct FillDownOp {
    const uint8_t* upmask;
    T* out;
    long long ncol;
    NoData nd;
    __device__ __forceinline__ void operator()(cell_t c, long long, int) const {
        uint32_t m = __ldg(upmask + c);
        if (!m) return;
        T acc = ld_cg(out + c);
        if (not_nodata(acc, nd)) return;
        bool changed = false;
        while (m) {
            const int k = 31 - __clz(m);
            m ^= 1u << k;
            const T v = ld_cg(out + ((long long)c + pfd_slot_off(k, ncol)));
            if (!not_nodata(v, nd)) continue;
            if (!not_nodata(acc, nd)) acc = v;
            else if (HOW == 0) acc = (v > acc) ? v : acc;   // max(data_out[idx0], data_out[idx_ds])
            else if (HOW == 1) acc = (v < acc) ? v : acc;
            else acc = acc_add(acc, v);
            changed = true;
        }
        if (changed) out[c] = acc;
    }
};

// ---- streams.stream_order ("classic" / Hack, streams.py:191-225): down-sweep -----------------------------
template <typename IDX>
struct ClassicOrderOp {
    const uint8_t* dir;
    const uint8_t* upmask;
    const uint8_t* mask;  // may be null
    const IDX* us_main;
    uint8_t* out;
    long long ncol;
    __device__ __forceinline__ void operator()(cell_t c, long long, int) const {
        if (mask && !__ldg(mask + c)) return;
        const uint32_t d = __ldg(dir + c);
        if (d >= 8u) {
            out[c] = 1;
            return;
        }
        const long long ds = (long long)c + pfd_slot_off((int)d, ncol);
        uint32_t m = __ldg(upmask + ds);
        int nup = 0;  // core.upstream_count(mask=mask): upstream cells of ds that are in the mask
        if (!mask) nup = __popc(m);
        else
            while (m) {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                nup += __ldg(mask + (ds + pfd_slot_off(k, ncol))) ? 1 : 0;
            }
        const uint8_t sds = ld_cg(out + ds);
        const bool side = nup > 1 && (long long)__ldg(us_main + ds) != (long long)c;
        out[c] = side ? (uint8_t)(sds + 1) : sds;
    }
};

// core.main_upstream (core.py:191-219): upstream neighbour with the largest uparea (> upa_min), first wins on ties
template <typename T, typename IDX>
__global__ void main_upstream_kernel(const uint8_t* __restrict__ upmask, const T* __restrict__ uparea, int64_t n,
                                     long long ncol, T upa_min, IDX* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t m = upmask[i];
        T best = upa_min;
        IDX arg = (IDX)-1;
        while (m) {  // ascending upstream index = the reference's scan order
            const int k = __ffs(m) - 1;
            m &= m - 1;
            const int64_t u = i + pfd_slot_off(k, ncol);
            const T v = uparea[u];
            if (v > best) {
                best = v;
                arg = (IDX)u;
            }
        }
        out[i] = arg;
    }
}

// core.upstream_count with a mask (core.py:50-61): upstream cells that are in the mask; -9 on nodata
__global__ void upstream_count_mask_kernel(const uint8_t* __restrict__ dir, const uint8_t* __restrict__ upmask,
                                           const uint8_t* __restrict__ mask, int64_t n, long long ncol,
                                           int8_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (dir[i] == PFD_DIR_NODATA) {
            out[i] = -9;
            continue;
        }
        uint32_t m = upmask[i];
        int cnt = 0;
        while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            cnt += mask[i + pfd_slot_off(k, ncol)] ? 1 : 0;
        }
        out[i] = (int8_t)cnt;
    }
}

// ---- streams.stream_distance (streams.py:272-315, interpreted Python in the reference): down-sweep -------------
// REAL: float32 distances, the length of a hop comes from a host-built table indexed by (row of the cell, row
// delta + 1, |column delta|) holding float32(gis_utils.distance(...)) -- NumPy-2 scalar semantics: the Python float
// is cast to float32, then added in float32. !REAL: int32 cell counts.
template <bool REAL>
struct StreamDistOp {
    const uint8_t* dir;
    const uint8_t* mask;   // may be null: distance to the outlet
    const float* hop;      // [nrow][3][2], REAL only
    void* out;
    long long ncol;
    __device__ __forceinline__ void operator()(cell_t c, long long, int) const {
        const uint32_t d = __ldg(dir + c);
        const bool stop = d >= 8u || (mask && __ldg(mask + c));
        if (REAL) {
            float* o = (float*)out;
            if (stop) {
                o[c] = 0.0f;
                return;
            }
            const long long ds = (long long)c + pfd_slot_off((int)d, ncol);
            const long long r0 = (long long)c / ncol;
            const int dr = pfd_slot_dr((int)d), dc = pfd_slot_dc((int)d);
            const float len = __ldg(hop + (r0 * 3 + (dr + 1)) * 2 + (dc != 0 ? 1 : 0));
            o[c] = __fadd_rn(ld_cg(o + ds), len);
        } else {
            int32_t* o = (int32_t*)out;
            if (stop) {
                o[c] = 0;
                return;
            }
            o[c] = ld_cg(o + ((long long)c + pfd_slot_off((int)d, ncol))) + 1;
        }
    }
};

template <typename T>
__device__ __forceinline__ T elev_sub(T a, T b);

// ---- dem.floodplains (dem.py:333-379, interpreted Python in the reference): down-sweep ------------------------
// drainh arrives pre-loaded by the host with float32(uparea ** b) at the drain cells (uparea >= upa_min) and -9999
// elsewhere (pow is evaluated on the host exactly as the reference does); drainz is scratch.
template <typename T>
struct FloodplainOp {
    const uint8_t* dir;
    const T* elevtn;
    float* drainh;
    float* drainz;
    int8_t* out;
    long long ncol;
    __device__ __forceinline__ void operator()(cell_t c, long long, int) const {
        const float h_own = drainh[c];
        if (h_own != -9999.0f) {  // drain cell
            drainz[c] = (float)__ldg(elevtn + c);
            out[c] = 1;
            return;
        }
        int8_t f = 0;
        const uint32_t d = __ldg(dir + c);
        const long long ds = ds_of(c, d, ncol);
        if (ds != (long long)c && ld_cg(out + ds) == 1) {
            const float z0 = ld_cg(drainz + ds), h0 = ld_cg(drainh + ds);
            const T dh = elev_sub<T>(__ldg(elevtn + c), (T)z0);  // elevtn's dtype (float32 - float32, float64 - float32)
            if (dh <= (T)h0) {
                f = 1;
                drainz[c] = z0;
                drainh[c] = h0;
            }
        }
        out[c] = f;
    }
};

// ---- core.rank as a replay (when the BFS ran without it) -----------------------------------------------
struct RankOp {
    int32_t* rank;
    __device__ __forceinline__ void operator()(cell_t c, long long, int lev) const { rank[c] = lev; }
};

// ---- streams.strahler_order -----------------------------------------------------------------------------
struct StrahlerOp {
    const uint8_t* upmask;
    const uint8_t* mask;  // may be null
    uint8_t* out;
    long long ncol;
    __device__ __forceinline__ void operator()(cell_t c, long long, int) const {
        uint32_t m = __ldg(upmask + c);
        uint8_t so = 0, smax = 0;  // strord[c], strmax[c] built from the pushes of the upstream cells
        while (m) {
            const int k = 31 - __clz(m);
            m ^= 1u << k;
            const long long u = (long long)c + pfd_slot_off(k, ncol);
            if (mask && !__ldg(mask + u)) continue;  // streams.py:252-253: masked-out cells do not push
            const uint8_t sto = ld_cg(out + u);
            if (so < sto)
                so = sto;
            else if (sto == so && smax == sto)
                so = (uint8_t)(so + 1);
            if (smax < sto) smax = sto;
        }
        const bool in = mask ? (__ldg(mask + c) != 0) : true;
        if (in && so == 0) so = 1;  // headwater
        out[c] = so;
    }
};

// ---- dem.height_above_nearest_drain ----------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T elev_sub(T a, T b);
template <>
__device__ __forceinline__ float elev_sub<float>(float a, float b) { return __fsub_rn(a, b); }
template <>
__device__ __forceinline__ double elev_sub<double>(double a, double b) { return __dsub_rn(a, b); }

template <typename T>
struct HandOp {
    const uint8_t* dir;
    const uint8_t* drain;
    const T* elevtn;
    double* out;
    long long ncol;
    __device__ __forceinline__ void operator()(cell_t c, long long, int) const {
        if (__ldg(drain + c) == 1) {
            out[c] = 0.0;
            return;
        }
        const uint32_t d = __ldg(dir + c);
        const long long ds = ds_of(c, d, ncol);
        const T dz = elev_sub<T>(__ldg(elevtn + c), __ldg(elevtn + ds));
        const double h_ds = (d < 8u) ? ld_cg(out + ds) : 0.0;  // a pit reads its own initial 0 (dem.py:323)
        out[c] = __dadd_rn(h_ds, (double)dz);
    }
};

// ---- element-wise helpers --------------------------------------------------------------------------------
template <typename T>
__global__ void fill_kernel(T* __restrict__ out, int64_t n, T v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = v;
}

// upstream_area("cell") init: ones, -9999 on nodata (pyflwdir.py:790-800)
__global__ void uparea_init_kernel(const uint8_t* __restrict__ dir, int64_t n, int32_t* __restrict__ out) {
    const int64_t i4 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4;
    if (i4 >= n) return;
    if (i4 + 4 <= n) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(dir + i4);
        int4 r;
        r.x = ((w & 0xFFu) == 0xFFu) ? -9999 : 1;
        r.y = (((w >> 8) & 0xFFu) == 0xFFu) ? -9999 : 1;
        r.z = (((w >> 16) & 0xFFu) == 0xFFu) ? -9999 : 1;
        r.w = (((w >> 24) & 0xFFu) == 0xFFu) ? -9999 : 1;
        *reinterpret_cast<int4*>(out + i4) = r;
    } else {
        for (int64_t i = i4; i < n; ++i) out[i] = (dir[i] == PFD_DIR_NODATA) ? -9999 : 1;
    }
}

// basins[outlets[k]] = ids[k]; serial "last one wins" like numpy fancy assignment is kept by letting the
// highest k win through an atomicMax on the writer index first.
template <typename IDX, typename U>
__global__ void scatter_ids_kernel(const IDX* __restrict__ idxs, const U* __restrict__ ids, int64_t k0, int64_t k1,
                                   int64_t n, U* __restrict__ out, unsigned int* __restrict__ flag) {
    for (int64_t k = k0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < k1; k += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = (int64_t)idxs[k];
        if (i < 0) i += n;  // numpy negative indexing
        if (i < 0 || i >= n) {
            atomicOr(flag, 4u);
            continue;
        }
        out[i] = ids[k];
    }
}
