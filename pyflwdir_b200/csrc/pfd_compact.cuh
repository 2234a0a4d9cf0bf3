// pfd_compact.cuh -- ordered stream compaction over the cell sequence, and the entry points built on it / next to it:
//   basins.subbasins_streamorder   (pyflwdir/basins.py:67-103): outlets numbered in seq[::-1] order
//   arithmetics.upstream_sum       (pyflwdir/arithmetics.py:150-169): element-wise, order-exact
#pragma once
#include "pfd_sweeps.cuh"

// ---------------------------------------------------------------------------------------------------------
// Ordered compaction. Position q in [0, m) maps to cell seq[q] (REVERSED = 1: seq[m-1-q], 2: q itself); the cells whose predicate holds
// are written to out_cells in ascending q and, optionally, label[cell] = 1-based ordinal. Every thread owns
// CP_PER_THREAD consecutive positions, so order is preserved with one block-wide exclusive scan.
// Pass 1 counts per chunk, scan_counts_kernel (pfd_parse.cuh) turns counts into offsets, pass 2 writes.
// ---------------------------------------------------------------------------------------------------------
#define CP_THREADS 256
#define CP_PER_THREAD 8
#define CP_CHUNK (CP_THREADS * CP_PER_THREAD)

template <class Pred, int REVERSED>
__device__ __forceinline__ uint32_t cp_flags(const cell_t* __restrict__ seq, long long m, const Pred& pred, long long q0,
                                             cell_t* cells) {
    uint32_t flags = 0;
#pragma unroll
    for (int e = 0; e < CP_PER_THREAD; ++e) {
        const long long q = q0 + e;
        if (q < m) {
            const cell_t c = (REVERSED == 2) ? (cell_t)q : __ldg(seq + (REVERSED == 1 ? m - 1 - q : q));
            cells[e] = c;
            if (pred(c)) flags |= 1u << e;
        }
    }
    return flags;
}

template <class Pred, int REVERSED>
__global__ void __launch_bounds__(CP_THREADS) compact_count_kernel(const cell_t* __restrict__ seq, long long m, Pred pred,
                                                                   uint32_t* __restrict__ blk_counts) {
    __shared__ uint32_t s_w[CP_THREADS / 32];
    cell_t cells[CP_PER_THREAD];
    const long long q0 = (long long)blockIdx.x * CP_CHUNK + (long long)threadIdx.x * CP_PER_THREAD;
    uint32_t cnt = __popc(cp_flags<Pred, REVERSED>(seq, m, pred, q0, cells));
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < CP_THREADS / 32; ++w) t += s_w[w];
        blk_counts[blockIdx.x] = t;
    }
}

template <class Pred, int REVERSED, typename LABEL>
__global__ void __launch_bounds__(CP_THREADS) compact_scatter_kernel(const cell_t* __restrict__ seq, long long m, Pred pred,
                                                                     const unsigned long long* __restrict__ blk_off,
                                                                     cell_t* __restrict__ out_cells, LABEL* __restrict__ label) {
    __shared__ uint32_t s_w[CP_THREADS / 32];
    cell_t cells[CP_PER_THREAD];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long q0 = (long long)blockIdx.x * CP_CHUNK + (long long)threadIdx.x * CP_PER_THREAD;
    const uint32_t flags = cp_flags<Pred, REVERSED>(seq, m, pred, q0, cells);
    const uint32_t cnt = __popc(flags);
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    unsigned long long o = blk_off[blockIdx.x] + (incl - cnt);
    for (int w = 0; w < warp; ++w) o += s_w[w];
#pragma unroll
    for (int e = 0; e < CP_PER_THREAD; ++e) {
        if (flags & (1u << e)) {
            out_cells[o] = cells[e];
            if (label) label[cells[e]] = (LABEL)(o + 1ull);
            ++o;
        }
    }
}

// basins.subbasins_streamorder (basins.py:90-100): an outlet is a cell of stream order >= min_sto (and inside the
// mask) whose downstream cell has another stream order, or a pit
struct SubbasinOutletPred {
    const uint8_t* dir;
    const uint8_t* strord;
    const uint8_t* mask;  // may be null
    int min_sto;
    long long ncol;
    __device__ __forceinline__ bool operator()(cell_t c) const {
        const uint32_t so = __ldg(strord + c);
        if ((mask && __ldg(mask + c) == 0) || (int)so < min_sto) return false;
        const uint32_t d = __ldg(dir + c);
        if (d >= 8u) return true;  // idx_ds == idx0
        return so != (uint32_t)__ldg(strord + ((long long)c + pfd_slot_off((int)d, ncol)));
    }
};

__global__ void max_u8_kernel(const uint8_t* __restrict__ a, int64_t n, unsigned int* __restrict__ out) {
    unsigned int mx = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        mx = max(mx, (unsigned int)a[i]);
    mx = __reduce_max_sync(0xFFFFFFFFu, mx);
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(out, mx);
}

// ---------------------------------------------------------------------------------------------------------
// arithmetics.upstream_sum (arithmetics.py:150-169). The reference walks idx0 = 0..N-1 and either adds data[idx0] to
// arr_sum[idx_ds] or -- when data[idx0] or data[idx_ds] is nodata -- ASSIGNS nodata to arr_sum[idx0], wiping what the
// upstream cells with a smaller index added before. Per cell D that is: the valid upstream neighbours below D in
// index order (slots 0..3), then the assignment at idx0 = D, then the neighbours above D (slots 4..7).
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void upstream_sum_kernel(const uint8_t* __restrict__ dir, const uint8_t* __restrict__ upmask, const T* __restrict__ data,
                                    int64_t n, long long ncol, NoData nd, T ndv, T* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t d = dir[i];
        T acc = (T)0;
        if (d != PFD_DIR_NODATA) {
            const bool valid = not_nodata<T>(data[i], nd);
            const bool wipe = d < 8u && (!valid || !not_nodata<T>(data[i + pfd_slot_off((int)d, ncol)], nd));
            const uint32_t m = upmask[i];
            if (valid) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (m & (1u << k)) {
                        const T v = data[i + pfd_slot_off(k, ncol)];
                        if (not_nodata<T>(v, nd)) acc = acc_add<T>(acc, v);
                    }
            }
            if (wipe) acc = ndv;
            if (valid) {
#pragma unroll
                for (int k = 4; k < 8; ++k)
                    if (m & (1u << k)) {
                        const T v = data[i + pfd_slot_off(k, ncol)];
                        if (not_nodata<T>(v, nd)) acc = acc_add<T>(acc, v);
                    }
            }
        }
        out[i] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------
// basins.subbasins_area (basins.py:194-233). The reference walks the sequence from down- to upstream; the cells
// that share a downstream cell are consecutive and ascending in it (core.idxs_seq), and they are the only ones that
// touch each other's state (upa_out of the parent, of themselves and of the parent's main-stem child). So one thread
// per PARENT replays its children in ascending index inside a level-synchronous sweep: bit-identical, no races.
// flag[c] = 1 marks the subbasin outlets; they are numbered afterwards in sequence order (ordered compaction).
// ---------------------------------------------------------------------------------------------------------
template <typename T> struct SubDiff {  // (a - b) as the reference's numba typing evaluates it, converted for `> area_min`
    static __device__ __forceinline__ double gap(T a, T b) { return (double)((long long)a - (long long)b); }
};
template <> struct SubDiff<float> {
    static __device__ __forceinline__ double gap(float a, float b) { return (double)__fsub_rn(a, b); }
};
template <> struct SubDiff<double> {
    static __device__ __forceinline__ double gap(double a, double b) { return __dsub_rn(a, b); }
};
template <typename T>
__device__ __forceinline__ T sub_wrap(T a, T b) {
    typedef typename AccT<T>::U U;
    return (T)((U)a - (U)b);
}
template <>
__device__ __forceinline__ float sub_wrap<float>(float a, float b) { return __fsub_rn(a, b); }
template <>
__device__ __forceinline__ double sub_wrap<double>(double a, double b) { return __dsub_rn(a, b); }

template <typename T, typename IDX>
struct SubbasinsAreaOp {
    const uint8_t* dir;
    const uint8_t* upmask;
    const IDX* us_main;
    const T* uparea;
    T* upa_out;     // pre-initialised with uparea
    uint8_t* flag;  // pre-initialised with 0
    double area_min;
    long long ncol;
    __device__ __forceinline__ void operator()(cell_t c, long long, int) const {
        if (__ldg(dir + c) >= 8u) flag[c] = 1;  // pit: a subbasin of its own (basins.py:211-214)
        uint32_t m = __ldg(upmask + c);
        if (!m) return;
        T upa_p = ld_cg(upa_out + c);  // upa_out[idx_ds]
        const T uparea_p = __ldg(uparea + c);
        const IDX mainv = __ldg(us_main + c);
        const long long main = (mainv == (IDX)-1) ? -1ll : (long long)mainv;
        while (m) {  // children in ascending index = their order in the sequence
            const int k = __ffs(m) - 1;
            m &= m - 1;
            const long long ch = (long long)c + pfd_slot_off(k, ncol);
            const T upa = __ldg(uparea + ch);
            if (SubDiff<T>::gap(upa_p, upa) > area_min && (double)upa > area_min) {
                const bool conf = SubDiff<T>::gap(uparea_p, upa) > area_min;
                const bool trib = main != ch;
                if (!conf || trib) {
                    flag[ch] = 1;
                    upa_out[ch] = upa;
                }
                if (trib) {
                    upa_p = sub_wrap<T>(upa_p, upa);
                    if (main >= 0) upa_out[main] = upa_p;
                }
            } else {
                upa_out[ch] = upa_p;
            }
        }
    }
};

struct FlagPred {
    const uint8_t* flag;
    __device__ __forceinline__ bool operator()(cell_t c) const { return __ldg(flag + c) != 0; }
};

// ---------------------------------------------------------------------------------------------------------
// arithmetics.moving_average / moving_median (arithmetics.py:67-147) over core._window (core.py:368-398): for every
// cell the n cells upstream along the main-upstream links, the cell, and the n cells downstream (optionally only
// while the stream order does not grow). One thread per cell; the window is visited in the reference's order
// (farthest upstream cell first), the weighted mean accumulates in float64 exactly like arithmetics._average, and
// the median uses numba's own median-of-three quick-select (numba/np/arraymath.py _partition / _select /
// _select_two), so that even the sign of a zero that ties with another zero comes out identical.
// ---------------------------------------------------------------------------------------------------------
#define MW_NMAX 64

template <typename T>
__device__ __forceinline__ bool mw_is_nodata(T v, double nodata, bool nan_nodata) {
    return nan_nodata ? isnan(v) : ((double)v == nodata);
}

// `w0 * v0` with float64 weights (the only weights dtype arithmetics.py:101 types with): a float64 product, no fma
__device__ __forceinline__ double mw_prod(double w, float v) { return __dmul_rn(w, (double)v); }
__device__ __forceinline__ double mw_prod(double w, double v) { return __dmul_rn(w, v); }

template <typename T>
__device__ __forceinline__ int mw_partition(T* A, int low, int high) {
    const int mid = (low + high) >> 1;
    T t;
    if (A[mid] < A[low]) { t = A[low]; A[low] = A[mid]; A[mid] = t; }
    if (A[high] < A[mid]) { t = A[high]; A[high] = A[mid]; A[mid] = t; }
    if (A[mid] < A[low]) { t = A[low]; A[low] = A[mid]; A[mid] = t; }
    const T pivot = A[mid];
    t = A[high]; A[high] = A[mid]; A[mid] = t;
    int i = low, j = high - 1;
    while (true) {
        while (i < high && A[i] < pivot) ++i;
        while (j >= low && pivot < A[j]) --j;
        if (i >= j) break;
        t = A[i]; A[i] = A[j]; A[j] = t;
        ++i;
        --j;
    }
    t = A[i]; A[i] = A[high]; A[high] = t;
    return i;
}
template <typename T>
__device__ __forceinline__ T mw_select(T* A, int k, int low, int high) {
    int i = mw_partition(A, low, high);
    while (i != k) {
        if (i < k) low = i + 1;
        else high = i - 1;
        i = mw_partition(A, low, high);
    }
    return A[k];
}
template <typename T>
__device__ __forceinline__ void mw_select_two(T* A, int k, int low, int high) {
    while (true) {
        const int i = mw_partition(A, low, high);
        if (i < k) low = i + 1;
        else if (i > k + 1) high = i - 1;
        else if (i == k) { mw_select(A, k + 1, i + 1, high); break; }
        else { mw_select(A, k, low, i - 1); break; }
    }
}

// MEDIAN = false: weighted mean (weights may be null = ones), MEDIAN = true: nan-median
template <typename T, typename IDX, bool MEDIAN>
__global__ void moving_window_kernel(const uint8_t* __restrict__ dir, const IDX* __restrict__ us_main,
                                     const uint8_t* __restrict__ strord, const T* __restrict__ data,
                                     const double* __restrict__ weights, int64_t n, long long ncol, int nwin, double nodata,
                                     T* __restrict__ out, unsigned int* __restrict__ flag) {
    const bool nan_nodata = isnan(nodata);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const T d0 = data[i];
        if ((double)d0 == nodata) {  // `data[idx0] == nodata` (never true for a NaN nodata)
            out[i] = (T)nodata;
            continue;
        }
        // n upstream cells along the main-upstream links
        long long up[MW_NMAX];
        int nu = 0;
        long long cur = i;
        while (nu < nwin) {
            const IDX u = us_main[cur];
            if (u == (IDX)-1) break;
            if ((long long)u < 0 || (long long)u >= n) {
                atomicOr(flag, 8u);
                break;
            }
            cur = (long long)u;
            up[nu++] = cur;
        }
        const uint32_t so0 = strord ? (uint32_t)strord[i] : 0u;
        if (!MEDIAN) {
            double v = 0.0, w = 0.0;
            auto add = [&](long long c) {
                const T v0 = data[c];
                if (mw_is_nodata<T>(v0, nodata, nan_nodata)) return;
                if (weights) {
                    const double w0 = weights[c];
                    v = __dadd_rn(v, mw_prod(w0, v0));
                    w = __dadd_rn(w, w0);
                } else {
                    v = __dadd_rn(v, mw_prod(1.0, v0));
                    w = __dadd_rn(w, 1.0);
                }
            };
            for (int k = nu - 1; k >= 0; --k) add(up[k]);
            add(i);
            cur = i;
            for (int t = 0; t < nwin; ++t) {
                const uint32_t d = dir[cur];
                if (d >= 8u) break;  // pit or nodata
                const long long ds = cur + pfd_slot_off((int)d, ncol);
                if (strord && (uint32_t)strord[ds] > so0) break;
                cur = ds;
                add(cur);
            }
            out[i] = (T)((w != 0.0) ? __ddiv_rn(v, w) : nodata);
        } else {
            T a[2 * MW_NMAX + 1];
            int m = 0;
            auto push = [&](long long c) {
                const T v0 = data[c];
                if (isnan(v0) || (!nan_nodata && (double)v0 == nodata)) return;  // -> NaN -> dropped by nanmedian
                a[m++] = v0;
            };
            for (int k = nu - 1; k >= 0; --k) push(up[k]);
            push(i);
            cur = i;
            for (int t = 0; t < nwin; ++t) {
                const uint32_t d = dir[cur];
                if (d >= 8u) break;
                const long long ds = cur + pfd_slot_off((int)d, ncol);
                if (strord && (uint32_t)strord[ds] > so0) break;
                cur = ds;
                push(cur);
            }
            T r;
            if (m == 0) {
                r = (T)nan("");
            } else if ((m & 1) == 0) {
                const int half = m >> 1;
                mw_select_two(a, half - 1, 0, m - 1);
                r = (T)((double)acc_add<T>(a[half - 1], a[half]) / 2.0);
            } else {
                r = mw_select(a, m >> 1, 0, m - 1);
            }
            out[i] = r;
        }
    }
}
