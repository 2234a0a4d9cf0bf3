// pfd_local.cuh -- local traces and region post-processing around the ordered flow graph:
//   Flwdir.downstream                    (pyflwdir/flwdir.py:394-410)
//   core._trace / path / snap            (pyflwdir/core.py:309-364, 400-480)
//   core.inflow_idxs / outflow_idxs      (pyflwdir/core.py:483-514)
//   basins.interbasin_mask               (pyflwdir/basins.py:23-64)
//   regions.region_outlets               (pyflwdir/regions.py:132-163)
//   regions.region_slices / region_bounds (pyflwdir/regions.py:58-129, scipy.ndimage.find_objects)
#pragma once
#include "pfd_compact.cuh"

// ---- Flwdir.downstream: data_out[mask] = data[idxs_ds[mask]]; pits and nodata cells keep their own value --------
template <typename W>
__global__ void downstream_kernel(const uint8_t* __restrict__ dir, const W* __restrict__ data, int64_t n, long long ncol,
                                  W* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t d = dir[i];
        out[i] = data[d < 8u ? i + pfd_slot_off((int)d, ncol) : i];
    }
}

// ---- core._trace for a batch of start cells (core.path / core.snap): one thread per start ------------------------
// Downstream traces read the 1-byte dir graph, upstream traces the caller's idxs_us_main. hop = NULL (cells) or the
// host-built float64 table of gis_utils.distance per (row of the cell, row delta + 1, column delta != 0).
// paths = NULL: count / end / distance only; else the trace of start i is written at paths[offsets[i] ...].
// flag bits: 8 = index outside the raster, 16 = trace longer than the raster (a loop without a stop condition).
template <typename IDX, typename OUT, bool UP>
__global__ void trace_kernel(const uint8_t* __restrict__ dir, const IDX* __restrict__ us_main, const uint8_t* __restrict__ mask,
                             const int64_t* __restrict__ starts, int64_t n0, int64_t n, long long ncol, int has_max,
                             double max_length, const double* __restrict__ hop, const long long* __restrict__ offsets,
                             int64_t* __restrict__ counts, int64_t* __restrict__ ends, double* __restrict__ dists,
                             OUT* __restrict__ paths, unsigned int* __restrict__ flag) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n0; t += (int64_t)gridDim.x * blockDim.x) {
        long long cur = starts[t];
        if (cur < 0 || cur >= n) {
            atomicOr(flag, 8u);
            counts[t] = 0;
            ends[t] = -1;
            dists[t] = 0.0;
            continue;
        }
        OUT* dst = paths ? paths + offsets[t] : nullptr;
        long long cnt = 0;
        double dist = 0.0, d = 1.0;
        if (dst) dst[cnt] = (OUT)cur;
        ++cnt;
        while (!mask || mask[cur] == 0) {
            long long nx;
            if (!UP) {
                const uint32_t c = dir[cur];
                if (c >= 8u) break;  // pit (idx1 == idx0) or nodata (idx1 == mv)
                nx = cur + pfd_slot_off((int)c, ncol);
            } else {
                const IDX u = us_main[cur];
                if (u == (IDX)-1 || (long long)u == cur) break;
                nx = (long long)u;
                if (nx < 0 || nx >= n) {
                    atomicOr(flag, 8u);
                    break;
                }
            }
            if (hop) {
                const long long r0 = cur / ncol, r1 = nx / ncol;
                const long long drow = r1 - r0;
                if (drow < -1 || drow > 1) {
                    atomicOr(flag, 8u);
                    break;
                }
                d = hop[(r0 * 3 + (drow + 1)) * 2 + ((nx - r1 * ncol) != (cur - r0 * ncol) ? 1 : 0)];
            }
            if (has_max && __dadd_rn(dist, d) > max_length) break;
            dist = __dadd_rn(dist, d);
            cur = nx;
            if (has_max ? !(d > 0.0) : (cnt > n)) {  // more cells than the raster holds (or hops of no length): a loop
                atomicOr(flag, 16u);
                break;
            }
            if (dst) dst[cnt] = (OUT)cur;
            ++cnt;
        }
        counts[t] = cnt;
        ends[t] = cur;
        dists[t] = dist;
    }
}

// exclusive scan of the trace lengths (a handful of starts: one thread)
__global__ void trace_offsets_kernel(const int64_t* __restrict__ counts, int64_t n0, long long* __restrict__ offsets) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        long long o = 0;
        for (int64_t i = 0; i < n0; ++i) {
            offsets[i] = o;
            o += counts[i];
        }
        offsets[n0] = o;
    }
}

// ---- core.outflow_idxs: down-sweep. state byte: bit 0 = the reference's `mask`, bit 1 = "appended" -----------------
struct OutflowOp {
    const uint8_t* dir;
    const uint8_t* region;
    uint8_t* st;  // pre-initialised with 1
    long long ncol;
    __device__ __forceinline__ void operator()(cell_t c, long long, int) const {
        const uint32_t d = __ldg(dir + c);
        const bool pit = d >= 8u;
        const long long ds = pit ? (long long)c : (long long)c + pfd_slot_off((int)d, ncol);
        const uint32_t m_ds = pit ? 1u : ((uint32_t)ld_cg(st + ds) & 1u);
        const bool out = m_ds && __ldg(region + c) && (pit || !__ldg(region + ds));
        st[c] = out ? (uint8_t)2 : (uint8_t)m_ds;
    }
};

// ---- core.inflow_idxs: up-sweep, parent-centric. The reference overwrites mask[idx_ds] once per upstream cell while it
// walks seq[::-1]; the last writer is the upstream cell that comes FIRST in seq = the one with the smallest index
// (siblings are consecutive and ascending in core.idxs_seq). ------------------------------------------------------------
struct InflowOp {
    const uint8_t* upmask;
    const uint8_t* region;
    uint8_t* st;  // pre-initialised with 1
    long long ncol;
    __device__ __forceinline__ void operator()(cell_t c, long long, int) const {
        uint32_t m = __ldg(upmask + c);
        if (!m) return;  // headwater: mask stays True
        const bool rc = __ldg(region + c) != 0;
        bool first = true;
        while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            const long long u = (long long)c + pfd_slot_off(k, ncol);
            const uint32_t mu = (uint32_t)ld_cg(st + u) & 1u;
            const bool in = mu && rc && !__ldg(region + u);
            if (in) st[u] = (uint8_t)(mu | 2u);
            if (first) {
                st[c] = in ? (uint8_t)0 : (uint8_t)mu;
                first = false;
            }
        }
    }
};

struct StateBit2Pred {
    const uint8_t* st;
    __device__ __forceinline__ bool operator()(cell_t c) const { return (__ldg(st + c) & 2u) != 0; }
};

// ---- basins.interbasin_mask ------------------------------------------------------------------------------------------
// step 1 (stream given): a cell is True if it or any cell upstream of it is a stream cell (up-sweep, OR of the children)
struct AnyUpstreamOp {
    const uint8_t* upmask;
    uint8_t* st;  // pre-initialised with the stream mask
    long long ncol;
    __device__ __forceinline__ void operator()(cell_t c, long long, int) const {
        if (ld_cg(st + c)) return;
        uint32_t m = __ldg(upmask + c);
        while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            if (ld_cg(st + ((long long)c + pfd_slot_off(k, ncol)))) {
                st[c] = 1;
                return;
            }
        }
    }
};
// step 2: propagate upstream, cut where the flow enters the region (down-sweep)
struct InterbasinOp {
    const uint8_t* dir;
    const uint8_t* region;
    uint8_t* st;
    long long ncol;
    __device__ __forceinline__ void operator()(cell_t c, long long, int) const {
        const uint32_t d = __ldg(dir + c);
        if (d >= 8u) return;  // pit: mask[idx0] = mask[idx_ds] is a no-op
        const long long ds = (long long)c + pfd_slot_off((int)d, ncol);
        uint8_t v = ld_cg(st + ds);
        if (!__ldg(region + c) && __ldg(region + ds)) v = 0;
        st[c] = v;
    }
};
__global__ void and_mask_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int64_t n, uint8_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (a[i] && b[i]) ? 1 : 0;
}

// ---- regions.region_outlets: cell inside a region (label > 0) whose downstream cell carries another label, or a pit ---
template <typename T>
struct RegionOutletPred {
    const uint8_t* dir;
    const T* regions;
    long long ncol;
    __device__ __forceinline__ bool operator()(cell_t c) const {
        const T lb = __ldg(regions + c);
        if (!(lb > (T)0)) return false;
        const uint32_t d = __ldg(dir + c);
        if (d >= 8u) return true;
        return __ldg(regions + ((long long)c + pfd_slot_off((int)d, ncol))) != lb;
    }
};
template <typename T>
__global__ void gather_labels_kernel(const cell_t* __restrict__ cells, int64_t m, const T* __restrict__ regions, int64_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (int64_t)regions[cells[i]];
}

// ---- regions.region_slices (scipy.ndimage.find_objects): bounding rows / columns of every label > 0 ------------------
template <typename T>
__global__ void max_label_kernel(const T* __restrict__ regions, int64_t n, unsigned long long* __restrict__ out) {
    unsigned long long mx = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const T v = regions[i];
        if (v > (T)0) mx = max(mx, (unsigned long long)v);
    }
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(out, mx);
}
// box[l] = (min row, max row, min col, max col) of label l + 1. Only cells on the edge of a run of equal labels can
// hold an extreme, so interior cells issue no atomics.
__global__ void region_box_init_kernel(int4* __restrict__ box, int64_t nlab) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nlab; i += (int64_t)gridDim.x * blockDim.x)
        box[i] = make_int4(0x7FFFFFFF, -1, 0x7FFFFFFF, -1);
}
template <typename T>
__global__ void region_box_kernel(const T* __restrict__ regions, long long nrow, long long ncol, int4* __restrict__ box) {
    const int64_t n = nrow * ncol;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const T lb = regions[i];
        if (!(lb > (T)0)) continue;
        const long long r = i / ncol, c = i - r * ncol;
        int* b = reinterpret_cast<int*>(box + ((unsigned long long)lb - 1ull));
        if (r == 0 || regions[i - ncol] != lb) atomicMin(b + 0, (int)r);
        if (r == nrow - 1 || regions[i + ncol] != lb) atomicMax(b + 1, (int)r);
        if (c == 0 || regions[i - 1] != lb) atomicMin(b + 2, (int)c);
        if (c == ncol - 1 || regions[i + 1] != lb) atomicMax(b + 3, (int)c);
    }
}
struct BoxPresentPred {
    const int4* box;
    __device__ __forceinline__ bool operator()(cell_t l) const { return __ldg(reinterpret_cast<const int*>(box + l) + 1) >= 0; }
};
// slices[k] = (row start, row stop, col start, col stop) of the k-th present label; labels[k] = its value
__global__ void region_slices_kernel(const cell_t* __restrict__ present, int64_t m, const int4* __restrict__ box,
                                     int64_t* __restrict__ labels, int4* __restrict__ slices) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const cell_t l = present[i];
        const int4 b = box[l];
        labels[i] = (int64_t)l + 1;
        slices[i] = make_int4(b.x, b.y + 1, b.z, b.w + 1);
    }
}

// ---------------------------------------------------------------------------------------------------------
// streams.streams (pyflwdir/streams.py:131-188): the stream segments between confluences of the masked network.
// The reference walks seq[::-1] and starts a segment at every masked cell no earlier segment ran through; a segment
// runs downstream (also across unmasked cells) until a cell with more than one masked upstream neighbour or a pit.
// "Some segment runs through c" is a 1-bit up-sweep: through[c] = mask[c] or (nup[c] <= 1 and any upstream neighbour
// is run through); a masked cell starts a segment unless it is reached that way. The starts are compacted in
// seq[::-1] order, every start is traced twice (lengths -> exclusive scans -> cells), one thread per start.
// ---------------------------------------------------------------------------------------------------------
struct StreamStartOp {
    const uint8_t* upmask;
    const uint8_t* mask;  // may be null = every cell
    uint8_t* st;          // bit 0 = a segment runs through the cell, bit 1 = the cell starts a segment; pre-set to 0
    long long ncol;
    __device__ __forceinline__ void operator()(cell_t c, long long, int) const {
        uint32_t m = __ldg(upmask + c);
        int nup = 0;
        bool any = false;
        while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            const long long u = (long long)c + pfd_slot_off(k, ncol);
            nup += (!mask || __ldg(mask + u)) ? 1 : 0;
            any = any || (ld_cg(st + u) & 1u);
        }
        const bool in = !mask || __ldg(mask + c);
        const bool reached = nup <= 1 && any;
        st[c] = (uint8_t)(((in || reached) ? 1u : 0u) | ((in && !reached) ? 2u : 0u));
    }
};

// number of masked upstream neighbours > 1 (core.upstream_count with a mask, core.py:50-61)
__device__ __forceinline__ bool stream_confluence(const uint8_t* __restrict__ upmask, const uint8_t* __restrict__ mask, long long c,
                                                  long long ncol) {
    uint32_t m = upmask[c];
    if (!mask) return __popc(m) > 1;
    int nup = 0;
    while (m) {
        const int k = __ffs(m) - 1;
        m &= m - 1;
        nup += mask[c + pfd_slot_off(k, ncol)] ? 1 : 0;
    }
    return nup > 1;
}

// split of a segment of l cells (streams.py:168-180): k pieces of n cells (+ the shared end point)
__device__ __forceinline__ void stream_split(long long l, long long max_len, long long& k, long long& n) {
    k = 1;
    n = l;
    if (max_len > 0 && l > max_len && ((double)l / (double)max_len) > 1.5) {
        k = (long long)rint((double)l / (double)max_len);  // Python round(): half to even
        n = (long long)rint((double)l / (double)k);
    }
}
__device__ __forceinline__ long long stream_piece_len(long long i, long long k, long long n, long long l) {
    const long long from = min(i * n, l);
    const long long to = (i + 1 == k) ? l : min(n * (i + 1) + 1, l);
    return max(to - from, 0ll);
}

// pass 1: length of every segment (bit 31: it ends in a pit), number of output pieces and indices
__global__ void stream_count_kernel(const uint8_t* __restrict__ dir, const uint8_t* __restrict__ upmask, const uint8_t* __restrict__ mask,
                                    const cell_t* __restrict__ starts, long long nstart, long long ncol, long long max_len,
                                    uint32_t* __restrict__ len, uint32_t* __restrict__ npiece, uint32_t* __restrict__ ncell) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nstart; t += (long long)gridDim.x * blockDim.x) {
        long long cur = starts[t], l = 1;
        bool pit;
        while (true) {
            const uint32_t d = dir[cur];
            pit = d >= 8u;
            if (pit) break;
            cur += pfd_slot_off((int)d, ncol);
            ++l;
            if (stream_confluence(upmask, mask, cur, ncol)) break;
        }
        long long k, n;
        stream_split(l, max_len, k, n);
        long long cells = 0;
        if (k == 1) cells = l;
        else
            for (long long i = 0; i < k; ++i) cells += stream_piece_len(i, k, n, l);
        len[t] = (uint32_t)l | (pit ? 0x80000000u : 0u);
        npiece[t] = (uint32_t)(k + (pit ? 1 : 0));
        ncell[t] = (uint32_t)(cells + (pit ? 2 : 0));
    }
}

// pass 2: write the indices of every piece and its start offset
__global__ void stream_write_kernel(const uint8_t* __restrict__ dir, const cell_t* __restrict__ starts, long long nstart, long long ncol,
                                    long long max_len, const uint32_t* __restrict__ len, const unsigned long long* __restrict__ piece_off,
                                    const unsigned long long* __restrict__ cell_off, long long* __restrict__ offsets,
                                    cell_t* __restrict__ cells) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nstart; t += (long long)gridDim.x * blockDim.x) {
        const uint32_t lw = len[t];
        const long long l = (long long)(lw & 0x7FFFFFFFu);
        const bool pit = (lw >> 31) != 0;
        long long k, n;
        stream_split(l, max_len, k, n);
        unsigned long long po = piece_off[t], base = cell_off[t];
        long long cur = starts[t], i = 0;
        offsets[po] = (long long)base;
        for (long long p = 0; p < l; ++p) {
            if (p > 0) cur += pfd_slot_off((int)dir[cur], ncol);
            cells[base + (unsigned long long)(p - i * n)] = (cell_t)cur;
            if (i + 1 < k && p == n * (i + 1)) {  // shared end point: last index of piece i, first of piece i + 1
                base += (unsigned long long)(n + 1);
                ++i;
                offsets[po + (unsigned long long)i] = (long long)base;
                cells[base] = (cell_t)cur;
            }
        }
        unsigned long long end = base + (unsigned long long)(l - min(i * n, l));
        for (long long j = i + 1; j < k; ++j) offsets[po + (unsigned long long)j] = (long long)end;  // pieces beyond the end are empty
        if (pit) {
            offsets[po + (unsigned long long)k] = (long long)end;
            cells[end] = (cell_t)cur;
            cells[end + 1] = (cell_t)cur;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// basins.subbasins_pfafstetter (pyflwdir/basins.py:106-191). The reference pops labels from a FIFO; a label owns a
// "stem" (the main-upstream chain that carries it), picks its four largest unlabelled tributaries, labels them (odd
// digits) and carves the stem into interbasins (even digits); the new labels go to the back of the FIFO, so the FIFO is
// processed depth by depth, and the stems and tributary subtrees of the labels of one depth are disjoint. One depth =
// one round here, one thread per label: count the candidate tributaries along the stem -> scan -> collect them, restore
// sequence order, numba's argsort on -uparea (ties!), label, emit children / outlets into per-label slots -> ordered
// compaction into the next round's label list and the outlet list.
// ---------------------------------------------------------------------------------------------------------
#define PF_MV 0xFFFFFFFFu

__global__ void pf_prepare_kernel(uint8_t* __restrict__ strord, int64_t n, int depth) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if ((int)strord[i] > depth + 1) strord[i] = 0;
}
__global__ void pf_pos_kernel(const cell_t* __restrict__ seq, int64_t m, uint32_t* __restrict__ pos) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < m; q += (int64_t)gridDim.x * blockDim.x) pos[seq[q]] = (uint32_t)q;
}
template <typename IDX>
__global__ void pf_main32_kernel(const IDX* __restrict__ us_main, int64_t n, uint32_t* __restrict__ out, unsigned int* __restrict__ flag) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const IDX u = us_main[i];
        uint32_t v = PF_MV;
        if (u != (IDX)-1) {
            if ((long long)u < 0 || (long long)u >= n) atomicOr(flag, 8u);
            else v = (uint32_t)u;
        }
        out[i] = v;
    }
}
template <typename T>
__global__ void pf_to_double_kernel(const T* __restrict__ a, int64_t n, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (double)a[i];
}

struct PfGraph {
    const uint8_t* dir;
    const uint8_t* upmask;
    const uint8_t* strord;
    const uint32_t* main;   // idxs_us_main as uint32, PF_MV at headwaters
    const uint32_t* pos;    // position of a cell inside seq
    const double* uparea;
    int32_t* pb;            // pfaf_branch
    uint8_t* inidx;         // 1 where the cell is already in the outlet list (`idx1 not in idxs`)
    long long ncol;
    int depth;
};

// label its outlet, then the main stem upstream of it while the stream order is non-zero (basins.py:134-141,163-169)
__device__ __forceinline__ void pf_label_stem(const PfGraph& G, uint32_t cell, long long label) {
    G.pb[cell] = (int32_t)label;
    while (true) {
        const uint32_t u = G.main[cell];
        if (u == PF_MV || G.strord[u] == 0) break;
        cell = u;
        G.pb[cell] = (int32_t)label;
    }
}

__device__ __forceinline__ long long pf_pow10(int e) {
    long long p = 1;
    for (int i = 0; i < e; ++i) p *= 10;
    return p;
}

__global__ void pf_init_pits_kernel(PfGraph G, const cell_t* __restrict__ pits, long long npits, long long* __restrict__ lab,
                                    uint32_t* __restrict__ lab_out, cell_t* __restrict__ outlets) {
    long long base = 1;
    for (int d0 = 1; d0 < G.depth; ++d0) base += pf_pow10(d0);
    const long long p = pf_pow10(G.depth);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npits; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t c = pits[i];
        const long long label = base + (i + 1) * p;
        lab[i] = label;
        lab_out[i] = c;
        outlets[i] = c;
        G.inidx[c] = 1;
        pf_label_stem(G, c, label);
    }
}

// visits the unlabelled tributaries attached to the stem of (label, outlet): cells t that drain into a stem cell c with
// strord[t] > 0 and strord[t] > strord[c] (basins.py:107-114) and pfaf_branch[t] == 0
template <class F>
__device__ __forceinline__ void pf_for_candidates(const PfGraph& G, long long label, uint32_t outlet, F f) {
    uint32_t c = outlet;
    while ((long long)G.pb[c] == label) {
        uint32_t m = G.upmask[c];
        const uint32_t sc = G.strord[c];
        while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t t = (uint32_t)((long long)c + pfd_slot_off(k, G.ncol));
            const uint32_t st = G.strord[t];
            if (st > 0 && st > sc && G.pb[t] == 0) f(t);
        }
        const uint32_t u = G.main[c];
        if (u == PF_MV) break;
        c = u;
    }
}

__global__ void pf_count_kernel(PfGraph G, const long long* __restrict__ lab, const uint32_t* __restrict__ lab_out, long long nlab,
                                uint32_t* __restrict__ cnt) {
    for (long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x; l < nlab; l += (long long)gridDim.x * blockDim.x) {
        uint32_t n = 0;
        pf_for_candidates(G, lab[l], lab_out[l], [&](uint32_t) { ++n; });
        cnt[l] = n;
    }
}

// numba's `a < b` for float keys (numba/np/numpy_support.py lt_floats): NaNs sort last
__device__ __forceinline__ bool pf_lt(double a, double b) { return a < b || (isnan(b) && !isnan(a)); }

// np.argsort as numba compiles it (numba/misc/quicksort.py), on keys A[0..n) with the index array R
__device__ void pf_numba_argsort(const double* A, long long n, uint32_t* R) {
    for (long long i = 0; i < n; ++i) R[i] = (uint32_t)i;
    if (n < 2) return;
    long long slo[100], shi[100];
    int ns = 1;
    slo[0] = 0;
    shi[0] = n - 1;
    while (ns > 0) {
        --ns;
        long long low = slo[ns], high = shi[ns];
        while (high - low >= 15) {
            const long long mid = (low + high) >> 1;
            uint32_t t;
            if (pf_lt(A[R[mid]], A[R[low]])) { t = R[low]; R[low] = R[mid]; R[mid] = t; }
            if (pf_lt(A[R[high]], A[R[mid]])) { t = R[high]; R[high] = R[mid]; R[mid] = t; }
            if (pf_lt(A[R[mid]], A[R[low]])) { t = R[low]; R[low] = R[mid]; R[mid] = t; }
            const double pivot = A[R[mid]];
            t = R[high]; R[high] = R[mid]; R[mid] = t;
            long long i = low, j = high - 1;
            while (true) {
                while (i < high && pf_lt(A[R[i]], pivot)) ++i;
                while (j >= low && pf_lt(pivot, A[R[j]])) --j;
                if (i >= j) break;
                t = R[i]; R[i] = R[j]; R[j] = t;
                ++i;
                --j;
            }
            t = R[i]; R[i] = R[high]; R[high] = t;
            if (high - i > i - low) {
                if (high > i && ns < 100) { slo[ns] = i + 1; shi[ns] = high; ++ns; }
                high = i - 1;
            } else {
                if (i > low && ns < 100) { slo[ns] = low; shi[ns] = i - 1; ++ns; }
                low = i + 1;
            }
        }
        for (long long i = low + 1; i <= high; ++i) {
            const uint32_t k = R[i];
            const double v = A[k];
            long long j = i;
            while (j > low && pf_lt(v, A[R[j - 1]])) {
                R[j] = R[j - 1];
                --j;
            }
            R[j] = k;
        }
    }
}

// ascending heap sort of the candidate cells by their sequence position (unique keys)
__device__ void pf_sort_by_pos(const uint32_t* __restrict__ pos, uint32_t* c, long long n) {
    auto sift = [&](long long root, long long end) {
        while (true) {
            long long child = 2 * root + 1;
            if (child > end) break;
            if (child + 1 <= end && pos[c[child]] < pos[c[child + 1]]) ++child;
            if (pos[c[root]] < pos[c[child]]) {
                const uint32_t t = c[root]; c[root] = c[child]; c[child] = t;
                root = child;
            } else break;
        }
    };
    for (long long s = (n - 2) / 2; s >= 0; --s) sift(s, n - 1);
    for (long long e = n - 1; e > 0; --e) {
        const uint32_t t = c[0]; c[0] = c[e]; c[e] = t;
        sift(0, e - 1);
    }
}

// one label: basins.py:142-187. Children / outlets go to the label's 8 slots.
__global__ void pf_process_kernel(PfGraph G, const long long* __restrict__ lab, const uint32_t* __restrict__ lab_out, long long nlab,
                                  int d0, const unsigned long long* __restrict__ off, uint32_t* __restrict__ cand,
                                  double* __restrict__ key, uint32_t* __restrict__ R, long long* __restrict__ child_lab,
                                  uint32_t* __restrict__ child_out, uint32_t* __restrict__ nchild, uint32_t* __restrict__ out_cells,
                                  uint32_t* __restrict__ nout, unsigned int* __restrict__ flag) {
    for (long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x; l < nlab; l += (long long)gridDim.x * blockDim.x) {
        const long long pfaf0 = lab[l];
        const long long n = (long long)(off[l + 1] - off[l]);
        uint32_t nc = 0, no = 0;
        if (n > 0) {
            uint32_t* c = cand + off[l];
            double* kk = key + off[l];
            uint32_t* r = R + off[l];
            long long w = 0;
            pf_for_candidates(G, pfaf0, lab_out[l], [&](uint32_t t) { c[w++] = t; });
            pf_sort_by_pos(G.pos, c, n);  // the order of idxs_trib (sequence order)
            for (long long i = 0; i < n; ++i) kk[i] = -G.uparea[c[i]];
            pf_numba_argsort(kk, n, r);   // idxs0[np.argsort(-uparea[idxs0])]
            const int n4 = (int)min(n, 4ll);
            uint32_t t4[4], o4[4];
            double k4[4];
            for (int i = 0; i < n4; ++i) {
                t4[i] = c[r[i]];
                const uint32_t d = G.dir[t4[i]];
                k4[i] = -G.uparea[(long long)t4[i] + pfd_slot_off((int)d, G.ncol)];
            }
            pf_numba_argsort(k4, n4, o4);  // down- to upstream along the stem
            const long long p = pf_pow10(G.depth - d0);
            long long pfaf_int_ds = pfaf0;
            for (int i = 0; i < n4; ++i) {
                const uint32_t idx = t4[o4[i]];
                out_cells[l * 8 + no++] = idx;
                G.inidx[idx] = 1;
                const uint32_t ds = (uint32_t)((long long)idx + pfd_slot_off((int)G.dir[idx], G.ncol));
                uint32_t idx1 = G.main[ds];
                const long long pfaf_sub = pfaf0 + (long long)(i * 2 + 1) * p;
                pf_label_stem(G, idx, pfaf_sub);
                if (d0 < G.depth) {
                    child_lab[l * 8 + nc] = pfaf_sub;
                    child_out[l * 8 + nc] = idx;
                    ++nc;
                }
                if (idx1 == PF_MV) {  // the reference would index with -1 here
                    atomicOr(flag, 32u);
                    break;
                }
                if (!G.inidx[idx1]) {
                    out_cells[l * 8 + no++] = idx1;
                    G.inidx[idx1] = 1;
                    const long long pfaf_int = pfaf0 + (long long)(i + 1) * 2 * p;
                    const uint32_t int_outlet = idx1;
                    G.pb[idx1] = (int32_t)pfaf_int;
                    while (true) {
                        const uint32_t u = G.main[idx1];
                        if (u == PF_MV || (long long)G.pb[u] != pfaf_int_ds) break;
                        idx1 = u;
                        G.pb[idx1] = (int32_t)pfaf_int;
                    }
                    pfaf_int_ds = pfaf_int;
                    if (d0 < G.depth) {
                        child_lab[l * 8 + nc] = pfaf_int;
                        child_out[l * 8 + nc] = int_outlet;
                        ++nc;
                    }
                }
            }
        }
        nchild[l] = nc;
        nout[l] = no;
    }
}

__global__ void pf_gather_kernel(long long nlab, const long long* __restrict__ child_lab, const uint32_t* __restrict__ child_out,
                                 const uint32_t* __restrict__ nchild, const unsigned long long* __restrict__ child_off,
                                 const uint32_t* __restrict__ out_cells, const uint32_t* __restrict__ nout,
                                 const unsigned long long* __restrict__ out_off, long long* __restrict__ next_lab,
                                 uint32_t* __restrict__ next_out, cell_t* __restrict__ outlets_tail) {
    for (long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x; l < nlab; l += (long long)gridDim.x * blockDim.x) {
        for (uint32_t j = 0; j < nchild[l]; ++j) {
            next_lab[child_off[l] + j] = child_lab[l * 8 + j];
            next_out[child_off[l] + j] = child_out[l * 8 + j];
        }
        for (uint32_t j = 0; j < nout[l]; ++j) outlets_tail[out_off[l] + j] = out_cells[l * 8 + j];
    }
}

// pfaf_branch % 10**depth with Python's sign rule, as int64 (numba: int32 % int64 -> int64)
__global__ void pf_mod_kernel(const int32_t* __restrict__ pb, int64_t n, long long mod, int64_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        long long m = (long long)pb[i] % mod;
        if (m < 0) m += mod;
        out[i] = m;
    }
}

// ---- rivers.classify_estuary (pyflwdir/rivers.py:11-53): down-sweep. A cell whose downstream cell is estuary joins it
// if the river keeps widening downstream fast enough, else it marks the downstream cell as the upstream end (2). The
// parent's value is only tested for != 0, so the siblings' concurrent `= 2` writes do not change any decision. ----------
template <typename TD, typename TW>
struct EstuaryOp {
    const uint8_t* dir;
    const TD* rivdst;
    const TW* rivwth;
    int8_t* est;  // pre-initialised: 1 at the pits with elevtn <= max_elevtn, else 0
    double min_convergence;
    long long ncol;
    __device__ __forceinline__ void operator()(cell_t c, long long, int) const {
        const uint32_t d = __ldg(dir + c);
        if (d >= 8u) return;  // idx == idx_ds
        const long long ds = (long long)c + pfd_slot_off((int)d, ncol);
        if (ld_cg(est + ds) == 0) return;
        const TD dst_ds = __ldg(rivdst + ds);
        const TD dx = elev_sub<TD>(__ldg(rivdst + c), dst_ds);
        const TW dw = elev_sub<TW>(__ldg(rivwth + ds), __ldg(rivwth + c));
        bool conv;
        if (sizeof(TD) == 4 && sizeof(TW) == 4) conv = (double)__fdiv_rn((float)dw, (float)dx) > min_convergence;  // float32 / float32
        else conv = __ddiv_rn((double)dw, (double)dx) > min_convergence;
        if ((dst_ds == (TD)0 && dw <= (TW)0) || (dx > (TD)0 && conv)) est[c] = 1;
        else est[ds] = 2;  // most upstream estuary link
    }
};
