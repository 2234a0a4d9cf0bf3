// pfd_verify.cuh -- element-wise verification of finished outputs against their DEFINING RECURRENCES.
// Verification plumbing (tests, bench.py's parity_ok): every output of the hot path is a deterministic function of
// the values of the cell's graph neighbours, so one independent pass over the finished arrays that re-evaluates that
// function per cell and compares bit for bit proves the whole array at ANY raster size (the oracle can only be run
// at sizes it finishes in seconds). The kernels below share nothing with the solvers they check: they read the 1-byte
// direction graph and apply the reference's per-cell statement once.
//   rank    core.rank             pyflwdir/core.py:17-47       rank[i] = rank[ds] + 1; pit 0; no path to a pit -1; nodata -9999
//   basins  core.fillnodata_upstream / basins.basins core.py:120-146, basins.py:12-18   basins[i] = basins[ds]; pit k -> k + 1
//   uparea  streams.accuflux      pyflwdir/streams.py:15-41    accu[i] = 1 + sum(accu[upstream]) (int32 wrap); cells outside seq 1
//   idxs_ds core_d8.from_array    pyflwdir/core_d8.py:42-67
//   strord  streams.strahler_order pyflwdir/streams.py:228-269
//   hand    dem.height_above_nearest_drain pyflwdir/dem.py:299-330
//   accu    streams.accuflux, any dtype: the running sum in descending upstream index (what seq[::-1] produces)
#pragma once
#include "pfd_common.cuh"
#include "pfd_sweeps.cuh"

struct VerifyCounts {
    unsigned long long bad[8];  // 0 idxs_ds, 1 rank, 2 basins, 3 uparea, 4 pit ids, 5 generic
};

__device__ __forceinline__ void vf_flag(unsigned long long* slot, bool bad) {
    const unsigned m = __ballot_sync(__activemask(), bad);
    if (bad && (threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) atomicAdd(slot, (unsigned long long)__popc(m));
}

template <typename IDX>
__global__ void verify_flow_kernel(const uint8_t* __restrict__ dir, const uint8_t* __restrict__ upmask, int64_t n, int64_t ncol,
                                   const IDX* __restrict__ idxs, const int32_t* __restrict__ rank,
                                   const int32_t* __restrict__ uparea, const uint32_t* __restrict__ basins, VerifyCounts* out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t d = dir[i];
        const bool nodata = d == PFD_DIR_NODATA, pit = d == PFD_DIR_PIT || d == PFD_DIR_FPIT;
        const int64_t ds = (d < 8u) ? i + pfd_slot_off((int)d, ncol) : i;
        bool b_idx = false, b_rank = false, b_bas = false, b_upa = false;
        if (idxs) b_idx = idxs[i] != (nodata ? (IDX)-1 : (IDX)ds);
        if (rank) {
            const int32_t r = rank[i];
            if (nodata) b_rank = r != -9999;
            else if (pit) b_rank = r != 0;
            else {
                const int32_t rd = rank[ds];
                b_rank = (rd >= 0) ? (r != rd + 1) : (r != -1 || rd != -1);
            }
        }
        if (basins) {
            const uint32_t b = basins[i];
            if (nodata) b_bas = b != 0u;
            else if (pit) b_bas = b == 0u;  // the numbering itself is checked over the pit list
            else b_bas = b != basins[ds];
            if (rank && !nodata && (rank[i] < 0) != (b == 0u)) b_bas = true;  // basin 0 <=> the cell drains to no pit
        }
        if (uparea) {
            const int32_t a = uparea[i];
            if (nodata) b_upa = a != -9999;
            else if (rank && rank[i] < 0) b_upa = a != 1;  // not in seq: keeps its own unit weight
            else {
                uint32_t m = upmask[i];
                uint32_t acc = 1u;
                while (m) {
                    const int k = __ffs(m) - 1;
                    m &= m - 1;
                    acc += (uint32_t)uparea[i + pfd_slot_off(k, ncol)];
                }
                b_upa = a != (int32_t)acc;
            }
        }
        vf_flag(&out->bad[0], b_idx);
        vf_flag(&out->bad[1], b_rank);
        vf_flag(&out->bad[2], b_bas);
        vf_flag(&out->bad[3], b_upa);
    }
}

// pit k (ascending linear index, the reference's order) carries basin id id_off + k + 1
__global__ void verify_pit_ids_kernel(const cell_t* __restrict__ pits, int64_t npits, const uint32_t* __restrict__ basins,
                                      VerifyCounts* out) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < npits; k += (int64_t)gridDim.x * blockDim.x) {
        const bool bad = basins[pits[k]] != (uint32_t)(k + 1) || (k > 0 && pits[k - 1] >= pits[k]);
        vf_flag(&out->bad[4], bad);
    }
}

// streams.strahler_order (streams.py:250-269) re-evaluated per cell from the finished orders of its upstream cells
__global__ void verify_strahler_kernel(const uint8_t* __restrict__ dir, const uint8_t* __restrict__ upmask,
                                       const int32_t* __restrict__ rank, const uint8_t* __restrict__ mask, int64_t n,
                                       int64_t ncol, const uint8_t* __restrict__ so, VerifyCounts* out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint8_t want = 0;
        if (dir[i] != PFD_DIR_NODATA && rank[i] >= 0) {
            uint32_t m = upmask[i];
            uint8_t o = 0, mx = 0;
            while (m) {
                const int k = 31 - __clz(m);
                m ^= 1u << k;
                const int64_t u = i + pfd_slot_off(k, ncol);
                if (mask && !mask[u]) continue;
                const uint8_t s = so[u];
                if (o < s) o = s;
                else if (s == o && mx == s) o = (uint8_t)(o + 1);
                if (mx < s) mx = s;
            }
            if ((!mask || mask[i]) && o == 0) o = 1;
            want = o;
        }
        vf_flag(&out->bad[5], so[i] != want);
    }
}

// dem.height_above_nearest_drain (dem.py:316-329): hand[i] = hand[ds] + (elevtn[i] - elevtn[ds]) hop by hop
template <typename T>
__global__ void verify_hand_kernel(const uint8_t* __restrict__ dir, const int32_t* __restrict__ rank,
                                   const uint8_t* __restrict__ drain, const T* __restrict__ elevtn, int64_t n, int64_t ncol,
                                   const double* __restrict__ hand, VerifyCounts* out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t d = dir[i];
        double want = -9999.0;
        if (d != PFD_DIR_NODATA && rank[i] >= 0) {
            if (drain[i] == 1) want = 0.0;
            else {
                const int64_t ds = (d < 8u) ? i + pfd_slot_off((int)d, ncol) : i;
                const T dz = elev_sub<T>(elevtn[i], elevtn[ds]);
                want = __dadd_rn((d < 8u) ? hand[ds] : 0.0, (double)dz);
            }
        }
        const double got = hand[i];
        vf_flag(&out->bad[5], __double_as_longlong(got) != __double_as_longlong(want));
    }
}

// streams.accuflux (streams.py:36-40) over the walk order = running sum in DESCENDING upstream index
template <typename T>
__global__ void verify_accuflux_kernel(const uint8_t* __restrict__ dir, const uint8_t* __restrict__ upmask,
                                       const int32_t* __restrict__ rank, const T* __restrict__ data, NoData nd, int64_t n,
                                       int64_t ncol, const T* __restrict__ accu, VerifyCounts* out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        T want = data[i];
        if (dir[i] != PFD_DIR_NODATA && rank[i] >= 0) {
            uint32_t m = upmask[i];
            while (m) {
                const int k = 31 - __clz(m);
                m ^= 1u << k;
                const T a = accu[i + pfd_slot_off(k, ncol)];
                if (not_nodata(want, nd) && not_nodata(a, nd)) want = acc_add(want, a);
            }
        }
        const T got = accu[i];
        bool bad;
        if (sizeof(T) == 8) bad = *reinterpret_cast<const unsigned long long*>(&got) != *reinterpret_cast<const unsigned long long*>(&want);
        else if (sizeof(T) == 4) bad = *reinterpret_cast<const uint32_t*>(&got) != *reinterpret_cast<const uint32_t*>(&want);
        else bad = got != want;
        vf_flag(&out->bad[5], bad);
    }
}
