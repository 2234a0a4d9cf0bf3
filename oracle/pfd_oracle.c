/*
 * pfd_oracle.c -- CPU ORACLE for the D8 flow-network hot path of Deltares/pyflwdir v0.5.12.
 *
 * *** TEST INFRASTRUCTURE, NOT PRODUCT CODE. ***
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library, and only as the checker / the CPU arm that is timed BESIDE the GPU path. Nothing
 * under pyflwdir_b200/ imports, links or calls it.
 *
 * It is a plain-C, single-threaded restatement (the reference's numba kernels are single-threaded
 * scalar loops too) of:
 *   core_d8.drdc / from_array / to_array     /root/reference/pyflwdir/core_d8.py:22-39,42-67,86-102
 *   core.rank / upstream_count / upstream_matrix / idxs_seq / fillnodata_upstream / pit_indices
 *                                             /root/reference/pyflwdir/core.py:17-47,50-61,67-84,87-117,120-146,225-232
 *   streams.accuflux / accuflux_ds / strahler_order
 *                                             /root/reference/pyflwdir/streams.py:15-41,44-70,228-269
 *   dem.height_above_nearest_drain            /root/reference/pyflwdir/dem.py:299-330
 *   dem.fill_depressions (pfd_oracle_fill.inc)  /root/reference/pyflwdir/dem.py:17-143
 * Parity is PINNED: tests/golden/make_golden.py imports the real reference (numba) in the build
 * container and writes golden vectors; tests/test_oracle.py checks this library against them (and
 * against the live reference whenever /root/reference is present).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../pyflwdir_b200/csrc/pfd_synth.h"

/* ---- core_d8.drdc (core_d8.py:22-39), evaluated for all 256 codes exactly as numba does:
 *      np.int8(2 - np.log2(dd)) etc. truncate toward zero. Legal codes: core_d8.py:19 _all. ---- */
static int8_t orc_dr[256], orc_dc[256];
static int orc_drdc_ready = 0;

static void orc_init_drdc(void) {
    if (orc_drdc_ready) return;
    for (int dd = 0; dd < 256; ++dd) {
        int dr = 0, dc = 0;
        if (dd <= 8) {
            if (dd >= 2) {
                dr = 1;
                dc = (int)(int8_t)(2.0 - log2((double)dd));
            } else {
                dr = 0;
                dc = dd;
            }
        } else if (dd <= 128) {
            if (dd == 16) {
                dr = 0;
                dc = -1;
            } else {
                dr = -1;
                dc = (int)(int8_t)(log2((double)dd) - 6.0);
            }
        }
        orc_dr[dd] = (int8_t)dr;
        orc_dc[dd] = (int8_t)dc;
    }
    orc_drdc_ready = 1;
}

void orc_drdc_table(int8_t* dr, int8_t* dc) {
    orc_init_drdc();
    memcpy(dr, orc_dr, 256);
    memcpy(dc, orc_dc, 256);
}

/* core_d8.check_values (core_d8.py:115-122): 1 iff every value is one of the 11 legal codes */
int orc_d8_check_values(const uint8_t* flwdir, int64_t size) {
    static const uint8_t all[11] = {32, 64, 128, 16, 0, 1, 8, 4, 2, 247, 255};
    for (int64_t i = 0; i < size; ++i) {
        int found = 0;
        for (int k = 0; k < 11; ++k)
            if (all[k] == flwdir[i]) {
                found = 1;
                break;
            }
        if (!found) return 0;
    }
    return 1;
}

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

#define IDX int32_t
#define SFX i32
#include "pfd_oracle_body.inc"
#undef IDX
#undef SFX

#define IDX uint32_t
#define SFX u32
#include "pfd_oracle_body.inc"
#undef IDX
#undef SFX

#define IDX int64_t
#define SFX i64
#include "pfd_oracle_body.inc"
#undef IDX
#undef SFX

/* ---- dem.fill_depressions (dem.py:17-143): one instance per elevation type ---- */
#define FILL_T float
#define FILL_SFX f32
#include "pfd_oracle_fill.inc"
#undef FILL_T
#undef FILL_SFX
#define FILL_T double
#define FILL_SFX f64
#include "pfd_oracle_fill.inc"
#undef FILL_T
#undef FILL_SFX

/* ---- synthetic input (not part of the reference): SURVEY.md §8(d) generator, host version ---- */
/* The generator (NOT part of the measured path) is split over rows with pthreads so that the CPU-only reference arm
 * of bench.py can build its 8192^2 input in seconds; every cell is computed independently, so the result does not
 * depend on the thread count (ORC_SYNTH_THREADS, default = online cores, capped at 64). */
#include <pthread.h>
#include <unistd.h>

typedef struct {
    int64_t r0, r1, nrow, ncol, nref;
    int octaves;
    uint32_t seed;
    float sea_level;
    const float* zin;
    float* z;
    uint8_t* d8;
} orc_synth_job;

static void* orc_synth_elev_rows(void* arg) {
    orc_synth_job* j = (orc_synth_job*)arg;
    for (int64_t r = j->r0; r < j->r1; ++r)
        for (int64_t c = 0; c < j->ncol; ++c) j->z[r * j->ncol + c] = pfd_synth_z(r, c, j->nref, j->octaves, j->seed);
    return NULL;
}

static void* orc_synth_d8_rows(void* arg) {
    orc_synth_job* j = (orc_synth_job*)arg;
    const int64_t nrow = j->nrow, ncol = j->ncol;
    for (int64_t r = j->r0; r < j->r1; ++r)
        for (int64_t c = 0; c < ncol; ++c) {
            float w[9];
            int valid[9];
            for (int k = 0; k < 9; ++k) {
                int64_t rr = r + k / 3 - 1, cc = c + k % 3 - 1;
                valid[k] = (rr >= 0 && rr < nrow && cc >= 0 && cc < ncol);
                w[k] = valid[k] ? j->zin[rr * ncol + cc] : 0.0f;
            }
            j->d8[r * ncol + c] = pfd_synth_d8_from_window(w, valid, j->sea_level);
        }
    return NULL;
}

static void orc_synth_run(void* (*fn)(void*), orc_synth_job proto) {
    long nt = sysconf(_SC_NPROCESSORS_ONLN);
    const char* env = getenv("ORC_SYNTH_THREADS");
    if (env) nt = atol(env);
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    if (nt > proto.nrow) nt = (long)proto.nrow;
    pthread_t th[64];
    orc_synth_job jobs[64];
    int started[64];
    for (long t = 0; t < nt; ++t) {
        jobs[t] = proto;
        jobs[t].r0 = proto.nrow * t / nt;
        jobs[t].r1 = proto.nrow * (t + 1) / nt;
        started[t] = (t > 0) && pthread_create(&th[t], NULL, fn, &jobs[t]) == 0;
    }
    fn(&jobs[0]);
    for (long t = 1; t < nt; ++t) {
        if (started[t]) pthread_join(th[t], NULL);
        else fn(&jobs[t]); /* thread creation failed: do the rows here */
    }
}

void orc_synth_elevation(int64_t nrow, int64_t ncol, int64_t nref, int octaves, uint32_t seed, float* z) {
    orc_synth_job j = {0, 0, nrow, ncol, nref, octaves, seed, 0.0f, NULL, z, NULL};
    orc_synth_run(orc_synth_elev_rows, j);
}

void orc_synth_d8(const float* z, int64_t nrow, int64_t ncol, float sea_level, uint8_t* d8) {
    orc_synth_job j = {0, 0, nrow, ncol, 0, 0, 0, sea_level, z, NULL, d8};
    orc_synth_run(orc_synth_d8_rows, j);
}
