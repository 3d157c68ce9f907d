/* One LiDAR scan through the C ABI, in plain C: local-map cube -> raw points to HBM -> de-skew -> voxel grid ->
 * iterated-EKF update -> map_incremental.  The calls are the ones INTEGRATION.md places into laserMapping.cpp /
 * IMU_Processing.hpp.  Build:  gcc -std=c99 -Iinclude examples/pipeline.c fast_lio_b200/libfastlio_b200.so -lm
 * Run on a machine with a CUDA GPU (there is no CPU path: fl_map_create fails loudly without one). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "fastlio_b200.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc_ = (call);                                                             \
        if (rc_ < 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, fl_last_error()); return 1; } \
    } while (0)

static float frand(unsigned* s) { *s = *s * 1664525u + 1013904223u; return (float)(*s >> 8) / 16777216.0f; }

int main(void) {
    enum { N_MAP = 200000, N_RAW = 40000, N_POSE = 21 };
    unsigned seed = 7u;
    int i, j;

    /* a flat floor at z = -1.5 m with one point per 0.5 m cell plays the map */
    float* map = (float*)malloc(sizeof(float) * 4 * N_MAP);
    for (i = 0; i < N_MAP; i++) {
        int ix = i % 448 - 224, iy = i / 448 - 223;
        map[4 * i + 0] = 0.5f * ix + 0.5f * frand(&seed);
        map[4 * i + 1] = 0.5f * iy + 0.5f * frand(&seed);
        map[4 * i + 2] = -1.5f + 0.01f * (frand(&seed) - 0.5f);
        map[4 * i + 3] = 10.0f;
    }
    fl_map_t* ikdtree = NULL;
    CHECK(fl_map_create(&ikdtree, 0, 0.5f));                   /* KD_TREE<PointType> ikdtree; set_downsample_param(0.5) */
    CHECK(fl_map_build(ikdtree, map, N_MAP));                  /* ikdtree.Build(feats_down_world->points)                */

    fl_filter_t* kf = NULL;
    double limit[23];
    for (i = 0; i < 23; i++) limit[i] = 0.001;
    CHECK(fl_filter_create(&kf, ikdtree, N_RAW));
    CHECK(fl_filter_set_params(kf, 4, limit, 0));              /* kf.init_dyn_share(..., NUM_MAX_ITERATIONS, epsi)        */

    fl_scan_t* scan = NULL;
    fl_localmap_t* cube = NULL;
    CHECK(fl_scan_create(&scan, ikdtree));
    CHECK(fl_localmap_create(&cube, 1000.0, 100.0f));          /* cube_side_length, det_range                              */

    /* state_ikfom: identity attitude at the origin, gravity down */
    double x[26] = {0}, P[23 * 23] = {0};
    x[6] = 1.0; x[10] = 1.0; x[25] = -9.809;
    for (i = 0; i < 23; i++) P[i * 23 + i] = 1e-3;

    /* a raw scan of the floor seen from the (static) sensor, time-stamped over 100 ms */
    float* raw = (float*)malloc(sizeof(float) * 4 * N_RAW);
    float* t_ms = (float*)malloc(sizeof(float) * N_RAW);
    for (i = 0; i < N_RAW; i++) {
        raw[4 * i + 0] = 40.0f * (frand(&seed) - 0.5f);
        raw[4 * i + 1] = 40.0f * (frand(&seed) - 0.5f);
        raw[4 * i + 2] = -1.5f + 0.02f * (frand(&seed) - 0.5f) + 0.03f;     /* 3 cm off: the update has something to correct */
        raw[4 * i + 3] = 10.0f;
        t_ms[i] = 100.0f * frand(&seed);
    }
    /* IMUpose (msg/Pose6D.msg) from the caller's forward propagation: here a sensor at rest */
    double pose[N_POSE * 22];
    for (i = 0; i < N_POSE; i++) {
        for (j = 0; j < 22; j++) pose[i * 22 + j] = 0.0;
        pose[i * 22 + 0] = 0.005 * i;                          /* offset_time */
        pose[i * 22 + 13] = pose[i * 22 + 17] = pose[i * 22 + 21] = 1.0;   /* rot = I */
    }

    int n_deleted = 0, n_down, out3[3];
    double solve_time = 0.0;
    CHECK(fl_localmap_segment(cube, ikdtree, x, NULL, &n_deleted));          /* lasermap_fov_segment()                  */
    CHECK(fl_scan_upload(scan, raw, t_ms, N_RAW));
    CHECK(fl_scan_undistort(scan, pose, N_POSE, x));                         /* p_imu->Process(...): de-skew             */
    n_down = fl_scan_voxel_downsample(scan, 0.5f);                           /* downSizeFilterSurf.filter(...)            */
    if (n_down < 0) { fprintf(stderr, "voxel grid: %s\n", fl_last_error()); return 1; }
    CHECK(fl_filter_update_scan(kf, scan, x, P, 0.001, &solve_time));        /* kf.update_iterated_dyn_share_modified    */
    CHECK(fl_filter_map_incremental(kf, 0.5, 1, out3));                      /* map_incremental()                        */

    printf("scan: %d raw -> %d down-sampled; state z = %+.4f m (the floor was 3 cm off); update %.3f ms on the device;\n"
           "map_incremental: %d to add, %d without down-sampling, Add_Points returned %d; map now %d points\n",
           N_RAW, n_down, x[2], 1e3 * solve_time, out3[0], out3[1], out3[2], fl_map_validnum(ikdtree));

    fl_scan_destroy(scan);
    fl_localmap_destroy(cube);
    fl_map_destroy(ikdtree);                                   /* any order: the filter keeps its map alive */
    fl_filter_destroy(kf);
    free(map); free(raw); free(t_ms);
    return fabs(x[2]) < 1.0 ? 0 : 1;
}
