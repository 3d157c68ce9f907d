// extern "C" boundary (include/fastlio_b200.h).  Thin: argument checks, handle plumbing,
// status codes.  No torch types, no exceptions.
#include <atomic>
#include <chrono>
#include <cstring>
#include <mutex>
#include <new>

#include "../../include/fastlio_b200.h"
#include "filter.h"
#include "scan.h"

namespace fl {
const char* last_error();
int nccl_unique_id(void* out128);
}  // namespace fl

// A map handle is shared by the filters and scan front ends created on it: they keep it alive, so destroying the
// handles in any order is safe (fl_map_destroy only drops the caller's reference).
struct fl_map {
    fl::Map* impl;
    std::mutex mu;
    std::atomic<int> refs{1};
};
static void map_retain(fl_map* m) { m->refs.fetch_add(1, std::memory_order_relaxed); }
static void map_release(fl_map* m) {
    if (m->refs.fetch_sub(1, std::memory_order_acq_rel) == 1) {
        delete m->impl;
        delete m;
    }
}
struct fl_filter {
    fl::Filter* impl;
    fl_map* map;
    fl::DeviceBuffer flush;
};

struct fl_scan {
    fl::ScanFrontEnd* impl;
    fl_map* map;
};
struct fl_localmap {
    fl::LocalMapCube cube;
};

static_assert(sizeof(fl_pass_log_t) == sizeof(fl::PassLog), "pass-log layouts must match");

extern "C" {

const char* fl_last_error(void) { return fl::last_error(); }
int fl_version(void) { return 100; }
int fl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int fl_host_register(const void* ptr, unsigned long long bytes) {
    if (!ptr || !bytes) return FL_ERR_ARG;
    FL_CUDA(cudaHostRegister(const_cast<void*>(ptr), (size_t)bytes, cudaHostRegisterDefault));
    return FL_OK;
}
int fl_host_unregister(const void* ptr) {
    if (!ptr) return FL_ERR_ARG;
    FL_CUDA(cudaHostUnregister(const_cast<void*>(ptr)));
    return FL_OK;
}

// ------------------------------------------------------------------------------------ map
int fl_map_create(fl_map_t** out, int device, float downsample_size) {
    if (!out) { fl::set_last_error("fl_map_create: null out"); return FL_ERR_ARG; }
    *out = nullptr;
    int n = fl_device_count();
    if (n <= 0) { fl::set_last_error("fl_map_create: no CUDA device visible (this library has no CPU path)"); return FL_ERR_CUDA; }
    if (device < 0 || device >= n) { fl::set_last_error("fl_map_create: device %d out of range [0, %d)", device, n); return FL_ERR_ARG; }
    fl_map* m = new (std::nothrow) fl_map();
    if (!m) return FL_ERR_CAPACITY;
    m->impl = new (std::nothrow) fl::Map(device, downsample_size);
    if (!m->impl) { delete m; return FL_ERR_CAPACITY; }
    int rc = m->impl->init();
    if (rc != FL_OK) { delete m->impl; delete m; return rc; }
    *out = m;
    return FL_OK;
}
int fl_map_destroy(fl_map_t* m) {
    if (!m) return FL_OK;
    map_release(m);
    return FL_OK;
}
#define MAP_GUARD(m)                                                             \
    if (!(m) || !(m)->impl) { fl::set_last_error("null map handle"); return FL_ERR_ARG; } \
    std::lock_guard<std::mutex> _lk((m)->mu)

int fl_map_set_downsample(fl_map_t* m, float v) { MAP_GUARD(m); m->impl->set_downsample(v); return FL_OK; }
int fl_map_build(fl_map_t* m, const float* pts, int n) { MAP_GUARD(m); return m->impl->build(pts, n); }
int fl_map_size(fl_map_t* m) { MAP_GUARD(m); return m->impl->size(); }
int fl_map_validnum(fl_map_t* m) { MAP_GUARD(m); return m->impl->validnum(); }
int fl_map_knn(fl_map_t* m, const float* q, int nq, int k, float* out_pts, float* out_d2, int* out_cnt) {
    MAP_GUARD(m);
    if (nq > 0 && (!q || !out_pts || !out_d2 || !out_cnt)) { fl::set_last_error("fl_map_knn: null buffer"); return FL_ERR_ARG; }
    return m->impl->knn(q, nq, k, out_pts, out_d2, out_cnt);
}
int fl_map_add_points(fl_map_t* m, const float* pts, int n, int downsample_on) {
    MAP_GUARD(m);
    int added = 0;
    int rc = m->impl->add_points(pts, n, downsample_on != 0, &added);
    return rc == FL_OK ? added : rc;
}
int fl_map_delete_boxes(fl_map_t* m, const float* boxes6, int nb) {
    MAP_GUARD(m);
    int deleted = 0;
    int rc = m->impl->delete_boxes(boxes6, nb, &deleted);
    return rc == FL_OK ? deleted : rc;
}
int fl_map_add_boxes(fl_map_t* m, const float* boxes6, int nb) {
    MAP_GUARD(m);
    int revived = 0;
    int rc = m->impl->add_boxes(boxes6, nb, &revived);
    return rc == FL_OK ? revived : rc;
}
int fl_map_acquire_removed(fl_map_t* m, float* out, int cap) {
    MAP_GUARD(m);
    int n = 0;
    int rc = m->impl->acquire_removed(out, cap, &n);
    return rc == FL_OK ? n : rc;
}
int fl_map_flatten(fl_map_t* m, float* out, int cap) {
    MAP_GUARD(m);
    int n = 0;
    int rc = m->impl->flatten(out, cap, &n);
    return rc == FL_OK ? n : rc;
}
int fl_map_tree_range(fl_map_t* m, float* box6) { MAP_GUARD(m); if (!box6) return FL_ERR_ARG; return m->impl->tree_range(box6); }
int fl_map_rebuild(fl_map_t* m) { MAP_GUARD(m); return m->impl->rebuild(); }
int fl_map_stats(fl_map_t* m, int* out4) {
    MAP_GUARD(m);
    if (!out4) return FL_ERR_ARG;
    out4[0] = m->impl->view().n_main; out4[1] = m->impl->overflow_leaves();
    out4[2] = m->impl->view().n_levels; out4[3] = m->impl->rebuild_count();
    return FL_OK;
}

int fl_map_set_cell_directory(fl_map_t* m, int on, float cell_size) {
    MAP_GUARD(m);
    m->impl->set_cell_directory(on != 0, cell_size > 0.f ? cell_size : 0.f);
    return m->impl->build_directory();
}
int fl_map_dir_stats(fl_map_t* m, int* out6) {
    MAP_GUARD(m);
    if (!out6) return FL_ERR_ARG;
    return m->impl->dir_stats(out6);
}

// ------------------------------------------------------------------------------------ filter
#define FILTER_GUARD(f)                                                                        \
    if (!(f) || !(f)->impl) { fl::set_last_error("null filter handle"); return FL_ERR_ARG; }   \
    std::lock_guard<std::mutex> _lk((f)->map->mu)

int fl_filter_create(fl_filter_t** out, fl_map_t* map, int max_points) {
    if (!out || !map || !map->impl) { fl::set_last_error("fl_filter_create: null argument"); return FL_ERR_ARG; }
    *out = nullptr;
    fl_filter* f = new (std::nothrow) fl_filter();
    if (!f) return FL_ERR_CAPACITY;
    f->map = map;
    f->impl = new (std::nothrow) fl::Filter(map->impl, max_points);
    if (!f->impl) { delete f; return FL_ERR_CAPACITY; }
    int rc = f->impl->init();
    if (rc != FL_OK) { delete f->impl; delete f; return rc; }
    map_retain(map);
    *out = f;
    return FL_OK;
}
int fl_filter_destroy(fl_filter_t* f) {
    if (!f) return FL_OK;
    f->flush.release();
    delete f->impl;
    map_release(f->map);
    delete f;
    return FL_OK;
}
int fl_filter_set_params(fl_filter_t* f, int max_iter, const double* limit23, int extr) { FILTER_GUARD(f); return f->impl->set_params(max_iter, limit23, extr); }
int fl_filter_set_solver(fl_filter_t* f, int mode) { FILTER_GUARD(f); if (mode < 0 || mode > 1) return FL_ERR_ARG; f->impl->set_solver(mode); return FL_OK; }
int fl_filter_set_fused(fl_filter_t* f, int on) { FILTER_GUARD(f); f->impl->set_fused(on != 0); return FL_OK; }
int fl_filter_set_search(fl_filter_t* f, int mode) { FILTER_GUARD(f); if (mode < 0 || mode > 1) return FL_ERR_ARG; f->impl->set_search_mode(mode); return FL_OK; }
int fl_filter_update(fl_filter_t* f, const float* body, int nq, double* x26, double* P, double R, double* solve_time_s) {
    FILTER_GUARD(f);
    return f->impl->update(body, nq, x26, P, R, solve_time_s);
}
int fl_filter_map_incremental(fl_filter_t* f, double fsm, int ekf_inited, int* out3) {
    FILTER_GUARD(f);
    int a = 0, b = 0, c = 0;
    int rc = f->impl->map_incremental(fsm, ekf_inited, &a, &b, &c);
    if (out3) { out3[0] = a; out3[1] = b; out3[2] = c; }
    return rc;
}
int fl_filter_get_nearest(fl_filter_t* f, float* out_pts, int* out_cnt, int nq) { FILTER_GUARD(f); return f->impl->get_nearest(out_pts, out_cnt, nq); }
int fl_filter_get_selected(fl_filter_t* f, unsigned char* out, int nq) { FILTER_GUARD(f); if (!out) return FL_ERR_ARG; return f->impl->get_selected(out, nq); }
int fl_filter_get_pass_logs(fl_filter_t* f, fl_pass_log_t* out, int cap) {
    FILTER_GUARD(f);
    if (!out || cap < 0) return FL_ERR_ARG;
    int n = 0;
    int rc = f->impl->get_pass_logs(reinterpret_cast<fl::PassLog*>(out), cap, &n);
    return rc == FL_OK ? n : rc;
}
int fl_filter_upload_scan(fl_filter_t* f, const float* body, int nq) { FILTER_GUARD(f); return f->impl->upload_scan(body, nq); }
int fl_filter_upload_state(fl_filter_t* f, const double* x26, const double* P, double R) {
    FILTER_GUARD(f);
    if (!x26 || !P) return FL_ERR_ARG;
    return f->impl->upload_state(x26, P, R);
}
int fl_filter_run(fl_filter_t* f) { FILTER_GUARD(f); return f->impl->run_passes(); }
int fl_filter_download_state(fl_filter_t* f, double* x26, double* P, int* n_pass) { FILTER_GUARD(f); return f->impl->download_state(x26, P, n_pass); }
int fl_filter_sync(fl_filter_t* f) { FILTER_GUARD(f); return f->impl->sync(); }
int fl_filter_debug_prof(fl_filter_t* f, long long* out16) {
    FILTER_GUARD(f);
    if (!out16) return FL_ERR_ARG;
    FL_CUDA(cudaMemcpy(out16, f->impl->ctl_device()->prof, sizeof(long long) * 16, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; i++) out16[12 + i] = f->impl->host_ns()[i];       // host-side ns of the last fl_filter_update
    return FL_OK;
}
int fl_filter_gpu_launches(fl_filter_t* f) { FILTER_GUARD(f); return f->impl->gpu_launches(); }

int fl_filter_time_resident(fl_filter_t* f, int reps, int flush_l2, float* ms_total) {
    FILTER_GUARD(f);
    if (reps < 1 || !ms_total) return FL_ERR_ARG;
    fl::Filter* F = f->impl;
    cudaStream_t st = F->stream();
    FL_CUDA(cudaSetDevice(F->map()->device()));
    const size_t flush_bytes = 256u << 20;       // > 126 MB of L2
    if (flush_l2) FL_CHECK(f->flush.reserve(flush_bytes));
    cudaEvent_t e0, e1;
    FL_CUDA(cudaEventCreate(&e0));
    FL_CUDA(cudaEventCreate(&e1));
    float total = 0.f;
    for (int r = 0; r < reps; r++) {
        FL_CHECK(F->restore_state());
        if (flush_l2) FL_CUDA(cudaMemsetAsync(f->flush.ptr, r & 0xff, flush_bytes, st));
        FL_CHECK(F->p2p_barrier());           // multi-GPU: all ranks enter the timed step together (no-op on one GPU)
        FL_CUDA(cudaEventRecord(e0, st));
        int rc = F->run_passes();
        if (rc != FL_OK) { cudaEventDestroy(e0); cudaEventDestroy(e1); return rc; }
        FL_CUDA(cudaEventRecord(e1, st));
        FL_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        FL_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        total += ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_total = total;
    return FL_OK;
}

// Wall time of `reps` whole fl_filter_update calls made back to back from native code -- what a C++ application
// (the reference is one) sees per scan: host buffers in, host buffers out, every copy inside.  Each repetition starts
// from the same prior, like the bench's resident loop.
int fl_filter_time_e2e(fl_filter_t* f, const float* body, int nq, const double* x26, const double* P, double R, int reps,
                       double* seconds, double* x26_out, double* P_out) {
    if (!f || !x26 || !P || !seconds || reps < 1) return FL_ERR_ARG;
    double x[fl::XLEN], Pw[fl::NDOF * fl::NDOF];
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; r++) {
        memcpy(x, x26, sizeof(x));
        memcpy(Pw, P, sizeof(Pw));
        int rc = fl_filter_update(f, body, nq, x, Pw, R, nullptr);
        if (rc != FL_OK) return rc;
    }
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (x26_out) memcpy(x26_out, x, sizeof(x));
    if (P_out) memcpy(P_out, Pw, sizeof(Pw));
    return FL_OK;
}

// Device time of the dominant phase alone: the kNN of one searching pass (k_update's search_only launch; k_search_c / k_search on the legacy path).
int fl_filter_time_search_pass(fl_filter_t* f, int reps, int flush_l2, float* ms_total) {
    FILTER_GUARD(f);
    if (reps < 1 || !ms_total) return FL_ERR_ARG;
    fl::Filter* F = f->impl;
    cudaStream_t st = F->stream();
    FL_CUDA(cudaSetDevice(F->map()->device()));
    const size_t flush_bytes = 256u << 20;
    if (flush_l2) FL_CHECK(f->flush.reserve(flush_bytes));
    cudaEvent_t e0, e1;
    FL_CUDA(cudaEventCreate(&e0));
    FL_CUDA(cudaEventCreate(&e1));
    float total = 0.f;
    for (int r = 0; r < reps; r++) {
        FL_CHECK(F->restore_state());
        if (flush_l2) FL_CUDA(cudaMemsetAsync(f->flush.ptr, r & 0xff, flush_bytes, st));
        FL_CUDA(cudaEventRecord(e0, st));
        int rc = F->launch_search_only();
        if (rc != FL_OK) { cudaEventDestroy(e0); cudaEventDestroy(e1); return rc; }
        FL_CUDA(cudaEventRecord(e1, st));
        FL_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        FL_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        total += ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    // leave the control block as uploaded
    FL_CHECK(F->restore_state());
    *ms_total = total;
    return FL_OK;
}

// ------------------------------------------------------------------------------------ scan front end
#define SCAN_GUARD(s)                                                                        \
    if (!(s) || !(s)->impl) { fl::set_last_error("null scan handle"); return FL_ERR_ARG; }   \
    std::lock_guard<std::mutex> _lk((s)->map->mu)

int fl_scan_create(fl_scan_t** out, fl_map_t* map) {
    if (!out) return FL_ERR_ARG;
    *out = nullptr;
    if (!map || !map->impl) { fl::set_last_error("fl_scan_create: null map handle"); return FL_ERR_ARG; }
    fl_scan* s = new (std::nothrow) fl_scan();
    if (!s) return FL_ERR_CAPACITY;
    s->map = map;
    s->impl = new (std::nothrow) fl::ScanFrontEnd(map->impl);
    if (!s->impl) { delete s; return FL_ERR_CAPACITY; }
    int rc = s->impl->init();
    if (rc != FL_OK) { delete s->impl; delete s; return rc; }
    map_retain(map);
    *out = s;
    return FL_OK;
}
int fl_scan_destroy(fl_scan_t* s) {
    if (!s) return FL_OK;
    delete s->impl;
    map_release(s->map);
    delete s;
    return FL_OK;
}
int fl_scan_upload(fl_scan_t* s, const float* xyzi, const float* offset_ms, int n) { SCAN_GUARD(s); return s->impl->upload(xyzi, offset_ms, n); }
int fl_scan_undistort(fl_scan_t* s, const double* imu_pose22, int n_pose, const double* x26_end) {
    SCAN_GUARD(s);
    return s->impl->undistort(imu_pose22, n_pose, x26_end);
}
int fl_scan_voxel_downsample(fl_scan_t* s, float leaf_size) {
    SCAN_GUARD(s);
    int n = 0;
    int rc = s->impl->voxel_downsample(leaf_size, &n);
    return rc == FL_OK ? n : rc;
}
int fl_scan_download(fl_scan_t* s, int which, float* out_xyzi, int cap) {
    SCAN_GUARD(s);
    int n = 0;
    int rc = s->impl->download(which, out_xyzi, cap, &n);
    return rc == FL_OK ? n : rc;
}
int fl_filter_update_scan(fl_filter_t* f, fl_scan_t* s, double* x26, double* P, double R, double* solve_time_s) {
    FILTER_GUARD(f);
    if (!s || !s->impl || s->map != f->map) { fl::set_last_error("fl_filter_update_scan: the scan must live on the filter's map"); return FL_ERR_ARG; }
    return f->impl->update_device(s->impl->down_device(), s->impl->down_count(), x26, P, R, solve_time_s);
}

// ------------------------------------------------------------------------------------ local-map cube
int fl_localmap_create(fl_localmap_t** out, double cube_len, float det_range) {
    if (!out) return FL_ERR_ARG;
    *out = nullptr;
    if (!(cube_len > 0.0) || !(det_range > 0.f)) { fl::set_last_error("fl_localmap_create: cube_len and det_range must be > 0"); return FL_ERR_ARG; }
    fl_localmap* l = new (std::nothrow) fl_localmap{fl::LocalMapCube(cube_len, det_range)};
    if (!l) return FL_ERR_CAPACITY;
    *out = l;
    return FL_OK;
}
int fl_localmap_destroy(fl_localmap_t* l) { delete l; return FL_OK; }
int fl_localmap_segment(fl_localmap_t* l, fl_map_t* map, const double* pos_lid, float* boxes6_out, int* n_deleted) {
    if (n_deleted) *n_deleted = 0;
    if (!l || !pos_lid) { fl::set_last_error("fl_localmap_segment: null argument"); return FL_ERR_ARG; }
    float boxes[18];
    const int nb = l->cube.slide(pos_lid, boxes);
    if (boxes6_out) for (int i = 0; i < nb * 6; i++) boxes6_out[i] = boxes[i];
    if (nb > 0 && map) {                              // if (cub_needrm.size() > 0) ikdtree.Delete_Point_Boxes(cub_needrm)  (:275)
        int deleted = fl_map_delete_boxes(map, boxes, nb);
        if (deleted < 0) return deleted;
        if (n_deleted) *n_deleted = deleted;
    }
    return nb;
}
int fl_localmap_get(fl_localmap_t* l, float* box6) {
    if (!l || !box6) return FL_ERR_ARG;
    if (!l->cube.initialized()) { fl::set_last_error("fl_localmap_get: the cube is placed by the first fl_localmap_segment call"); return FL_ERR_STATE; }
    l->cube.get(box6);
    return FL_OK;
}

// ------------------------------------------------------------------------------------ multi-GPU
int fl_comm_unique_id(void* out128) { if (!out128) return FL_ERR_ARG; return fl::nccl_unique_id(out128); }
int fl_filter_comm_init(fl_filter_t* f, int nranks, int rank, const void* id128) {
    FILTER_GUARD(f);
    if (nranks > 1 && !id128) return FL_ERR_ARG;
    return f->impl->comm_init(nranks, rank, id128);
}
int fl_filter_p2p_handle(fl_filter_t* f, void* out64) { FILTER_GUARD(f); if (!out64) return FL_ERR_ARG; return f->impl->p2p_local_handle(out64); }
int fl_filter_p2p_connect(fl_filter_t* f, int nranks, int rank, const void* handles) { FILTER_GUARD(f); return f->impl->p2p_connect(nranks, rank, handles); }
int fl_filter_set_shard(fl_filter_t* f, int q_begin, int q_end) { FILTER_GUARD(f); return f->impl->set_shard(q_begin, q_end); }

}  // extern "C"
