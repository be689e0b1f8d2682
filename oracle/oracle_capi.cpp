// oracle_capi.cpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// C entry points over the CPU oracle, shaped like include/vmp_b200.h (same POD structs,
// prefix orc_ instead of vmp_) so that tests and bench.py's cpu_baseline leg can drive
// oracle and CUDA path with the same inputs.
#include <chrono>
#include <cstring>

#include "../include/vmp_b200.h"
#include "oracle.h"

using namespace orc;

namespace {
struct Handle {
    LIOBuilder lio;
};
inline Handle* H(void* h) { return reinterpret_cast<Handle*>(h); }

LIOConfig to_cfg(const vmp_config* c) {
    LIOConfig o;
    o.opti_max_iter = c->opti_max_iter;
    o.na = c->na; o.ng = c->ng; o.nba = c->nba; o.nbg = c->nbg;
    o.imu_init_num = c->imu_init_num;
    std::memcpy(o.r_il.a, c->r_il, sizeof(double) * 9);
    std::memcpy(o.p_il.a, c->p_il, sizeof(double) * 3);
    o.gravity_align = c->gravity_align != 0;
    o.estimate_ext = c->estimate_ext != 0;
    o.scan_resolution = c->scan_resolution;
    o.voxel_size = c->voxel_size;
    o.update_size_thresh = c->update_size_thresh;
    o.max_point_thresh = c->max_point_thresh;
    o.plane_thresh = c->plane_thresh;
    o.ranging_cov = c->ranging_cov; o.angle_cov = c->angle_cov;
    o.merge_thresh_for_angle = c->merge_thresh_for_angle;
    o.merge_thresh_for_distance = c->merge_thresh_for_distance;
    o.map_capacity = c->map_capacity;
    o.max_points_per_scan = c->max_points_per_scan;
    return o;
}
void to_state(const vmp_state* s, State& x) {
    std::memcpy(x.pos.a, s->pos, 24); std::memcpy(x.rot.a, s->rot, 72); std::memcpy(x.rot_ext.a, s->rot_ext, 72);
    std::memcpy(x.pos_ext.a, s->pos_ext, 24); std::memcpy(x.vel.a, s->vel, 24); std::memcpy(x.bg.a, s->bg, 24);
    std::memcpy(x.ba.a, s->ba, 24); std::memcpy(x.g.a, s->g, 24);
}
void from_state(const State& x, vmp_state* s) {
    std::memcpy(s->pos, x.pos.a, 24); std::memcpy(s->rot, x.rot.a, 72); std::memcpy(s->rot_ext, x.rot_ext.a, 72);
    std::memcpy(s->pos_ext, x.pos_ext.a, 24); std::memcpy(s->vel, x.vel.a, 24); std::memcpy(s->bg, x.bg.a, 24);
    std::memcpy(s->ba, x.ba.a, 24); std::memcpy(s->g, x.g.a, 24);
}
void fill_stats(const VoxelMap& m, vmp_update_stats* st) {
    if (!st) return;
    const MapCounters& c = m.counters;
    st->n_points = c.n_points; st->n_ins = c.n_ins; st->n_touch = c.n_touch; st->n_created = c.n_created;
    st->n_refit = c.n_refit; st->refit_points = c.refit_points; st->n_full = c.n_full;
    st->n_mergeprobe = c.n_mergeprobe; st->n_merge = c.n_merge; st->n_evicted = c.n_evicted;
    st->map_size = (int64_t)m.featmap.size();
    st->n_mergevox = c.n_mergevox;
    st->n_skipped = 0;                        // the reference inserts every point (int64 keys)
}
std::vector<CloudPoint> to_cloud3(const float* p, int n) {
    std::vector<CloudPoint> c((size_t)n);
    for (int i = 0; i < n; i++) { c[i].x = p[3 * i]; c[i].y = p[3 * i + 1]; c[i].z = p[3 * i + 2]; c[i].curvature = 0.f; }
    return c;
}
std::vector<PointWithCov> to_pvs(const double* pts, const double* cov, int n) {
    std::vector<PointWithCov> v((size_t)n);
    for (int i = 0; i < n; i++) {
        std::memcpy(v[i].point.a, pts + 3 * i, 24);
        std::memcpy(v[i].cov.a, cov + 9 * i, 72);
    }
    return v;
}
void fill_scan_stats(LIOBuilder& l, vmp_scan_stats* st, double ms) {
    if (!st) return;
    std::memset(st, 0, sizeof(*st));
    st->iters = l.kf.last_iters;
    st->converged = l.kf.last_converged ? 1 : 0;
    for (size_t i = 0; i < l.effect_nums.size() && i < 8; i++) st->effect_num[i] = l.effect_nums[i];
    fill_stats(*l.map, &st->map);
    st->gpu_ms = (float)ms;    // for the oracle: CPU milliseconds of the timed region
}
}  // namespace

extern "C" {

int orc_create(const vmp_config* cfg, void** out) {
    if (!cfg || !out) return VMP_ERR_INVALID_ARG;
    Handle* h = new Handle();
    h->lio.loadConfig(to_cfg(cfg));
    *out = h;
    return VMP_OK;
}
int orc_destroy(void* h) { delete H(h); return VMP_OK; }
int orc_set_threads(void* h, int n) { H(h)->lio.omp_threads = n < 1 ? 1 : n; return VMP_OK; }

int orc_map_build(void* h, const double* pts, const double* cov, int n, vmp_update_stats* st) {
    std::vector<PointWithCov> v = to_pvs(pts, cov, n);
    H(h)->lio.map->build(v);
    fill_stats(*H(h)->lio.map, st);
    return VMP_OK;
}
int orc_map_update(void* h, const double* pts, const double* cov, int n, vmp_update_stats* st) {
    std::vector<PointWithCov> v = to_pvs(pts, cov, n);
    H(h)->lio.map->update(v);
    fill_stats(*H(h)->lio.map, st);
    return VMP_OK;
}
// timed variant for the map-update microbench: returns CPU seconds spent inside update()
double orc_map_update_timed(void* h, const double* pts, const double* cov, int n, vmp_update_stats* st) {
    std::vector<PointWithCov> v = to_pvs(pts, cov, n);
    auto t0 = std::chrono::steady_clock::now();
    H(h)->lio.map->update(v);
    auto t1 = std::chrono::steady_clock::now();
    fill_stats(*H(h)->lio.map, st);
    return std::chrono::duration<double>(t1 - t0).count();
}

int orc_set_state(void* h, const vmp_state* x, const double* P) {
    if (x) to_state(x, H(h)->lio.kf.x());
    if (P) std::memcpy(H(h)->lio.kf.P().a, P, sizeof(double) * 529);
    return VMP_OK;
}
int orc_get_state(void* h, vmp_state* x, double* P) {
    if (x) from_state(H(h)->lio.kf.x(), x);
    if (P) std::memcpy(P, H(h)->lio.kf.P().a, sizeof(double) * 529);
    return VMP_OK;
}

int orc_set_scan(void* h, const float* pts, int n) {
    if (n > H(h)->lio.config.max_points_per_scan) return VMP_ERR_INVALID_ARG;
    H(h)->lio.setScan(to_cloud3(pts, n));
    return VMP_OK;
}
int orc_measure(void* h, const vmp_state* x, const double* P, double* Hout, double* bout, int* effect_num) {
    LIOBuilder& l = H(h)->lio;
    State s; to_state(x, s);
    std::memcpy(l.kf.P().a, P, sizeof(double) * 529);
    SharedState sh;
    l.effect_nums.clear();
    l.sharedUpdateFunc(s, sh);
    if (Hout) std::memcpy(Hout, sh.H.a, sizeof(double) * 144);
    if (bout) std::memcpy(bout, sh.b.a, sizeof(double) * 12);
    if (effect_num) *effect_num = l.effect_nums.back();
    return VMP_OK;
}
int orc_scan(void* h, vmp_state* x, double* P, const float* pts, int n, vmp_scan_stats* st) {
    LIOBuilder& l = H(h)->lio;
    if (n > l.config.max_points_per_scan) return VMP_ERR_INVALID_ARG;
    orc_set_state(h, x, P);
    std::vector<CloudPoint> c = to_cloud3(pts, n);
    auto t0 = std::chrono::steady_clock::now();
    l.setScan(c);
    l.hotPath();
    auto t1 = std::chrono::steady_clock::now();
    orc_get_state(h, x, P);
    fill_scan_stats(l, st, std::chrono::duration<double, std::milli>(t1 - t0).count());
    return VMP_OK;
}
int orc_first_scan(void* h, const vmp_state* x, const double* P, const float* pts, int n, vmp_update_stats* st) {
    LIOBuilder& l = H(h)->lio;
    orc_set_state(h, x, P);
    l.firstScan(to_cloud3(pts, n));
    l.status = LIO_MAPPING;
    fill_stats(*l.map, st);
    return VMP_OK;
}

int orc_dump_correspondences(void* h, int64_t* keys, uint8_t* status, double* residual, double* plane_norm, int n) {
    std::vector<ResidualData>& ri = H(h)->lio.data_group.residual_info;
    if ((size_t)n > ri.size()) return VMP_ERR_INVALID_ARG;
    for (int i = 0; i < n; i++) {
        if (keys) { keys[3 * i] = ri[i].key.x; keys[3 * i + 1] = ri[i].key.y; keys[3 * i + 2] = ri[i].key.z; }
        if (status) status[i] = ri[i].status;
        if (residual) residual[i] = ri[i].residual;
        if (plane_norm) std::memcpy(plane_norm + 3 * i, ri[i].plane_norm.a, 24);
    }
    return VMP_OK;
}
int orc_dump_world_points(void* h, double* pts, double* cov, int n) {
    std::vector<PointWithCov>& pv = H(h)->lio.last_pv_list;
    if ((size_t)n > pv.size()) return VMP_ERR_INVALID_ARG;
    for (int i = 0; i < n; i++) {
        if (pts) std::memcpy(pts + 3 * i, pv[i].point.a, 24);
        if (cov) std::memcpy(cov + 9 * i, pv[i].cov.a, 72);
    }
    return VMP_OK;
}
int orc_dump_map(void* h, vmp_plane* out, int cap, int* count) {
    VoxelMap& m = *H(h)->lio.map;
    int i = 0;
    for (const VoxelKey& k : m.cache) {     // front = most recently inserted-into (utils.cpp:161)
        if (i >= cap) break;
        const VoxelGrid& g = *m.featmap.at(k);
        vmp_plane& p = out[i];
        std::memset(&p, 0, sizeof(p));
        p.key[0] = k.x; p.key[1] = k.y; p.key[2] = k.z;
        std::memcpy(p.mean, g.plane->mean.a, 24);
        std::memcpy(p.ppt, g.plane->ppt.a, 72);
        std::memcpy(p.norm, g.plane->norm.a, 24);
        std::memcpy(p.cov, g.plane->cov.a, 288);
        std::memcpy(p.center, g.center.a, 24);
        p.n = g.plane->n;
        p.n_temp = (int32_t)g.temp_points.size();
        p.newly_add_point = g.newly_add_point;
        p.flags = (g.is_init ? VMP_F_INIT : 0) | (g.is_plane ? VMP_F_PLANE : 0) |
                  (g.update_enable ? VMP_F_UPDATE_ENABLE : 0) | (g.merged ? VMP_F_MERGED : 0);
        p.group = g.group_id;
        p.lru_rank = (uint64_t)i;
        i++;
    }
    if (count) *count = (int)m.cache.size();
    return VMP_OK;
}
int orc_dump_evicted(void* h, int64_t* keys, int cap, int* count) {
    VoxelMap& m = *H(h)->lio.map;
    int n = (int)m.evicted.size();
    for (int i = 0; i < n && i < cap; i++) { keys[3 * i] = m.evicted[i].x; keys[3 * i + 1] = m.evicted[i].y; keys[3 * i + 2] = m.evicted[i].z; }
    if (count) *count = n;
    return VMP_OK;
}
int orc_map_size(void* h, int* count) { *count = (int)H(h)->lio.map->featmap.size(); return VMP_OK; }

int orc_lio_process(void* h, const vmp_imu* imus, int n_imu, float* cloud_xyzc, int n, double t0, double t1, vmp_scan_stats* st) {
    LIOBuilder& l = H(h)->lio;
    SyncPackage pk;
    pk.imus.resize((size_t)n_imu);
    for (int i = 0; i < n_imu; i++) {
        std::memcpy(pk.imus[i].acc.a, imus[i].acc, 24);
        std::memcpy(pk.imus[i].gyro.a, imus[i].gyro, 24);
        pk.imus[i].timestamp = imus[i].timestamp;
    }
    pk.cloud.resize((size_t)n);
    for (int i = 0; i < n; i++) { pk.cloud[i].x = cloud_xyzc[4 * i]; pk.cloud[i].y = cloud_xyzc[4 * i + 1]; pk.cloud[i].z = cloud_xyzc[4 * i + 2]; pk.cloud[i].curvature = cloud_xyzc[4 * i + 3]; }
    pk.cloud_start_time = t0; pk.cloud_end_time = t1;
    double ms = 0.0;
    if (l.status == LIO_MAPPING) {
        // same as process(), with the clock around the timed region lio_builder.cpp:224-246
        l.undistortCloud(pk);
        if (n > l.config.max_points_per_scan) return VMP_ERR_INVALID_ARG;
        std::vector<CloudPoint> filtered;                 // scan_filter.filter (lio_builder.cpp:215-219), outside the timed region
        const bool ds = l.config.scan_resolution > 0.0;
        if (ds) filtered = LIOBuilder::voxelGridFilter(pk.cloud, (float)l.config.scan_resolution);
        auto c0 = std::chrono::steady_clock::now();
        l.setScan(ds ? filtered : pk.cloud);
        l.hotPath();
        auto c1 = std::chrono::steady_clock::now();
        ms = std::chrono::duration<double, std::milli>(c1 - c0).count();
    } else {
        l.process(pk);
    }
    for (int i = 0; i < n; i++) { cloud_xyzc[4 * i] = pk.cloud[i].x; cloud_xyzc[4 * i + 1] = pk.cloud[i].y; cloud_xyzc[4 * i + 2] = pk.cloud[i].z; cloud_xyzc[4 * i + 3] = pk.cloud[i].curvature; }
    fill_scan_stats(l, st, ms);
    return VMP_OK;
}
// pcl::VoxelGrid leaf-centroid filter on an N x 4 float cloud (x y z curvature); out: up to cap points, *m = leaf count
int orc_downsample(void* h, const float* cloud_xyzc, int n, double leaf, float* out_xyzc, int cap, int* m) {
    (void)h;
    std::vector<CloudPoint> c((size_t)n);
    for (int i = 0; i < n; i++) c[i] = CloudPoint{cloud_xyzc[4 * i], cloud_xyzc[4 * i + 1], cloud_xyzc[4 * i + 2], cloud_xyzc[4 * i + 3]};
    const std::vector<CloudPoint> o = LIOBuilder::voxelGridFilter(c, (float)leaf);
    for (size_t i = 0; i < o.size() && (int)i < cap; i++) { out_xyzc[4 * i] = o[i].x; out_xyzc[4 * i + 1] = o[i].y; out_xyzc[4 * i + 2] = o[i].z; out_xyzc[4 * i + 3] = o[i].curvature; }
    if (m) *m = (int)o.size();
    return VMP_OK;
}
// LIOBuilder::lidar_cloud of the last scan (the filter input; downsampled when scan_resolution > 0)
int orc_get_lidar_cloud(void* h, float* out_xyzc, int cap, int* m) {
    const std::vector<CloudPoint>& o = H(h)->lio.lidar_cloud;
    for (size_t i = 0; i < o.size() && (int)i < cap; i++) { out_xyzc[4 * i] = o[i].x; out_xyzc[4 * i + 1] = o[i].y; out_xyzc[4 * i + 2] = o[i].z; out_xyzc[4 * i + 3] = o[i].curvature; }
    if (m) *m = (int)o.size();
    return VMP_OK;
}
// smallest margins of the gate decisions since creation: plane fit, 3-sigma gate, merge angle, merge distance
int orc_track_margins(void* h, int on) { H(h)->lio.map->track_margins = on != 0; return VMP_OK; }
int orc_gate_margins(void* h, double* out4) {
    const GateMargins& g = H(h)->lio.map->margins;
    out4[0] = g.plane; out4[1] = g.gate; out4[2] = g.merge_angle; out4[3] = g.merge_dist;
    return VMP_OK;
}
int orc_lio_state(void* h, vmp_state* x, double* P, int* status) {
    orc_get_state(h, x, P);
    if (status) *status = (int)H(h)->lio.status;
    return VMP_OK;
}

// prior of the last hotPath()/firstScan() and the state of measurement call k of the last update()
int orc_get_prior(void* h, vmp_state* x, double* P) {
    if (x) from_state(H(h)->lio.prior_x, x);
    if (P) std::memcpy(P, H(h)->lio.prior_P.a, sizeof(double) * 529);
    return VMP_OK;
}
int orc_get_iter_state(void* h, int k, vmp_state* x) {
    if (k < 0 || (size_t)k >= H(h)->lio.iter_states.size()) return VMP_ERR_INVALID_ARG;
    from_state(H(h)->lio.iter_states[k], x);
    return VMP_OK;
}

int orc_get_iter_Hb(void* h, int k, double* Hout, double* bout) {
    if (k < 0 || (size_t)k >= H(h)->lio.iter_H.size()) return VMP_ERR_INVALID_ARG;
    std::memcpy(Hout, H(h)->lio.iter_H[k].a, sizeof(double) * 144);
    std::memcpy(bout, H(h)->lio.iter_b[k].a, sizeof(double) * 12);
    return VMP_OK;
}

// ---- small-math exports for unit tests against numpy ----
void orc_eig3(const double* A, double* evals, double* evecs) {
    M3 a; std::memcpy(a.a, A, 72);
    M3 v; eig3_sym(a, evals, v);
    std::memcpy(evecs, v.a, 72);
}
void orc_inverse23(const double* A, double* out) {
    M23 a; std::memcpy(a.a, A, sizeof(double) * 529);
    M23 r = inverse(a);
    std::memcpy(out, r.a, sizeof(double) * 529);
}
void orc_so3_exp(const double* w, double* R) { V3 v = v3(w[0], w[1], w[2]); M3 r = so3_exp(v); std::memcpy(R, r.a, 72); }
void orc_so3_log(const double* R, double* w) { M3 r; std::memcpy(r.a, R, 72); V3 v = so3_log(r); std::memcpy(w, v.a, 24); }
void orc_left_jacobian(const double* w, double* J) { M3 r = so3_left_jacobian(v3(w[0], w[1], w[2])); std::memcpy(J, r.a, 72); }
void orc_calc_body_cov(double* pb, double range_inc, double degree_inc, double* cov) {
    V3 p = v3(pb[0], pb[1], pb[2]); M3 c;
    calcBodyCov(p, range_inc, degree_inc, c);
    std::memcpy(pb, p.a, 24); std::memcpy(cov, c.a, 72);
}
void orc_boxplus(vmp_state* x, const double* delta23) {
    State s; to_state(x, s); V23 d; std::memcpy(d.a, delta23, sizeof(double) * 23);
    s.plus(d); from_state(s, x);
}
void orc_boxminus(const vmp_state* a, const vmp_state* b, double* delta23) {
    State sa, sb; to_state(a, sa); to_state(b, sb);
    V23 d = sa.minus(sb); std::memcpy(delta23, d.a, sizeof(double) * 23);
}
void orc_predict(void* h, const double* acc, const double* gyro, double dt) {
    Input in; std::memcpy(in.acc.a, acc, 24); std::memcpy(in.gyro.a, gyro, 24);
    H(h)->lio.kf.predict(in, dt, H(h)->lio.data_group.Q);
}
void orc_voxel_key(void* h, const double* p, int64_t* key) {
    VoxelKey k = H(h)->lio.map->index(v3(p[0], p[1], p[2]));
    key[0] = k.x; key[1] = k.y; key[2] = k.z;
}

// ---- STDManager::buildVoxels (std_matcher/src/std_manager/descriptor.cpp:70-122), restated ----
// Pass 1 (:72-89): every point, in cloud order, into temp_voxels[VoxelKey::index(p, voxel_size)] (:42-47): sum += p, ppt += p p^T, count.
// Pass 2 (:91-121): voxels with count > voxel_min_point get mean = sum / n, cov = ppt / n - mean mean^T, an eigen-decomposition and
// is_plane = lambda_min < voxel_plane_thresh; plane voxels keep lamdas (min, mid, max) and the eigenvectors as norms' columns.
// The reference iterates an unordered_map (order unobservable): records come out in first-touch order.  Eigen::EigenSolver (general
// real Schur) is replaced by the symmetric solver of this oracle; tests/test_std_voxels.py pins eigenvalues and eigenvectors (up to
// sign) against LAPACK's general solver (numpy.linalg.eig, dgeev), which is the algorithm class EigenSolver implements.
int orc_std_build_voxels(const float* cloud_xyzi, int n, double voxel_size, int voxel_min_point, double voxel_plane_thresh,
                         vmp_std_voxel* out, int cap, int* count) {
    struct Node { V3 sum = V3::zero(); M3 ppt = M3::zero(); int n = 0; int order = 0; };
    std::unordered_map<VoxelKey, Node, VoxelKey::Hasher> temp_voxels;
    std::vector<VoxelKey> order;
    for (int i = 0; i < n; i++) {
        const float* p = cloud_xyzi + 4 * (size_t)i;
        const double x = (double)p[0], y = (double)p[1], z = (double)p[2];
        VoxelKey k{(int64_t)std::floor(x / voxel_size + 0.0), (int64_t)std::floor(y / voxel_size + 0.0), (int64_t)std::floor(z / voxel_size + 0.0)};
        auto it = temp_voxels.find(k);
        if (it == temp_voxels.end()) {
            it = temp_voxels.emplace(k, Node{}).first;
            it->second.order = (int)order.size();
            order.push_back(k);
        }
        Node& v = it->second;
        V3 pv = v3(x, y, z);
        v.sum = add(v.sum, pv);
        v.ppt = add(v.ppt, outer(pv, pv));
        v.n++;
    }
    if (count) *count = (int)order.size();
    for (int i = 0; i < (int)order.size() && i < cap; i++) {
        const Node& v = temp_voxels[order[i]];
        vmp_std_voxel& o = out[i];
        std::memset(&o, 0, sizeof(o));
        o.key[0] = order[i].x; o.key[1] = order[i].y; o.key[2] = order[i].z;
        o.count = v.n;
        for (int k = 0; k < 3; k++) o.sum[k] = v.sum.a[k];
        for (int k = 0; k < 9; k++) o.ppt[k] = v.ppt.a[k];
        if (!(v.n > voxel_min_point)) continue;
        o.flags |= VMP_STD_F_VALID;
        const double nd = (double)v.n;
        V3 mean = divs(v.sum, nd);
        M3 cov = sub(divs(v.ppt, nd), outer(mean, mean));
        double ev[3];
        M3 evec;
        eig3_sym(cov, ev, evec);
        for (int k = 0; k < 3; k++) o.mean[k] = mean.a[k];
        if (ev[0] < voxel_plane_thresh) {
            o.flags |= VMP_STD_F_PLANE;
            for (int k = 0; k < 3; k++) { o.lamdas[k] = ev[k]; for (int q = 0; q < 3; q++) o.norms[3 * k + q] = evec(q, k); }
        }
    }
    return VMP_OK;
}

}  // extern "C"
