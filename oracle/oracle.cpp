// oracle.cpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md, oracle.h).
// CPU restatement of the reference hot path; every function cites what it follows.
#include "oracle.h"

#include <cfloat>
#include <climits>

#include <algorithm>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

// =============================================================== ieskf.cpp

// ieskf.cpp:11-21
void State::plus(const V23& d) {
    pos = add(pos, v3(d[0], d[1], d[2]));
    rot = mul(rot, so3_exp(v3(d[3], d[4], d[5])));
    rot_ext = mul(rot_ext, so3_exp(v3(d[6], d[7], d[8])));
    pos_ext = add(pos_ext, v3(d[9], d[10], d[11]));
    vel = add(vel, v3(d[12], d[13], d[14]));
    bg = add(bg, v3(d[15], d[16], d[17]));
    ba = add(ba, v3(d[18], d[19], d[20]));
    V2 dg; dg[0] = d[21]; dg[1] = d[22];
    g = mul(so3_exp(mul(getBx(), dg)), g);
}

// ieskf.cpp:23-33
void State::plus24(const V24& d) {
    pos = add(pos, v3(d[0], d[1], d[2]));
    rot = mul(rot, so3_exp(v3(d[3], d[4], d[5])));
    rot_ext = mul(rot_ext, so3_exp(v3(d[6], d[7], d[8])));
    pos_ext = add(pos_ext, v3(d[9], d[10], d[11]));
    vel = add(vel, v3(d[12], d[13], d[14]));
    bg = add(bg, v3(d[15], d[16], d[17]));
    ba = add(ba, v3(d[18], d[19], d[20]));
    g = mul(so3_exp(v3(d[21], d[22], d[23])), g);
}

// ieskf.cpp:35-67
V23 State::minus(const State& other) const {
    V23 delta = V23::zero();
    V3 t;
    t = sub(pos, other.pos);                      delta[0] = t[0];  delta[1] = t[1];  delta[2] = t[2];
    t = so3_log(mul(tr(other.rot), rot));         delta[3] = t[0];  delta[4] = t[1];  delta[5] = t[2];
    t = so3_log(mul(tr(other.rot_ext), rot_ext)); delta[6] = t[0];  delta[7] = t[1];  delta[8] = t[2];
    t = sub(pos_ext, other.pos_ext);              delta[9] = t[0];  delta[10] = t[1]; delta[11] = t[2];
    t = sub(vel, other.vel);                      delta[12] = t[0]; delta[13] = t[1]; delta[14] = t[2];
    t = sub(bg, other.bg);                        delta[15] = t[0]; delta[16] = t[1]; delta[17] = t[2];
    t = sub(ba, other.ba);                        delta[18] = t[0]; delta[19] = t[1]; delta[20] = t[2];

    const double v_sin = norm(mul(hat(g), other.g));
    const double v_cos = dot(g, other.g);
    const double theta = std::atan2(v_sin, v_cos);
    V2 res;
    if (v_sin < 1e-11) {
        if (std::fabs(theta) > 1e-11) { res[0] = 3.1415926; res[1] = 0.0; }
        else { res[0] = 0.0; res[1] = 0.0; }
    } else {
        // theta / v_sin * other.getBx().transpose() * hat(other.g) * g   (left to right)
        Mat<2, 3> a = scale(tr(other.getBx()), theta / v_sin);
        Mat<2, 3> b = mul(a, hat(other.g));
        res = mul(b, g);
    }
    delta[21] = res[0]; delta[22] = res[1];
    return delta;
}

// ieskf.cpp:69-77
M32 State::getBx() const {
    M32 r;
    r(0, 0) = -g[1];
    r(0, 1) = -g[2];
    r(1, 0) = GRAVITY - g[1] * g[1] / (GRAVITY + g[0]);
    r(1, 1) = -g[2] * g[1] / (GRAVITY + g[0]);
    r(2, 0) = -g[2] * g[1] / (GRAVITY + g[0]);
    r(2, 1) = GRAVITY - g[2] * g[2] / (GRAVITY + g[0]);
    return divs(r, GRAVITY);
}
// ieskf.cpp:79-83
M32 State::getMx() const { return mul(neg(hat(g)), getBx()); }
// ieskf.cpp:85-90
M32 State::getMx(const V2& res) const {
    M32 bx = getBx();
    V3 bu = mul(bx, res);
    M3 a = mul(neg(so3_exp(bu)), hat(g));
    M3 b = mul(a, tr(so3_left_jacobian(bu)));
    return mul(b, bx);
}
// ieskf.cpp:92-95
M23x State::getNx() const {
    const double s = 1 / GRAVITY / GRAVITY;
    return mul(scale(tr(getBx()), s), hat(g));
}

// ieskf.cpp:101-123
void IESKF::predict(const Input& inp, double dt, const M12& Q) {
    V24 delta = V24::zero();
    V3 t = scale(x_.vel, dt);
    delta[0] = t[0]; delta[1] = t[1]; delta[2] = t[2];
    const V3 w = sub(inp.gyro, x_.bg);
    const V3 a = sub(inp.acc, x_.ba);
    t = scale(w, dt);
    delta[3] = t[0]; delta[4] = t[1]; delta[5] = t[2];
    t = scale(add(mul(x_.rot, a), x_.g), dt);
    delta[12] = t[0]; delta[13] = t[1]; delta[14] = t[2];

    M23 F = M23::identity();
    set_block(F, 0, 12, scale(M3::identity(), dt));
    set_block(F, 3, 3, so3_exp(scale(neg(w), dt)));
    set_block(F, 3, 15, scale(neg(right_jacobian(scale(w, dt))), dt));
    set_block(F, 12, 3, scale(mul(neg(x_.rot), hat(a)), dt));
    set_block(F, 12, 18, scale(neg(x_.rot), dt));
    set_block(F, 12, 21, scale(x_.getMx(), dt));
    set_block(F, 21, 21, mul(x_.getNx(), x_.getMx()));

    M23x12 G = M23x12::zero();
    set_block(G, 3, 0, scale(neg(right_jacobian(scale(w, dt))), dt));
    set_block(G, 12, 3, scale(neg(x_.rot), dt));
    set_block(G, 15, 6, scale(M3::identity(), dt));
    set_block(G, 18, 9, scale(M3::identity(), dt));
    x_.plus24(delta);
    P_ = add(mul(mul(F, P_), tr(F)), mul(mul(G, Q), tr(G)));
}

// ieskf.cpp:125-156
void IESKF::update() {
    State predict_x = x_;
    SharedState shared;
    shared.iter_num = 0;
    V23 delta = V23::zero();
    last_iters = 0;
    last_converged = false;
    for (size_t i = 0; i < max_iter_; i++) {
        func_(x_, shared);
        H_ = M23::zero();
        b_ = V23::zero();
        delta = x_.minus(predict_x);
        M23 J = M23::identity();
        set_block(J, 3, 3, right_jacobian(v3(delta[3], delta[4], delta[5])));
        set_block(J, 6, 6, right_jacobian(v3(delta[6], delta[7], delta[8])));
        V2 dg; dg[0] = delta[21]; dg[1] = delta[22];
        set_block(J, 21, 21, mul(x_.getNx(), predict_x.getMx(dg)));
        // P_.inverse() is evaluated twice per iteration in the reference (Q6); same value both times
        const M23 JtPinv = mul(tr(J), inverse(P_));
        b_ = add(b_, mul(JtPinv, delta));
        H_ = add(H_, mul(JtPinv, J));
        for (int r = 0; r < 12; r++) {
            for (int c = 0; c < 12; c++) H_(r, c) += shared.H(r, c);
            b_[r] += shared.b[r];
        }
        delta = mul(neg(inverse(H_)), b_);
        x_.plus(delta);
        shared.iter_num += 1;
        last_iters++;
        double mx = delta[0];
        for (int k = 1; k < 23; k++) if (delta[k] > mx) mx = delta[k];
        if (mx < eps_) { last_converged = true; break; }    // signed max (Q5)
    }
    M23 L = M23::identity();
    set_block(L, 3, 3, right_jacobian(v3(delta[3], delta[4], delta[5])));
    set_block(L, 6, 6, right_jacobian(v3(delta[6], delta[7], delta[8])));
    V2 dg; dg[0] = delta[21]; dg[1] = delta[22];
    set_block(L, 21, 21, mul(x_.getNx(), predict_x.getMx(dg)));
    P_ = mul(mul(L, inverse(H_)), tr(L));
}

// =============================================================== commons.cpp

// commons.cpp:18-45.  DEG2RAD is PCL's macro ((x)*0.017453293)
void calcBodyCov(V3& pb, double range_inc, double degree_inc, M3& cov) {
    if (pb[2] == 0) pb[2] = 0.001;                                  // Q16: edits the caller's point
    const double range = std::sqrt(pb[0] * pb[0] + pb[1] * pb[1] + pb[2] * pb[2]);
    const double range_var = range_inc * range_inc;
    const double sn = std::sin(degree_inc * 0.017453293);
    Mat<2, 2> direction_var = Mat<2, 2>::zero();
    direction_var(0, 0) = sn * sn;
    direction_var(1, 1) = sn * sn;
    V3 direction = normalized(pb);
    M3 direction_hat = hat(direction);
    V3 base_vector1 = normalized(v3(1.0, 1.0, -(direction[0] + direction[1]) / direction[2]));
    V3 base_vector2 = normalized(cross(base_vector1, direction));
    M32 N;
    N(0, 0) = base_vector1[0]; N(0, 1) = base_vector2[0];
    N(1, 0) = base_vector1[1]; N(1, 1) = base_vector2[1];
    N(2, 0) = base_vector1[2]; N(2, 1) = base_vector2[2];
    M32 A = mul(scale(direction_hat, range), N);
    cov = add(mul(scale(direction, range_var), tr(direction)), mul(mul(A, direction_var), tr(A)));
}

// =============================================================== voxel_map.cpp

// voxel_map.cpp:11-27
VoxelGrid::VoxelGrid(int mpt, int upt, double pth, VoxelKey pos, VoxelMap* m)
    : max_point_thresh(mpt), update_point_thresh(upt), plane_thresh(pth), position(pos) {
    merged = false;
    group_id = m->count++;              // reference: process-wide static VoxelGrid::count (Q22/Q25)
    is_init = false;
    is_plane = false;
    temp_points.reserve(max_point_thresh);
    newly_add_point = 0;
    plane = std::make_shared<Plane>();
    update_enable = true;
    map = m;
}

// voxel_map.cpp:29-34
void VoxelGrid::addToPlane(const PointWithCov& pv) {
    plane->mean = add(plane->mean, divs(sub(pv.point, plane->mean), plane->n + 1.0));
    plane->ppt = add(plane->ppt, outer(pv.point, pv.point));
    plane->n += 1;
}

// voxel_map.cpp:36-40
void VoxelGrid::addPoint(const PointWithCov& pv) {
    addToPlane(pv);
    temp_points.push_back(pv);
}

// voxel_map.cpp:42-95
void VoxelGrid::pushPoint(const PointWithCov& pv) {
    MapCounters& ct = map->counters;
    if (!is_init) {
        addToPlane(pv);
        temp_points.push_back(pv);
        ct.n_ins++;
        updatePlane();
    } else {
        if (is_plane) {
            if (update_enable) {
                addToPlane(pv);
                temp_points.push_back(pv);
                ct.n_ins++;
                newly_add_point++;
                if (newly_add_point >= update_point_thresh) {
                    updatePlane();
                    newly_add_point = 0;
                }
                if (temp_points.size() >= (size_t)max_point_thresh) {
                    update_enable = false;
                    std::vector<PointWithCov>().swap(temp_points);
                }
            } else {
                ct.n_full++;
                merge();
            }
        } else {
            if (update_enable) {
                addToPlane(pv);
                temp_points.push_back(pv);
                ct.n_ins++;
                newly_add_point++;
                if (newly_add_point >= update_point_thresh) {
                    updatePlane();
                    newly_add_point = 0;
                }
                if (temp_points.size() >= (size_t)max_point_thresh) {
                    update_enable = false;
                    std::vector<PointWithCov>().swap(temp_points);
                }
            } else {
                ct.n_full++;                                        // Q12: inert
            }
        }
    }
}

// voxel_map.cpp:97-136
void VoxelGrid::updatePlane() {
    assert(temp_points.size() == (size_t)plane->n);                 // Q20: live in the shipped build
    if (plane->n < update_point_thresh) return;
    is_init = true;
    const double nd = static_cast<double>(plane->n);
    M3 cov = sub(divs(plane->ppt, nd), outer(plane->mean, plane->mean));
    double evals[3];
    M3 evecs;
    eig3_sym(cov, evals, evecs);
    map->counters.n_refit++;
    if (map->track_margins) { const double mg = std::abs(evals[0] - plane_thresh); if (mg < map->margins.plane) map->margins.plane = mg; }
    if (evals[0] > plane_thresh) {
        is_plane = false;                                           // Q13: old norm / cov stay
        return;
    }
    is_plane = true;
    const M3 J_Q = divs(M3::identity(), nd);
    V3 plane_norm = v3(evecs(0, 0), evecs(1, 0), evecs(2, 0));
    map->counters.refit_points += (int64_t)temp_points.size();
    for (PointWithCov& pv : temp_points) {
        Mat<6, 3> J;
        M3 F = M3::zero();
        for (int m = 1; m < 3; m++) {
            const V3 vm = v3(evecs(0, m), evecs(1, m), evecs(2, m));
            // (p - mean)^T / (n (l0 - lm)) * (v_m n^T + n v_m^T)
            const Mat<1, 3> lhs = divs(tr(sub(pv.point, plane->mean)), (plane->n) * (evals[0] - evals[m]));
            const M3 S = add(outer(vm, plane_norm), outer(plane_norm, vm));
            const Mat<1, 3> Fm = mul(lhs, S);
            F(m, 0) = Fm[0]; F(m, 1) = Fm[1]; F(m, 2) = Fm[2];
        }
        set_block(J, 0, 0, mul(evecs, F));
        set_block(J, 3, 0, J_Q);
        plane->cov = add(plane->cov, mul(mul(J, pv.cov), tr(J)));   // Q7: accumulates across refits
    }
    const double axis_distance = -dot(plane->mean, plane_norm);
    if (axis_distance < 0.0) plane_norm = neg(plane_norm);
    plane->norm = plane_norm;
    center = plane->mean;
}

static inline double trace3(const M6& c, int o) { return c(o, o) + c(o + 1, o + 1) + c(o + 2, o + 2); }

// voxel_map.cpp:138-186
void VoxelGrid::merge() {
    map->counters.n_mergeprobe++;
    if (merge_epoch != map->epoch) { merge_epoch = map->epoch; map->counters.n_mergevox++; }   // distinct (full plane voxel, update) pairs: SURVEY 8(d) N_mergeprobe
    VoxelKey near[6] = {
        {position.x - 1, position.y, position.z}, {position.x, position.y - 1, position.z},
        {position.x, position.y, position.z - 1}, {position.x + 1, position.y, position.z},
        {position.x, position.y + 1, position.z}, {position.x, position.y, position.z + 1}};
    for (VoxelKey& k : near) {
        auto it = map->featmap.find(k);
        if (it == map->featmap.end()) continue;
        std::shared_ptr<VoxelGrid> nb = it->second;
        if (nb->group_id == group_id || nb->update_enable || !nb->is_plane) continue;
        const double norm_distance = 1.0 - dot(nb->plane->norm, plane->norm);
        const double axis_distance = std::abs(dot(nb->plane->norm, nb->plane->mean) - dot(plane->norm, plane->mean));
        if (map->track_margins) {
            const double ma = std::abs(norm_distance - map->merge_thresh_for_angle), md = std::abs(axis_distance - map->merge_thresh_for_distance);
            if (ma < map->margins.merge_angle) map->margins.merge_angle = ma;
            if (md < map->margins.merge_dist) map->margins.merge_dist = md;
        }
        if (norm_distance > map->merge_thresh_for_angle || axis_distance > map->merge_thresh_for_distance) continue;
        const double tn0 = trace3(plane->cov, 0), tm0 = trace3(plane->cov, 3);
        const double tn1 = trace3(nb->plane->cov, 0), tm1 = trace3(nb->plane->cov, 3);
        const double tc0 = tn0 + tm0, tc1 = tn1 + tm1;
        // Q9: precedence exactly as written in the reference (voxel_map.cpp:166-167)
        V3 new_mean = add(scale(nb->plane->mean, tm0), divs(scale(plane->mean, tm1), tm0 + tm1));
        V3 new_norm = add(scale(nb->plane->norm, tn0), divs(scale(plane->norm, tn1), tn0 + tn1));
        M6 new_cov = divs(add(scale(nb->plane->cov, tc0 * tc0), scale(plane->cov, tc1 * tc1)), (tc0 + tc1) * (tc0 + tc1));
        nb->group_id = group_id;
        merged = true;
        nb->merged = true;
        if (-dot(new_mean, new_norm) < 0.0) new_norm = neg(new_norm);
        plane->mean = new_mean; plane->norm = new_norm; plane->cov = new_cov;
        nb->plane->mean = new_mean; nb->plane->norm = new_norm; nb->plane->cov = new_cov;
        map->counters.n_merge++;
    }
}

// voxel_map.cpp:188-192
VoxelMap::VoxelMap(int mpt, int upt, double pth, double vs, int cap)
    : max_point_thresh(mpt), update_point_thresh(upt), plane_thresh(pth), voxel_size(vs), capacity(cap) {}

// voxel_map.cpp:194-198
VoxelKey VoxelMap::index(const V3& p) const {
    VoxelKey k;
    k.x = static_cast<int64_t>(std::floor(p[0] / voxel_size));
    k.y = static_cast<int64_t>(std::floor(p[1] / voxel_size));
    k.z = static_cast<int64_t>(std::floor(p[2] / voxel_size));
    return k;
}

// voxel_map.cpp:200-230
void VoxelMap::build(std::vector<PointWithCov>& pvs) {
    evicted.clear();
    counters = MapCounters();
    epoch++;
    counters.n_points = (int64_t)pvs.size();
    for (PointWithCov& pv : pvs) {
        VoxelKey k = index(pv.point);
        auto it = featmap.find(k);
        if (it == featmap.end()) {
            featmap[k] = std::make_shared<VoxelGrid>(max_point_thresh, update_point_thresh, plane_thresh, k, this);
            cache.push_front(k);
            featmap[k]->cache_it = cache.begin();
            counters.n_created++;
            if (cache.size() > (size_t)capacity) {
                evicted.push_back(cache.back());
                counters.n_evicted++;
                featmap.erase(cache.back());
                cache.pop_back();
            }
        } else {
            cache.splice(cache.begin(), cache, featmap[k]->cache_it);
        }
        std::shared_ptr<VoxelGrid>& vg = featmap[k];
        if (vg->touch_epoch != epoch) { vg->touch_epoch = epoch; counters.n_touch++; }
        vg->addPoint(pv);
        counters.n_ins++;
    }
    for (auto it = featmap.begin(); it != featmap.end(); it++) it->second->updatePlane();
}

// voxel_map.cpp:232-256
void VoxelMap::update(std::vector<PointWithCov>& pvs) {
    evicted.clear();
    counters = MapCounters();
    epoch++;
    counters.n_points = (int64_t)pvs.size();
    for (PointWithCov& pv : pvs) {
        VoxelKey k = index(pv.point);
        auto it = featmap.find(k);
        if (it == featmap.end()) {
            featmap[k] = std::make_shared<VoxelGrid>(max_point_thresh, update_point_thresh, plane_thresh, k, this);
            cache.push_front(k);
            featmap[k]->cache_it = cache.begin();
            counters.n_created++;
            if (cache.size() > (size_t)capacity) {
                evicted.push_back(cache.back());
                counters.n_evicted++;
                featmap.erase(cache.back());
                cache.pop_back();
            }
        } else {
            cache.splice(cache.begin(), cache, featmap[k]->cache_it);
        }
        std::shared_ptr<VoxelGrid>& vg = featmap[k];
        if (vg->touch_epoch != epoch) { vg->touch_epoch = epoch; counters.n_touch++; }
        vg->pushPoint(pv);
    }
}

// voxel_map.cpp:258-276
bool VoxelMap::buildResidual(ResidualData& data, std::shared_ptr<VoxelGrid> vg) {
    data.is_valid = false;
    if (vg->is_plane) {
        V3 p2m = sub(data.point_world, vg->plane->mean);
        data.plane_norm = vg->plane->norm;
        data.plane_mean = vg->plane->mean;
        data.residual = dot(data.plane_norm, p2m);
        Mat<1, 6> J_nq;
        J_nq[0] = p2m[0]; J_nq[1] = p2m[1]; J_nq[2] = p2m[2];
        J_nq[3] = -data.plane_norm[0]; J_nq[4] = -data.plane_norm[1]; J_nq[5] = -data.plane_norm[2];
        double sigma_l = mul(mul(J_nq, data.plane_cov), tr(J_nq))[0];                 // == 0 (Q1)
        sigma_l += mul(mul(tr(data.plane_norm), data.cov_world), data.plane_norm)[0];
        if (std::abs(data.residual) < 3.0 * std::sqrt(sigma_l)) data.is_valid = true;
        if (track_margins) {          // instrumentation, off by default (the timed baselines run without it)
            const double mg = std::abs(std::abs(data.residual) - 3.0 * std::sqrt(sigma_l));
#ifdef _OPENMP
#pragma omp critical(orc_margin)
#endif
            if (mg < margins.gate) margins.gate = mg;
        }
    }
    return data.is_valid;
}

// =============================================================== lio_builder.cpp

// lio_builder.cpp:5-26
void LIOBuilder::loadConfig(const LIOConfig& cfg) {
    config = cfg;
    status = IMU_INIT;
    data_group.Q = M12::identity();
    for (int i = 0; i < 3; i++) {
        data_group.Q(i, i) = config.ng;
        data_group.Q(3 + i, 3 + i) = config.na;
        data_group.Q(6 + i, 6 + i) = config.nbg;
        data_group.Q(9 + i, 9 + i) = config.nba;
    }
    map = std::make_shared<VoxelMap>(config.max_point_thresh, config.update_size_thresh, config.plane_thresh,
                                     config.voxel_size, config.map_capacity);
    map->merge_thresh_for_angle = config.merge_thresh_for_angle;
    map->merge_thresh_for_distance = config.merge_thresh_for_distance;
    lidar_cloud.clear();
    kf.set_share_function([this](State& s, SharedState& d) { sharedUpdateFunc(s, d); });
    // deviation, stated: the reference never calls setMaxIter (Q4) so it always runs <=5;
    // opti_max_iter defaults to 5, which reproduces that.
    kf.setMaxIter(config.opti_max_iter);
    data_group.residual_info.assign((size_t)config.max_points_per_scan, ResidualData());   // reference: resize(10000), Q3
}

// lio_builder.cpp:28-63
bool LIOBuilder::initializeImu(std::vector<IMUData>& imus) {
    data_group.imu_cache.insert(data_group.imu_cache.end(), imus.begin(), imus.end());
    if (data_group.imu_cache.size() < (size_t)config.imu_init_num) return false;
    V3 acc_mean = V3::zero(), gyro_mean = V3::zero();
    for (const auto& imu : data_group.imu_cache) {
        acc_mean = add(acc_mean, imu.acc);
        gyro_mean = add(gyro_mean, imu.gyro);
    }
    acc_mean = divs(acc_mean, static_cast<double>(data_group.imu_cache.size()));
    gyro_mean = divs(gyro_mean, static_cast<double>(data_group.imu_cache.size()));
    data_group.gravity_norm = norm(acc_mean);
    kf.x().rot_ext = config.r_il;
    kf.x().pos_ext = config.p_il;
    kf.x().bg = gyro_mean;
    if (config.gravity_align) {
        kf.x().rot = rot_from_two_vectors(normalized(neg(acc_mean)), v3(0.0, 0.0, -1.0));
        kf.x().initG(v3(0, 0, -1.0));
    } else {
        kf.x().initG(neg(acc_mean));
    }
    M23& P = kf.P();
    P = M23::identity();
    for (int i = 0; i < 3; i++) {
        P(6 + i, 6 + i) = 0.00001;
        P(9 + i, 9 + i) = 0.00001;
        P(15 + i, 15 + i) = 0.0001;
        P(18 + i, 18 + i) = 0.0001;
    }
    P(21, 21) = 0.00001;
    P(22, 22) = 0.00001;
    data_group.last_imu = imus.back();
    return true;
}

// lio_builder.cpp:65-153
void LIOBuilder::undistortCloud(SyncPackage& package) {
    data_group.imu_cache.clear();
    data_group.imu_cache.push_back(data_group.last_imu);
    data_group.imu_cache.insert(data_group.imu_cache.end(), package.imus.begin(), package.imus.end());

    const double imu_time_end = data_group.imu_cache.back().timestamp;
    const double cloud_time_begin = package.cloud_start_time;
    const double cloud_time_end = package.cloud_end_time;
    // std::sort is not stable; ties in curvature are broken by stable_sort here (the
    // synthetic driver never produces ties)
    std::stable_sort(package.cloud.begin(), package.cloud.end(),
                     [](const CloudPoint& a, const CloudPoint& b) { return a.curvature < b.curvature; });

    data_group.imu_poses_cache.clear();
    data_group.imu_poses_cache.push_back(Pose{0.0, data_group.last_acc, data_group.last_gyro, kf.x().vel, kf.x().pos, kf.x().rot});

    V3 acc_val = V3::zero(), gyro_val = V3::zero();
    double dt = 0.0;
    Input inp; inp.acc = V3::zero(); inp.gyro = V3::zero();

    for (size_t i = 0; i + 1 < data_group.imu_cache.size(); i++) {
        IMUData& head = data_group.imu_cache[i];
        IMUData& tail = data_group.imu_cache[i + 1];
        if (tail.timestamp < data_group.last_cloud_end_time) continue;
        gyro_val = scale(add(head.gyro, tail.gyro), 0.5);
        acc_val = scale(add(head.acc, tail.acc), 0.5);
        acc_val = divs(scale(acc_val, 9.81), data_group.gravity_norm);
        if (head.timestamp < data_group.last_cloud_end_time) dt = tail.timestamp - data_group.last_cloud_end_time;
        else dt = tail.timestamp - head.timestamp;
        inp.acc = acc_val;
        inp.gyro = gyro_val;
        kf.predict(inp, dt, data_group.Q);
        data_group.last_gyro = sub(gyro_val, kf.x().bg);
        data_group.last_acc = add(mul(kf.x().rot, sub(acc_val, kf.x().ba)), kf.x().g);
        const double offset = tail.timestamp - cloud_time_begin;
        data_group.imu_poses_cache.push_back(Pose{offset, data_group.last_acc, data_group.last_gyro, kf.x().vel, kf.x().pos, kf.x().rot});
    }
    dt = cloud_time_end - imu_time_end;
    kf.predict(inp, dt, data_group.Q);

    data_group.last_imu = package.imus.back();
    data_group.last_cloud_end_time = cloud_time_end;

    const M3 cur_rot = kf.x().rot;
    const V3 cur_pos = kf.x().pos;
    const M3 cur_rot_ext = kf.x().rot_ext;
    const V3 cur_pos_ext = kf.x().pos_ext;

    if (package.cloud.empty()) return;
    std::vector<CloudPoint>& pts = package.cloud;
    size_t ip = pts.size() - 1;
    for (size_t kp = data_group.imu_poses_cache.size() - 1; kp != 0; kp--) {
        const Pose& head = data_group.imu_poses_cache[kp - 1];
        const Pose& tail = data_group.imu_poses_cache[kp];
        const M3 imu_rot = head.rot;
        const V3 imu_pos = head.pos, imu_vel = head.vel, imu_acc = tail.acc, imu_gyro = tail.gyro;
        for (; pts[ip].curvature / double(1000) > head.offset; ip--) {
            dt = pts[ip].curvature / double(1000) - head.offset;
            const V3 point = v3(pts[ip].x, pts[ip].y, pts[ip].z);
            const M3 point_rot = mul(imu_rot, so3_exp(scale(imu_gyro, dt)));
            const V3 point_pos = add(add(imu_pos, scale(imu_vel, dt)), scale(scale(scale(imu_acc, 0.5), dt), dt));
            const V3 inner = sub(add(mul(point_rot, add(mul(cur_rot_ext, point), cur_pos_ext)), point_pos), cur_pos);
            const V3 pc = mul(tr(cur_rot_ext), sub(mul(tr(cur_rot), inner), cur_pos_ext));
            pts[ip].x = (float)pc[0];
            pts[ip].y = (float)pc[1];
            pts[ip].z = (float)pc[2];
            if (ip == 0) break;
        }
    }
}

// lio_builder.cpp:155-163.  pcl::transformPointCloud(Matrix4f): float32, association of
// PCL >= 1.10's SSE Transformer::se3: x' = m00 x + (m01 y + (m02 z + tx)), separate mul/add.
std::vector<CloudPoint> LIOBuilder::lidarToWorld(const std::vector<CloudPoint>& cloud) {
    const M3 Rd = mul(kf.x().rot, kf.x().rot_ext);
    const V3 td = add(mul(kf.x().rot, kf.x().pos_ext), kf.x().pos);
    float m[3][3], t[3];
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) m[i][j] = (float)Rd(i, j); t[i] = (float)td[i]; }
    std::vector<CloudPoint> out(cloud.size());
    for (size_t i = 0; i < cloud.size(); i++) {
        const float x = cloud[i].x, y = cloud[i].y, z = cloud[i].z;
        float o[3];
        for (int r = 0; r < 3; r++) {
            const float p0 = m[r][0] * x, p1 = m[r][1] * y, p2 = m[r][2] * z;
            const float s2 = p2 + t[r];
            const float s1 = p1 + s2;
            o[r] = p0 + s1;
        }
        out[i] = cloud[i];
        out[i].x = o[0]; out[i].y = o[1]; out[i].z = o[2];
    }
    return out;
}

// pcl::VoxelGrid<PointT>::applyFilter (PCL 1.10 filters/impl/voxel_grid.hpp, restated from its published algorithm; PCL is
// not in this image and the reference does not pin its version: "parity unpinned").  Call site: lio_builder.cpp:215-219.
//   * bounding box of the finite points (getMinMax3D), float32
//   * if the leaf grid would overflow int32 PCL warns and returns the input unchanged
//   * leaf index ijk = floor(p * inverse_leaf) - min_b per axis, linear idx = i + j*dx + k*dx*dy
//   * std::sort by idx; the order of equal keys is unspecified in PCL -> this restatement keeps the original point
//     order inside a leaf (a valid outcome of std::sort and the one the device reproduces)
//   * one output point per leaf in ascending idx order = CentroidPoint: float32 sums in that order, divided by n
//     (AccumulatorXYZ / AccumulatorCurvature; intensity and normal are not read downstream and not carried here)
std::vector<CloudPoint> LIOBuilder::voxelGridFilter(const std::vector<CloudPoint>& cloud, float leaf) {
    const float inv = 1.0f / leaf;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    size_t n_finite = 0;
    for (const CloudPoint& p : cloud) {
        if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
        const float v[3] = {p.x, p.y, p.z};
        for (int a = 0; a < 3; a++) { if (v[a] < mn[a]) mn[a] = v[a]; if (v[a] > mx[a]) mx[a] = v[a]; }
        n_finite++;
    }
    std::vector<CloudPoint> out;
    if (n_finite == 0) return out;
    int64_t d[3];
    for (int a = 0; a < 3; a++) d[a] = (int64_t)((mx[a] - mn[a]) * inv) + 1;
    if (d[0] * d[1] * d[2] > (int64_t)INT32_MAX) return cloud;       // "Leaf size is too small": output = input
    int min_b[3], max_b[3], div_b[3];
    for (int a = 0; a < 3; a++) {
        min_b[a] = (int)std::floor(mn[a] * inv);
        max_b[a] = (int)std::floor(mx[a] * inv);
        div_b[a] = max_b[a] - min_b[a] + 1;
    }
    const int mul1 = div_b[0], mul2 = div_b[0] * div_b[1];
    std::vector<std::pair<unsigned, unsigned>> iv;                   // (idx, cloud_point_index)
    iv.reserve(cloud.size());
    for (size_t i = 0; i < cloud.size(); i++) {
        const CloudPoint& p = cloud[i];
        if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
        const int i0 = (int)(std::floor(p.x * inv) - (float)min_b[0]);
        const int i1 = (int)(std::floor(p.y * inv) - (float)min_b[1]);
        const int i2 = (int)(std::floor(p.z * inv) - (float)min_b[2]);
        iv.emplace_back((unsigned)(i0 + i1 * mul1 + i2 * mul2), (unsigned)i);
    }
    std::stable_sort(iv.begin(), iv.end(), [](const std::pair<unsigned, unsigned>& a, const std::pair<unsigned, unsigned>& b) { return a.first < b.first; });
    for (size_t first = 0; first < iv.size();) {
        size_t last = first + 1;
        while (last < iv.size() && iv[last].first == iv[first].first) last++;
        float sx = 0.0f, sy = 0.0f, sz = 0.0f, sc = 0.0f;
        for (size_t l = first; l < last; l++) {
            const CloudPoint& p = cloud[iv[l].second];
            sx += p.x; sy += p.y; sz += p.z; sc += p.curvature;
        }
        const float cnt = (float)(last - first);
        out.push_back(CloudPoint{sx / cnt, sy / cnt, sz / cnt, sc / cnt});
        first = last;
    }
    return out;
}

// MAP_INIT body, lio_builder.cpp:188-208
void LIOBuilder::firstScan(const std::vector<CloudPoint>& cloud) {
    prior_x = kf.x();
    prior_P = kf.P();
    std::vector<CloudPoint> point_world = lidarToWorld(cloud);
    std::vector<PointWithCov> pv_list;
    pv_list.reserve(cloud.size());
    const M3 r_wl = mul(kf.x().rot, kf.x().rot_ext);
    const M3 Prr = get_block<3, 3>(kf.P(), 3, 3), Ppp = get_block<3, 3>(kf.P(), 0, 0);
    for (size_t i = 0; i < point_world.size(); i++) {
        PointWithCov pv;
        pv.point = v3(point_world[i].x, point_world[i].y, point_world[i].z);
        V3 point_body = v3(cloud[i].x, cloud[i].y, cloud[i].z);
        M3 point_cov;
        calcBodyCov(point_body, config.ranging_cov, config.angle_cov, point_cov);
        const M3 cm = hat(point_body);
        pv.cov = add(add(mul(mul(r_wl, point_cov), tr(r_wl)), mul(mul(cm, Prr), tr(cm))), Ppp);
        pv_list.push_back(pv);
    }
    map->build(pv_list);
    last_pv_list.swap(pv_list);
}

// lio_builder.cpp:224-229
void LIOBuilder::setScan(const std::vector<CloudPoint>& cloud) {
    lidar_cloud = cloud;
    const int size = (int)lidar_cloud.size();
    if ((size_t)size > data_group.residual_info.size()) {
        std::fprintf(stderr, "oracle: scan of %d points exceeds max_points_per_scan=%zu (reference: UB, Q3)\n", size,
                     data_group.residual_info.size());
        std::abort();
    }
    for (int i = 0; i < size; i++) {
        ResidualData& r = data_group.residual_info[i];
        r.point_lidar = v3(lidar_cloud[i].x, lidar_cloud[i].y, lidar_cloud[i].z);
        calcBodyCov(r.point_lidar, config.ranging_cov, config.angle_cov, r.cov_lidar);
    }
}

// lio_builder.cpp:230-246
void LIOBuilder::hotPath() {
    const int size = (int)lidar_cloud.size();
    effect_nums.clear();
    iter_states.clear();
    iter_H.clear();
    iter_b.clear();
    prior_x = kf.x();
    prior_P = kf.P();
    kf.update();
    std::vector<CloudPoint> point_world = lidarToWorld(lidar_cloud);
    std::vector<PointWithCov> pv_list;
    const M3 r_wl = mul(kf.x().rot, kf.x().rot_ext);
    const M3 Prr = get_block<3, 3>(kf.P(), 3, 3), Ppp = get_block<3, 3>(kf.P(), 0, 0);
    for (int i = 0; i < size; i++) {
        PointWithCov pv;
        pv.point = v3(point_world[i].x, point_world[i].y, point_world[i].z);
        const M3 cov = data_group.residual_info[i].cov_lidar;
        const M3 cm = hat(data_group.residual_info[i].point_lidar);
        pv.cov = add(add(mul(mul(r_wl, cov), tr(r_wl)), mul(mul(cm, Prr), tr(cm))), Ppp);
        pv_list.push_back(pv);
    }
    map->update(pv_list);
    last_pv_list.swap(pv_list);
}

// lio_builder.cpp:175-248
void LIOBuilder::process(SyncPackage& package) {
    if (status == IMU_INIT) {
        if (initializeImu(package.imus)) {
            status = MAP_INIT;
            data_group.last_cloud_end_time = package.cloud_end_time;
        }
    } else if (status == MAP_INIT) {
        undistortCloud(package);
        firstScan(package.cloud);
        status = LIO_MAPPING;
    } else {
        undistortCloud(package);
        if (config.scan_resolution > 0.0) {
            // scan_filter.setLeafSize(r, r, r) takes floats (lio_builder.cpp:13-14)
            setScan(voxelGridFilter(package.cloud, (float)config.scan_resolution));
        } else {
            setScan(package.cloud);                     // pcl::copyPointCloud
        }
        hotPath();
    }
}

// lio_builder.cpp:250-311
void LIOBuilder::sharedUpdateFunc(State& state, SharedState& shared) {
    iter_states.push_back(state);
    const M3 r_wl = mul(state.rot, state.rot_ext);
    const V3 p_wl = add(mul(state.rot, state.pos_ext), state.pos);
    const int size = (int)lidar_cloud.size();
    const M3 Prr = get_block<3, 3>(kf.P(), 3, 3), Ppp = get_block<3, 3>(kf.P(), 0, 0);
    std::vector<ResidualData>& ri = data_group.residual_info;

#ifdef _OPENMP
#pragma omp parallel for num_threads(omp_threads)
#endif
    for (int i = 0; i < size; i++) {
        ri[i].point_world = add(mul(r_wl, ri[i].point_lidar), p_wl);
        const M3 cm = hat(ri[i].point_lidar);
        ri[i].cov_world = add(add(mul(mul(r_wl, ri[i].cov_lidar), tr(r_wl)), mul(mul(cm, Prr), tr(cm))), Ppp);
        VoxelKey position = map->index(ri[i].point_world);
        ri[i].key = position;
        auto iter = map->featmap.find(position);
        if (iter != map->featmap.end()) {
            map->buildResidual(ri[i], iter->second);
            ri[i].status = (uint8_t)(1 | (iter->second->is_plane ? 2 : 0) | (ri[i].is_valid ? 4 : 0));
        } else {
            ri[i].status = (uint8_t)(ri[i].is_valid ? 4 : 0);          // Q2: stale record survives
        }
    }

    shared.H = M12::zero();
    shared.b = V12::zero();
    int effect_num = 0;
    for (int i = 0; i < size; i++) {
        if (!ri[i].is_valid) continue;
        effect_num++;
        Mat<1, 12> J = Mat<1, 12>::zero();
        const V3 plane_norm = ri[i].plane_norm;
        Mat<1, 6> Jn;
        const V3 d = sub(ri[i].point_world, ri[i].plane_mean);
        Jn[0] = d[0]; Jn[1] = d[1]; Jn[2] = d[2];
        Jn[3] = -plane_norm[0]; Jn[4] = -plane_norm[1]; Jn[5] = -plane_norm[2];
        double r_cov = mul(mul(Jn, ri[i].plane_cov), tr(Jn))[0];                       // == 0 (Q1)
        const Mat<1, 3> nt = tr(plane_norm);
        r_cov += mul(mul(mul(mul(nt, r_wl), ri[i].cov_lidar), tr(r_wl)), plane_norm)[0];
        const double r_info = r_cov < 0.0002 ? 5000 : 1.0 / r_cov;                     // Q19
        J[0] = plane_norm[0]; J[1] = plane_norm[1]; J[2] = plane_norm[2];
        const Mat<1, 3> jr = mul(mul(neg(nt), state.rot), hat(add(mul(state.rot_ext, ri[i].point_lidar), state.pos_ext)));
        J[3] = jr[0]; J[4] = jr[1]; J[5] = jr[2];
        if (config.estimate_ext) {
            const Mat<1, 3> je = mul(mul(neg(nt), r_wl), hat(ri[i].point_lidar));
            J[6] = je[0]; J[7] = je[1]; J[8] = je[2];
            const Mat<1, 3> jp = mul(nt, state.rot);
            J[9] = jp[0]; J[10] = jp[1]; J[11] = jp[2];
        }
        // H += J^T * r_info * J ; b += J^T * r_info * residual
        for (int a = 0; a < 12; a++) {
            const double ja = J[a] * r_info;
            for (int c = 0; c < 12; c++) shared.H(a, c) += ja * J[c];
            shared.b[a] += ja * ri[i].residual;
        }
    }
    effect_nums.push_back(effect_num);
    iter_H.push_back(shared.H);
    iter_b.push_back(shared.b);
    if (effect_num < 1) std::fprintf(stderr, "NO EFFECTIVE POINT\n");
}

}  // namespace orc
