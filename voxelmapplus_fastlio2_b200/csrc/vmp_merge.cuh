// vmp_merge.cuh — VoxelGrid::merge() (voxel_map.cpp:138-186) on the device.  Included by
// vmp_map.cu inside namespace vmp.
//
// merge() runs for EVERY point that lands in a full plane voxel (Q11) and is a no-op almost
// always.  Phase 1 (k_merge_prefilter, parallel, 8 lanes per voxel): which of the voxels that
// received such points could merge with some neighbour at all, given the current planes and
// groups, ignoring when during the scan the neighbour becomes eligible.  Phase 2
// (k_merge_serial, one warp): an event simulation in point order over those voxels only.
// Per event the six neighbour look-ups run on six lanes; the accept/reject decisions are taken
// strictly in the reference's order (-x -y -z +x +y +z, own plane updated in between, Q10); the
// 6x6 covariance blend is spread over the lanes.  Exactness argument:
//   * an event is a no-op unless some neighbour pair passes the thresholds on the CURRENT planes
//     and groups while that neighbour is alive, full and a plane AT THAT TIME;
//   * planes / groups change only through a successful merge; after one, the pairs (Y, X) of every
//     voxel Y adjacent to a changed voxel X are re-examined, and the changed voxels themselves are
//     re-examined against all their neighbours;
//   * a voxel whose only passing pairs become eligible later in the scan (neighbour not yet created
//     / not yet full) sleeps until the earliest such time instead of replaying every point.
#pragma once

__device__ __forceinline__ unsigned long long nbr_key(unsigned long long pk, int d, bool& ok) {
    long long x, y, z;
    unpack_key(pk, x, y, z);
    switch (d) {
        case 0: x -= 1; break; case 1: y -= 1; break; case 2: z -= 1; break;
        case 3: x += 1; break; case 4: y += 1; break; default: z += 1; break;
    }
    ok = key_in_range(x) && key_in_range(y) && key_in_range(z);
    return ok ? pack_key(x, y, z) : KEY_EMPTY;
}

// thresholds of voxel_map.cpp:156-160 on explicit plane parameters
__device__ __forceinline__ bool plane_thresholds(const DevMap& m, const V3& mA, const V3& nA, const V3& mB, const V3& nB) {
    const double norm_distance = 1.0 - dot(nB, nA);
    const double axis_distance = fabs(dot(nB, mB) - dot(nA, mA));
    return !(norm_distance > m.th_angle || axis_distance > m.th_dist);
}
__device__ __forceinline__ void load_plane(const DevMap& m, int s, V3& mean, V3& nrm) {
    const double* h = m.hot + (size_t)s * 8;
    mean = v3(h[0], h[1], h[2]);
    nrm = v3(h[3], h[4], h[5]);
}

// could merge(A) succeed against the (final state of) incarnation B at some time of this scan?
__device__ __forceinline__ bool pair_static(const DevMap& m, int A, int B) {
    uint32_t fb; int nb;
    hot_get_fn(m.hot, B, fb, nb);
    if ((fb & F_UE) || !(fb & F_PLANE)) return false;
    if (m.sgroup[B] == m.sgroup[A]) return false;
    V3 mA, nA, mB, nB;
    load_plane(m, A, mA, nA);
    load_plane(m, B, mB, nB);
    return plane_thresholds(m, mA, nA, mB, nB);
}
// window (start, end) of point indices t of this scan at which B is alive and full: start < t < end
__device__ __forceinline__ void pair_window(const DevMap& m, int B, unsigned scan_id, int& start, int& end) {
    const int born = (m.born_scan[B] == scan_id) ? m.ft[B] : -1;
    const int full = (m.full_scan[B] == scan_id) ? m.full_idx[B] : -1;
    start = born > full ? born : full;
    end = m.evict_t[B];
}

// one neighbour direction of "could merge(A) succeed after time `after`": returns the earliest time
// bound (events with t > bound may succeed), T_INF if never.  after = -1 gives the plain static test.
__device__ int wake_dir(const DevMap& m, int A, int d, int after, unsigned scan_id) {
    bool ok;
    const unsigned long long nk = nbr_key(m.skey[A], d, ok);
    if (!ok) return T_INF;
    int best = T_INF;
    for (int B = hash_find(m, nk); B >= 0; B = m.ghost[B]) {
        if (!pair_static(m, A, B)) continue;
        int s, e;
        pair_window(m, B, scan_id, s, e);
        const int bound = s > after ? s : after;                 // first usable events are those with t > bound
        if (bound + 1 < e && bound < best) best = bound;
    }
    return best;
}
__device__ __forceinline__ int wake_warp(const DevMap& m, int A, int after, unsigned scan_id) {
    const int lane = threadIdx.x & 31;
    int w = lane < 6 ? wake_dir(m, A, lane, after, scan_id) : T_INF;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) { const int y = __shfl_xor_sync(0xffffffffu, w, o); w = y < w ? y : w; }
    return __shfl_sync(0xffffffffu, w, 0);
}

__device__ __forceinline__ int event_floor(const DevMap& m, int A, unsigned scan_id) {
    return (m.full_scan[A] == scan_id) ? m.full_idx[A] : -1;     // merge() runs only for points after the closing one
}

// 8 lanes per touched voxel, lanes 0..5 take one neighbour each; voxels whose merge() can succeed at
// some time of this scan go into the active set of k_merge_serial with their first relevant event
__global__ void __launch_bounds__(128) k_merge_prefilter(DevMap m, DevCtl* ctl) {
    const int V = ctl->n_touched;
    const unsigned scan_id = ctl->scan_id;
    const int sub = threadIdx.x & 7;
    const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~7);
    const int ngroups = (gridDim.x * blockDim.x) >> 3;
    const int vend = (V + ngroups - 1) / ngroups * ngroups;          // keep the 8-lane groups converged
    for (int vi = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; vi < vend; vi += ngroups) {
        int w = T_INF, A = -1;
        if (vi < V) {
            A = m.touched[vi];
            if (m.evn[A] > 0 && sub < 6) w = wake_dir(m, A, sub, event_floor(m, A, scan_id), scan_id);
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) { const int y = __shfl_xor_sync(gmask, w, o); w = y < w ? y : w; }
        if (w == T_INF) continue;                                   // uniform within the 8-lane group
        // first merge() call of A after the earliest time one of its pairs can pass
        const int c = m.cnt[A], off = m.seg_off[A];
        int best = T_INF;
        for (int q = sub; q < c; q += 8) { const int i = m.seg[off + q]; if (i > w && i < best) best = i; }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) { const int y = __shfl_xor_sync(gmask, best, o); best = y < best ? y : best; }
        if (best != T_INF && sub == 0) {
            const int k = atomicAdd(&ctl->n_hot, 1);
            m.act_slot[k] = A;
            m.act_t[k] = best;
        }
    }
}

// first point index of voxel A in this scan that is > after (warp-parallel), T_INF if none
__device__ int next_event_warp(const DevMap& m, int A, int after) {
    const int lane = threadIdx.x & 31;
    const int c = m.cnt[A], off = m.seg_off[A];
    int best = T_INF;
    for (int q = lane; q < c; q += 32) { const int i = m.seg[off + q]; if (i > after && i < best) best = i; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const int y = __shfl_xor_sync(0xffffffffu, best, o); best = y < best ? y : best; }
    return best;
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// VoxelGrid::merge() of voxel A at time t, executed by the whole warp.  Returns the number of
// successful pair merges; changed[0..n) (warp-uniform) are the neighbour slots that were modified.
__device__ int merge_at_warp(const DevMap& m, DevCtl* ctl, int A, int t, unsigned scan_id, int* changed) {
    const int lane = threadIdx.x & 31;
    // ---- parallel part: lane d < 6 resolves neighbour d as of time t and loads its record
    int B = -1;
    bool elig = false;
    unsigned long long gB = 0;
    V3 mB = v3(0, 0, 0), nB = v3(0, 0, 0);
    uint32_t fB = 0; int cntB = 0;
    if (lane < 6) {
        bool ok;
        const unsigned long long nk = nbr_key(m.skey[A], lane, ok);
        if (ok) {
            for (int X = hash_find(m, nk); X >= 0; X = m.ghost[X]) {          // incarnation alive at time t
                const int born = (m.born_scan[X] == scan_id) ? m.ft[X] : -1;
                if (born <= t && t < m.evict_t[X]) { B = X; break; }
            }
        }
        if (B >= 0) {
            hot_get_fn(m.hot, B, fB, cntB);
            const bool closed = !(fB & F_UE) && (m.full_scan[B] != scan_id || m.full_idx[B] < t);
            elig = closed && (fB & F_PLANE);
            gB = m.sgroup[B];
            load_plane(m, B, mB, nB);
        }
    }
    // ---- ordered part (warp-uniform control flow)
    V3 mA, nA;
    load_plane(m, A, mA, nA);
    const unsigned long long gA = m.sgroup[A];
    double* ca = m.cov + (size_t)A * 36;
    int nchg = 0;
    for (int d = 0; d < 6; d++) {
        const int Bd = __shfl_sync(0xffffffffu, B, d);
        const int el = __shfl_sync(0xffffffffu, elig ? 1 : 0, d);
        const unsigned long long g = __shfl_sync(0xffffffffu, gB, d);
        if (Bd < 0 || !el || g == gA) continue;
        V3 mb, nb;
#pragma unroll
        for (int k = 0; k < 3; k++) { mb[k] = shfl_d(mB[k], d); nb[k] = shfl_d(nB[k], d); }
        if (!plane_thresholds(m, mA, nA, mb, nb)) continue;
        // lanes hold the 36 covariance entries of both voxels (lane e: entry e, lanes 0..3 also entry 32+e)
        double* cb = m.cov + (size_t)Bd * 36;
        const double a0 = ca[lane], b0 = cb[lane];
        const double a1 = lane < 4 ? ca[32 + lane] : 0.0, b1 = lane < 4 ? cb[32 + lane] : 0.0;
        const double tn0 = shfl_d(a0, 0) + shfl_d(a0, 7) + shfl_d(a0, 14);
        const double tm0 = shfl_d(a0, 21) + shfl_d(a0, 28) + shfl_d(a1, 3);
        const double tn1 = shfl_d(b0, 0) + shfl_d(b0, 7) + shfl_d(b0, 14);
        const double tm1 = shfl_d(b0, 21) + shfl_d(b0, 28) + shfl_d(b1, 3);
        const double tc0 = tn0 + tm0, tc1 = tn1 + tm1;
        // Q9: operator precedence exactly as in voxel_map.cpp:166-167
        V3 nm, nn;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            nm[k] = mb[k] * tm0 + (mA[k] * tm1) / (tm0 + tm1);
            nn[k] = nb[k] * tn0 + (nA[k] * tn1) / (tn0 + tn1);
        }
        const double w0 = tc0 * tc0, w1 = tc1 * tc1, den = (tc0 + tc1) * (tc0 + tc1);
        const double c0 = (b0 * w0 + a0 * w1) / den;
        ca[lane] = c0; cb[lane] = c0;
        if (lane < 4) { const double c1 = (b1 * w0 + a1 * w1) / den; ca[32 + lane] = c1; cb[32 + lane] = c1; }
        if (-dot(nm, nn) < 0.0) nn = neg(nn);
        mA = nm; nA = nn;                                         // own plane is updated before the next neighbour (Q10)
        if (lane < 3) {
            m.hot[(size_t)A * 8 + lane] = nm[lane];  m.hot[(size_t)A * 8 + 3 + lane] = nn[lane];
            m.hot[(size_t)Bd * 8 + lane] = nm[lane]; m.hot[(size_t)Bd * 8 + 3 + lane] = nn[lane];
        }
        if (lane == 0) {
            m.sgroup[Bd] = gA;
            uint32_t fa; int na;
            hot_get_fn(m.hot, A, fa, na);
            hot_set_fn(m.hot, A, fa | F_MERGED, na);
            ctl->st.n_merge += 1;
        }
        if (lane == d) hot_set_fn(m.hot, Bd, fB | F_MERGED, cntB);      // neighbour's MERGED flag, by the lane that loaded it
        __syncwarp();
        changed[nchg++] = Bd;
    }
    return nchg;
}

// Ordered simulation of the merge() calls that can have an effect (one warp).
__global__ void __launch_bounds__(32) k_merge_serial(DevMap m, DevCtl* ctl) {
    const int lane = threadIdx.x;
    int na = ctl->n_hot;                 // active set (voxel, first relevant event) prepared by k_merge_prefilter
    if (na == 0) return;
    if (lane == 0) ctl->dbg[0] = na;
    int n_events = 0, n_react = 0;
    const unsigned scan_id = ctl->scan_id;
    while (na > 0) {
        // earliest pending event
        int bt = T_INF, bk = -1;
        for (int k = lane; k < na; k += 32) { const int t = m.act_t[k]; if (t < bt) { bt = t; bk = k; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int yt = __shfl_xor_sync(0xffffffffu, bt, o), yk = __shfl_xor_sync(0xffffffffu, bk, o);
            if (yt < bt) { bt = yt; bk = yk; }
        }
        const int A = m.act_slot[bk], t = bt;
        int changed[6];
        const int nchg = merge_at_warp(m, ctl, A, t, scan_id, changed);
        n_events++;
        if (nchg > 0) {
            // planes / groups of A and changed[] moved.  (a) every voxel Y adjacent to a changed voxel X:
            // only its pair with X can have flipped; (b) a changed neighbour itself: all of its pairs.
            // One (X, direction) or one changed voxel per lane.
            const int nx = nchg + 1;
            const int ncand = nx * 6;
            // (b) first: the changed neighbours themselves, six directions on six lanes each
            for (int q = 0; q < nchg; q++) {
                const int Yb = changed[q];
                if (m.cnt[Yb] == 0 || m.evn[Yb] == 0) continue;
                const int fl = event_floor(m, Yb, scan_id);
                const int wb = wake_warp(m, Yb, t > fl ? t : fl, scan_id);
                if (wb == T_INF) continue;
                const int nt = next_event_warp(m, Yb, wb);
                if (nt == T_INF) continue;
                int found = -1;
                for (int k = lane; k < na; k += 32) if (m.act_slot[k] == Yb) found = k;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { const int y = __shfl_xor_sync(0xffffffffu, found, o); found = y > found ? y : found; }
                if (found >= 0) { if (lane == 0 && nt < m.act_t[found]) m.act_t[found] = nt; }
                else if (na >= m.nmax) { if (lane == 0) atomicOr(&ctl->err, E_QUEUE); }
                else { if (lane == 0) { m.act_slot[na] = Yb; m.act_t[na] = nt; } na++; }
                __syncwarp();
            }
            for (int c0 = 0; c0 < ncand; c0 += 32) {
                const int c = c0 + lane;
                int Y = -1, w = T_INF;
                if (c < nx * 6) {
                    const int X = (c / 6 == 0) ? A : changed[c / 6 - 1];
                    bool ok;
                    const unsigned long long nk = nbr_key(m.skey[X], c % 6, ok);
                    Y = ok ? hash_find(m, nk) : -1;
                    if (Y == A || (Y >= 0 && (m.cnt[Y] == 0 || m.evn[Y] == 0))) Y = -1;   // no merge() call of Y in this scan
                    for (int q = 0; q < nchg && Y >= 0; q++) if (changed[q] == Y) Y = -1;  // handled by (b)
                    if (Y >= 0 && pair_static(m, Y, X)) {
                        int s, e;
                        pair_window(m, X, scan_id, s, e);
                        const int fl = event_floor(m, Y, scan_id);
                        int bound = s > t ? s : t;
                        bound = fl > bound ? fl : bound;
                        if (bound + 1 < e) w = bound;
                    }
                    if (w == T_INF) Y = -1;
                }
                unsigned hotmask = __ballot_sync(0xffffffffu, Y >= 0);
                while (hotmask) {
                    const int src = __ffs(hotmask) - 1;
                    hotmask &= hotmask - 1;
                    const int Yh = __shfl_sync(0xffffffffu, Y, src);
                    const int wh = __shfl_sync(0xffffffffu, w, src);
                    const int nt = next_event_warp(m, Yh, wh);
                    if (nt == T_INF) continue;
                    int found = -1;
                    for (int k = lane; k < na; k += 32) if (m.act_slot[k] == Yh) found = k;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) { const int y = __shfl_xor_sync(0xffffffffu, found, o); found = y > found ? y : found; }
                    if (found >= 0) { if (lane == 0 && nt < m.act_t[found]) m.act_t[found] = nt; }
                    else if (na >= m.nmax) { if (lane == 0) atomicOr(&ctl->err, E_QUEUE); }
                    else { if (lane == 0) { m.act_slot[na] = Yh; m.act_t[na] = nt; } na++; n_react++; }
                    __syncwarp();
                }
            }
        }
        // advance A: sleep until the earliest time one of its pairs can pass again, drop it if none can
        int nt = T_INF;
        const int w = wake_warp(m, A, t, scan_id);
        if (w != T_INF) nt = next_event_warp(m, A, w);
        if (nt != T_INF) { if (lane == 0) m.act_t[bk] = nt; }
        else { if (lane == 0) { m.act_slot[bk] = m.act_slot[na - 1]; m.act_t[bk] = m.act_t[na - 1]; } na--; }
        __syncwarp();
    }
    if (lane == 0) { ctl->dbg[1] = n_events; ctl->dbg[2] = n_react; }
}
