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
            atomicAdd((unsigned long long*)&ctl->st.n_merge, 1ull);
        }
        if (lane == d) hot_set_fn(m.hot, Bd, fB | F_MERGED, cntB);      // neighbour's MERGED flag, by the lane that loaded it
        __syncwarp();
        changed[nchg++] = Bd;
    }
    return nchg;
}

// ------------------------------------------------------------------------- event simulation in rounds
// Active set in shared memory.  An entry is READY when no other entry has an earlier event within
// Manhattan distance MERGE_R of its voxel.  One event touches its voxel and the six neighbours
// (writes) and looks at the neighbours' neighbours (reads): footprint radius 2.  A voxel that is
// activated by an event lies within 2 of it, so a cascade of depth l stays within 2l + 2 of the
// root; with MERGE_R = 8 every event of depth <= 2 of one root is disjoint from the footprint of
// any entry that ran concurrently with (or before) that root.  Ready entries therefore commute
// with everything they do not wait for, and processing them in parallel (one warp each) gives the
// same result as the reference's strictly sequential order.  A successful merge at depth 3 would
// leave that guarantee; it is reported (E_MERGE_DEPTH) instead of being silently reordered.
constexpr int MERGE_R = 8;
constexpr int MERGE_CAP = 2048;
constexpr int MERGE_MAX_DEPTH = 2;

struct ActiveSet {
    int slot[MERGE_CAP];        // voxel slot | depth << 28
    int t[MERGE_CAP];           // next event (point index), T_INF = retired
    short kx[MERGE_CAP], ky[MERGE_CAP], kz[MERGE_CAP];   // voxel coordinate relative to the first entry (clamped)
    unsigned char ready[MERGE_CAP];
    int n;                      // entries (including retired ones until compaction)
    int ox, oy, oz;
};

__device__ __forceinline__ void as_set_key(ActiveSet& as, int k, unsigned long long pk) {
    long long x, y, z;
    unpack_key(pk, x, y, z);
    const long long dx = x - as.ox, dy = y - as.oy, dz = z - as.oz;
    // clamping only ever makes two voxels look closer -> more waiting, never less
    as.kx[k] = (short)(dx < -30000 ? -30000 : dx > 30000 ? 30000 : dx);
    as.ky[k] = (short)(dy < -30000 ? -30000 : dy > 30000 ? 30000 : dy);
    as.kz[k] = (short)(dz < -30000 ? -30000 : dz > 30000 ? 30000 : dz);
}

// insert voxel Y (first relevant event nt, cascade depth) or pull its pending event earlier; whole warp calls
__device__ void as_activate(const DevMap& m, DevCtl* ctl, ActiveSet& as, int Y, int nt, int depth) {
    const int lane = threadIdx.x & 31;
    const int n = as.n;                                   // entries appended concurrently by other warps are never Y:
    int found = -1;                                       // they lie > MERGE_R - 4 away from this warp's footprint
    for (int k = lane; k < n; k += 32) if ((as.slot[k] & 0x0FFFFFFF) == Y && as.t[k] != T_INF) found = k;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const int y = __shfl_xor_sync(0xffffffffu, found, o); found = y > found ? y : found; }
    if (lane == 0) {
        if (found >= 0) {
            atomicMin(&as.t[found], nt);
            const int d0 = as.slot[found] >> 28;
            if (depth > d0) as.slot[found] = Y | (depth << 28);
        } else {
            const int k = atomicAdd(&as.n, 1);
            if (k >= MERGE_CAP) { atomicOr(&ctl->err, E_QUEUE); atomicSub(&as.n, 1); }
            else {
                as.slot[k] = Y | (depth << 28);
                as.t[k] = nt;
                as.ready[k] = 0;
                as_set_key(as, k, m.skey[Y]);
                atomicAdd(&ctl->dbg[2], 1);
            }
        }
    }
    __syncwarp();
}

// one merge() call of entry j (voxel A at time t) and everything it triggers; whole warp calls
__device__ void process_event(const DevMap& m, DevCtl* ctl, ActiveSet& as, int j, unsigned scan_id) {
    const int lane = threadIdx.x & 31;
    const int A = as.slot[j] & 0x0FFFFFFF, depth = as.slot[j] >> 28, t = as.t[j];
    int changed[6];
    const int nchg = merge_at_warp(m, ctl, A, t, scan_id, changed);
    if (nchg > 0) {
        if (depth > MERGE_MAX_DEPTH && lane == 0) atomicOr(&ctl->err, E_MERGE_DEPTH);
        const int dn = depth + 1 > 7 ? 7 : depth + 1;
        // planes / groups of A and changed[] moved.  (b) a changed neighbour itself: all of its pairs;
        // (a) every voxel Y adjacent to a changed voxel X: only its pair with X can have flipped.
        for (int q = 0; q < nchg; q++) {
            const int Yb = changed[q];
            if (m.cnt[Yb] == 0 || m.evn[Yb] == 0) continue;
            const int fl = event_floor(m, Yb, scan_id);
            const int wb = wake_warp(m, Yb, t > fl ? t : fl, scan_id);
            const int nt = wb == T_INF ? T_INF : next_event_warp(m, Yb, wb);
            if (nt != T_INF) as_activate(m, ctl, as, Yb, nt, dn);
            else {
                // the partner has no pair left that can pass (typically: it now shares A's group): retire its
                // pending entry right away instead of spending a round on a no-op.  Its entry is not being
                // processed concurrently: it lies within 1 of A and has a later event.
                const int n = as.n;
                for (int k = lane; k < n; k += 32) if ((as.slot[k] & 0x0FFFFFFF) == Yb) as.t[k] = T_INF;
                __syncwarp();
            }
        }
        const int ncand = (nchg + 1) * 6;
        for (int c0 = 0; c0 < ncand; c0 += 32) {
            const int c = c0 + lane;
            int Y = -1, w = T_INF;
            if (c < ncand) {
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
                if (nt != T_INF) as_activate(m, ctl, as, Yh, nt, dn);
            }
        }
    }
    // advance A: sleep until the earliest time one of its pairs can pass again, retire it if none can
    int nt = T_INF;
    const int w = wake_warp(m, A, t, scan_id);
    if (w != T_INF) nt = next_event_warp(m, A, w);
    if (lane == 0) { as.t[j] = nt; atomicAdd(&ctl->dbg[1], 1); }
    __syncwarp();
}

__global__ void __launch_bounds__(512) k_merge_rounds(DevMap m, DevCtl* ctl) {
    __shared__ ActiveSet as;
    __shared__ int s_cnt;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n0 = ctl->n_hot;                       // (voxel, first relevant event) pairs prepared by k_merge_prefilter
    if (n0 == 0) return;
    if (n0 > MERGE_CAP) { if (tid == 0) atomicOr(&ctl->err, E_QUEUE); return; }
    const unsigned scan_id = ctl->scan_id;
    if (tid == 0) {
        long long x, y, z;
        unpack_key(m.skey[m.act_slot[0]], x, y, z);
        as.ox = (int)x; as.oy = (int)y; as.oz = (int)z;
        as.n = n0;
        ctl->dbg[0] = n0;
    }
    __syncthreads();
    for (int k = tid; k < n0; k += blockDim.x) {
        const int A = m.act_slot[k];
        as.slot[k] = A;
        as.t[k] = m.act_t[k];
        as_set_key(as, k, m.skey[A]);
    }
    __syncthreads();
    for (int round = 0; round < 100000; round++) {
        const int n = as.n;
        // readiness: no other entry with an earlier event within MERGE_R
        for (int j = tid; j < n; j += blockDim.x) {
            const int tj = as.t[j];
            bool rdy = tj != T_INF;
            if (rdy) {
                const int x = as.kx[j], y = as.ky[j], z = as.kz[j];
                for (int i = 0; i < n; i++) {
                    if (as.t[i] < tj && abs(as.kx[i] - x) + abs(as.ky[i] - y) + abs(as.kz[i] - z) <= MERGE_R) { rdy = false; break; }
                }
            }
            as.ready[j] = rdy ? 1 : 0;
        }
        __syncthreads();
        for (int j = wid; j < n; j += (blockDim.x >> 5)) {
            if (as.ready[j]) process_event(m, ctl, as, j, scan_id);      // warp-uniform branch
        }
        __syncthreads();
        // compaction of retired entries (one warp), keeps the rest in place order
        if (wid == 0) {
            const int nn = as.n;
            int w = 0;
            for (int base = 0; base < nn; base += 32) {
                const int k = base + lane;
                const bool live = k < nn && as.t[k] != T_INF;
                const int sl = live ? as.slot[k] : 0, tt = live ? as.t[k] : 0;
                const short kx = live ? as.kx[k] : 0, ky = live ? as.ky[k] : 0, kz = live ? as.kz[k] : 0;
                const unsigned bal = __ballot_sync(0xffffffffu, live);
                const int pos = w + __popc(bal & ((1u << lane) - 1));
                __syncwarp();
                if (live) { as.slot[pos] = sl; as.t[pos] = tt; as.kx[pos] = kx; as.ky[pos] = ky; as.kz[pos] = kz; }
                w += __popc(bal);
                __syncwarp();
            }
            if (lane == 0) { as.n = w; s_cnt = w; }
        }
        __syncthreads();
        if (s_cnt == 0) break;
    }
}
