// vmp_merge.cuh — VoxelGrid::merge() (voxel_map.cpp:138-186) on the device.  Included by
// vmp_map.cu inside namespace vmp.
//
// merge() runs for EVERY point that lands in a full plane voxel (Q11) and is a no-op almost
// always.  Phase 1 (k_merge_prefilter, parallel, 8 lanes per voxel): which of the voxels that
// received such points could merge with some neighbour at all, given the current planes and
// groups, ignoring when during the scan the neighbour becomes eligible.  Phase 2
// (k_merge_rounds, one CTA, one warp per event): an event simulation over those voxels only, in
// rounds of spatially independent events (see the footprint argument further down).
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

// Everything the merge logic wants to know about one voxel, fetched with INDEPENDENT loads (one memory latency);
// these kernels are chains of dependent look-ups, so the number of round trips is what they cost.
struct VoxRec {
    V3 mean, nrm;
    unsigned long long group;
    uint32_t flags;
    long long w6;                       // raw flags | n word of the hot record
    unsigned born_scan, full_scan;
    int ft, full_idx, evict_t, ghost;
};
__device__ __forceinline__ VoxRec load_rec(const DevMap& m, int s) {
    VoxRec r;
    const double2* h2 = reinterpret_cast<const double2*>(m.hot + (size_t)s * 8);        // 64-byte records
    const double2 q0 = h2[0], q1 = h2[1], q2 = h2[2], q3 = h2[3];
    r.group = m.sgroup[s];
    r.born_scan = m.born_scan[s]; r.full_scan = m.full_scan[s];
    r.ft = m.ft[s]; r.full_idx = m.full_idx[s]; r.evict_t = m.evict_t[s]; r.ghost = m.ghost[s];
    r.mean = v3(q0.x, q0.y, q1.x);
    r.nrm = v3(q1.y, q2.x, q2.y);
    r.w6 = __double_as_longlong(q3.x);
    r.flags = (uint32_t)(r.w6 & 0xFFFFFFFFll);
    return r;
}
// could merge(A) succeed against the (final state of) incarnation B at some time of this scan?
__device__ __forceinline__ bool pair_static(const DevMap& m, const V3& mA, const V3& nA, unsigned long long gA, const VoxRec& B) {
    if ((B.flags & F_UE) || !(B.flags & F_PLANE)) return false;
    if (B.group == gA) return false;
    return plane_thresholds(m, mA, nA, B.mean, B.nrm);
}
// window (start, end) of point indices t of this scan at which B is alive and full: start < t < end
__device__ __forceinline__ void pair_window(const VoxRec& B, unsigned scan_id, int& start, int& end) {
    const int born = (B.born_scan == scan_id) ? B.ft : -1;
    const int full = (B.full_scan == scan_id) ? B.full_idx : -1;
    start = born > full ? born : full;
    end = B.evict_t;
}
__device__ __forceinline__ int event_floor(const VoxRec& A, unsigned scan_id) {
    return (A.full_scan == scan_id) ? A.full_idx : -1;           // merge() runs only for points after the closing one
}

// one neighbour direction of "could merge(A) succeed after time `after`": returns the earliest time
// bound (events with t > bound may succeed), T_INF if never.  keyA / mA / nA / gA: A's key, plane and group.
__device__ __forceinline__ int wake_dir(const DevMap& m, unsigned long long keyA, const V3& mA, const V3& nA, unsigned long long gA,
                                        int d, int after, unsigned scan_id) {
    bool ok;
    const unsigned long long nk = nbr_key(keyA, d, ok);
    if (!ok) return T_INF;
    int best = T_INF;
    int B = hash_find(m, nk);
    while (B >= 0) {
        const VoxRec r = load_rec(m, B);
        if (pair_static(m, mA, nA, gA, r)) {
            int s, e;
            pair_window(r, scan_id, s, e);
            const int bound = s > after ? s : after;             // first usable events are those with t > bound
            if (bound + 1 < e && bound < best) best = bound;
        }
        B = r.ghost;
    }
    return best;
}

// 8 lanes per touched voxel, lanes 0..5 take one neighbour each; voxels whose merge() can succeed at
// some time of this scan go into the active set of k_merge_rounds with their first relevant event
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
            if (m.evn[A] > 0 && sub < 6) {
                const unsigned long long keyA = m.skey[A];
                const VoxRec ra = load_rec(m, A);
                w = wake_dir(m, keyA, ra.mean, ra.nrm, ra.group, sub, event_floor(ra, scan_id), scan_id);
            }
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

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// VoxelGrid::merge() of voxel A at time t, executed by the whole warp.  Returns the number of
// successful pair merges; changed[0..n) (warp-uniform) are the neighbour slots that were modified.
__device__ int merge_at_warp(const DevMap& m, DevCtl* ctl, int A, int t, unsigned scan_id, int* changed) {
    const int lane = threadIdx.x & 31;
    // A's own record and covariance (issued first: these loads overlap the neighbour look-ups below);
    // lanes hold the 36 covariance entries (lane e: entry e, lanes 0..3 also entry 32+e)
    const unsigned long long keyA = m.skey[A];
    const VoxRec ra = load_rec(m, A);
    double* ca = m.cov + (size_t)A * 36;
    double a0 = ca[lane], a1 = lane < 4 ? ca[32 + lane] : 0.0;
    const unsigned long long gA = ra.group;
    V3 mA = ra.mean, nA = ra.nrm;
    // ---- parallel part: lane d < 6 resolves neighbour d as of time t (one record fetch per incarnation)
    int B = -1;
    bool elig = false;
    VoxRec rb;
    rb.mean = v3(0, 0, 0); rb.nrm = v3(0, 0, 0); rb.w6 = 0;
    if (lane < 6) {
        bool ok;
        const unsigned long long nk = nbr_key(keyA, lane, ok);
        if (ok) {
            int X = hash_find(m, nk);
            while (X >= 0) {                                          // incarnation alive at time t
                const VoxRec r = load_rec(m, X);
                const int born = (r.born_scan == scan_id) ? r.ft : -1;
                if (born <= t && t < r.evict_t) { B = X; rb = r; break; }
                X = r.ghost;
            }
        }
        if (B >= 0) {
            const bool closed = !(rb.flags & F_UE) && (rb.full_scan != scan_id || rb.full_idx < t);
            // the neighbours are distinct voxels and gA does not change: group test once, up front
            elig = closed && (rb.flags & F_PLANE) && rb.group != gA;
        }
    }
    const unsigned elmask = __ballot_sync(0xffffffffu, elig) & 0x3Fu;
    // covariances of every eligible neighbour, fetched together
    double b0p[6], b1p[6];
#pragma unroll
    for (int d = 0; d < 6; d++) {
        b0p[d] = 0.0; b1p[d] = 0.0;
        const int Bd = __shfl_sync(0xffffffffu, B, d);
        if ((elmask >> d) & 1u) {
            const double* cb = m.cov + (size_t)Bd * 36;
            b0p[d] = cb[lane];
            if (lane < 4) b1p[d] = cb[32 + lane];
        }
    }
    // ---- ordered part (warp-uniform control flow): -x -y -z +x +y +z, own plane updated in between (Q10)
    int nchg = 0;
#pragma unroll
    for (int d = 0; d < 6; d++) {
        if (!((elmask >> d) & 1u)) continue;
        const int Bd = __shfl_sync(0xffffffffu, B, d);
        V3 mb, nb;
#pragma unroll
        for (int k = 0; k < 3; k++) { mb[k] = shfl_d(rb.mean[k], d); nb[k] = shfl_d(rb.nrm[k], d); }
        if (!plane_thresholds(m, mA, nA, mb, nb)) continue;
        double* cb = m.cov + (size_t)Bd * 36;
        const double b0 = b0p[d], b1 = b1p[d];
        const double tn0 = shfl_d(a0, 0) + shfl_d(a0, 7) + shfl_d(a0, 14);
        const double tm0 = shfl_d(a0, 21) + shfl_d(a0, 28) + shfl_d(a1, 3);
        const double tn1 = shfl_d(b0, 0) + shfl_d(b0, 7) + shfl_d(b0, 14);
        const double tm1 = shfl_d(b0, 21) + shfl_d(b0, 28) + shfl_d(b1, 3);
        const double tc0 = tn0 + tm0, tc1 = tn1 + tm1;
        // Q9: operator precedence exactly as in voxel_map.cpp:166-167.  The six divisions run on six lanes (lane k:
        // mean[k], lane 3 + k: normal[k]) and are broadcast: a single warp pays per instruction, not per lane
        V3 nm, nn;
        {
            const int kk = lane % 3;
            const bool isn = lane >= 3;
            const double xb = isn ? (kk == 0 ? nb[0] : kk == 1 ? nb[1] : nb[2]) : (kk == 0 ? mb[0] : kk == 1 ? mb[1] : mb[2]);
            const double xa = isn ? (kk == 0 ? nA[0] : kk == 1 ? nA[1] : nA[2]) : (kk == 0 ? mA[0] : kk == 1 ? mA[1] : mA[2]);
            const double t0 = isn ? tn0 : tm0, t1 = isn ? tn1 : tm1;
            const double v = xb * t0 + (xa * t1) / (t0 + t1);
#pragma unroll
            for (int k = 0; k < 3; k++) { nm[k] = shfl_d(v, k); nn[k] = shfl_d(v, 3 + k); }
        }
        const double w0 = tc0 * tc0, w1 = tc1 * tc1, den = (tc0 + tc1) * (tc0 + tc1);
        const double c0 = (b0 * w0 + a0 * w1) / den;
        a0 = c0; cb[lane] = c0;                                    // A's copy stays in registers until the end
        if (lane < 4) { const double c1 = (b1 * w0 + a1 * w1) / den; a1 = c1; cb[32 + lane] = c1; }
        if (-dot(nm, nn) < 0.0) nn = neg(nn);
        mA = nm; nA = nn;
        if (lane < 3) { m.hot[(size_t)Bd * 8 + lane] = nm[lane]; m.hot[(size_t)Bd * 8 + 3 + lane] = nn[lane]; }
        if (lane == 0) {
            m.sgroup[Bd] = gA;
            atomicAdd((unsigned long long*)&ctl->st.n_merge, 1ull);
            changed[nchg] = Bd;
        }
        if (lane == d) m.hot[(size_t)Bd * 8 + 6] = __longlong_as_double(rb.w6 | (long long)F_MERGED);   // by the lane that loaded it
        nchg++;
    }
    if (nchg > 0) {
        ca[lane] = a0;
        if (lane < 4) ca[32 + lane] = a1;
        if (lane < 3) { m.hot[(size_t)A * 8 + lane] = mA[lane]; m.hot[(size_t)A * 8 + 3 + lane] = nA[lane]; }
        if (lane == 0) m.hot[(size_t)A * 8 + 6] = __longlong_as_double(ra.w6 | (long long)F_MERGED);
    }
    __syncwarp();
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
    short rlist[MERGE_CAP];     // entries that are ready in the current round
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
            if (k >= MERGE_CAP) { atomicOr(&ctl->err, E_MERGE_CAP); atomicSub(&as.n, 1); }
            else {
                as.slot[k] = Y | (depth << 28);
                as.t[k] = nt;
                as_set_key(as, k, m.skey[Y]);
                atomicAdd(&ctl->dbg[2], 1);
            }
        }
    }
    __syncwarp();
}

// per-warp scratch of the follow-up work of one event
constexpr int FOLLOW_CAP = 52;          // (nchg + 1) wakes + 6 (nchg + 1) candidates, nchg <= 6
struct WarpScratch {
    int chg[8];                         // neighbour slots modified by the event
    int tgt[FOLLOW_CAP], aft[FOLLOW_CAP], res[FOLLOW_CAP];
    unsigned char kind[FOLLOW_CAP];     // 0: the event's own voxel, 1: a modified neighbour, 2: a voxel next to a modified one
};

// one merge() call of entry j (voxel A at time t) and everything it triggers; whole warp calls.
// The follow-up look-ups are independent chains of dependent global loads (hash probe -> slot -> plane ...), so
// they run one per lane, eight lanes per voxel: which voxels around the event can merge now, and when is their
// next merge() call.  The results are applied afterwards in the order of the reference.
__device__ void process_event(const DevMap& m, DevCtl* ctl, ActiveSet& as, WarpScratch& ws, int j, unsigned scan_id) {
    const int lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & 7;
    const int A = as.slot[j] & 0x0FFFFFFF, depth = as.slot[j] >> 28, t = as.t[j];
    const int nchg = merge_at_warp(m, ctl, A, t, scan_id, ws.chg);
    if (nchg > 0 && depth > MERGE_MAX_DEPTH && lane == 0) atomicOr(&ctl->err, E_MERGE_DEPTH);
    const int dn = depth + 1 > 7 ? 7 : depth + 1;
    // task groups of 8 lanes (6 used, one neighbour direction each):
    //   g = 0            the event's own voxel: earliest time one of its pairs can pass again
    //   g = 1..nchg      (b) a modified neighbour itself: all of its pairs
    //   g > nchg         (a) the voxels Y adjacent to X = A / a modified neighbour: only the pair (Y, X) can have flipped
    const int ngroups = nchg > 0 ? 2 * nchg + 2 : 1;
    int nq = 0;
    for (int g0 = 0; g0 < ngroups; g0 += 4) {
        const int g = g0 + grp;
        const bool wake_grp = g <= nchg;
        int Y = -1, w = T_INF;
        if (g < ngroups && sub < 6) {
            if (wake_grp) {
                const int X = g == 0 ? A : ws.chg[g - 1];
                const unsigned long long keyX = m.skey[X];
                const VoxRec rx = load_rec(m, X);
                const int cx = m.cnt[X], ex = m.evn[X];
                if (g == 0 || (cx != 0 && ex != 0)) {
                    int after = t;
                    if (g > 0) { const int fl = event_floor(rx, scan_id); after = t > fl ? t : fl; }
                    w = wake_dir(m, keyX, rx.mean, rx.nrm, rx.group, sub, after, scan_id);
                    Y = X;
                }
            } else {
                const int xi = g - nchg - 1;
                const int X = xi == 0 ? A : ws.chg[xi - 1];
                const unsigned long long keyX = m.skey[X];
                const VoxRec rx = load_rec(m, X);
                bool ok;
                const unsigned long long nk = nbr_key(keyX, sub, ok);
                Y = ok ? hash_find(m, nk) : -1;
                if (Y == A) Y = -1;
                for (int q = 0; q < nchg && Y >= 0; q++) if (ws.chg[q] == Y) Y = -1;  // handled by (b)
                if (Y >= 0) {
                    const int cy = m.cnt[Y], ey = m.evn[Y];
                    const VoxRec ry = load_rec(m, Y);
                    if (cy == 0 || ey == 0) Y = -1;                                   // no merge() call of Y in this scan
                    else if (pair_static(m, ry.mean, ry.nrm, ry.group, rx)) {
                        int s0, e0;
                        pair_window(rx, scan_id, s0, e0);
                        const int fl = event_floor(ry, scan_id);
                        int bound = s0 > t ? s0 : t;
                        bound = fl > bound ? fl : bound;
                        if (bound + 1 < e0) w = bound;
                    }
                }
                if (w == T_INF) Y = -1;
            }
        }
        int wmin = w;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) { const int y = __shfl_xor_sync(0xffffffffu, wmin, o); wmin = y < wmin ? y : wmin; }
        // queue the voxels whose next merge() call has to be looked up
        const bool q_me = g < ngroups && Y >= 0 && (wake_grp ? sub == 0 : true);
        const unsigned bal = __ballot_sync(0xffffffffu, q_me);
        if (q_me) {
            const int pos = nq + __popc(bal & ((1u << lane) - 1));
            if (pos < FOLLOW_CAP) {
                ws.tgt[pos] = Y;
                ws.aft[pos] = wake_grp ? wmin : w;
                ws.kind[pos] = (unsigned char)(g == 0 ? 0 : wake_grp ? 1 : 2);
            }
        }
        nq += __popc(bal);
    }
    if (nq > FOLLOW_CAP) { if (lane == 0) atomicOr(&ctl->err, E_QUEUE); nq = FOLLOW_CAP; }
    __syncwarp();
    // first point index of each queued voxel after its bound: four voxels at a time, eight lanes each
    for (int q0 = 0; q0 < nq; q0 += 4) {
        const int q = q0 + grp;
        int best = T_INF;
        if (q < nq) {
            const int Yq = ws.tgt[q], aq = ws.aft[q];
            if (aq != T_INF) {
                const int c = m.cnt[Yq], off = m.seg_off[Yq];
                for (int k = sub; k < c; k += 8) { const int i = m.seg[off + k]; if (i > aq && i < best) best = i; }
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) { const int y = __shfl_xor_sync(0xffffffffu, best, o); best = y < best ? y : best; }
        if (q < nq && sub == 0) ws.res[q] = best;
    }
    __syncwarp();
    // apply: activations / retirements of the neighbourhood, then A's own next event
    int ntA = T_INF;
    for (int q = 0; q < nq; q++) {
        const int kind = ws.kind[q], Yq = ws.tgt[q], nt = ws.res[q];
        if (kind == 0) ntA = nt;
        else if (nt != T_INF) as_activate(m, ctl, as, Yq, nt, dn);
        else if (kind == 1) {
            // the partner has no pair left that can pass (typically: it now shares A's group): retire its
            // pending entry right away instead of spending a round on a no-op.  Its entry is not being
            // processed concurrently: it lies within 1 of A and has a later event.
            const int n = as.n;
            for (int k = lane; k < n; k += 32) if ((as.slot[k] & 0x0FFFFFFF) == Yq) as.t[k] = T_INF;
            __syncwarp();
        }
    }
    if (lane == 0) { as.t[j] = ntA; atomicAdd(&ctl->dbg[1], 1); }
    __syncwarp();
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// pull everything merge() will ask about slot s into L2 (the fields of VoxRec, the covariance, the per-scan lists)
__device__ __forceinline__ void prefetch_slot(const DevMap& m, int s) {
    const char* h = reinterpret_cast<const char*>(m.hot + (size_t)s * 8);
    prefetch_l2(h); prefetch_l2(h + 32);
    const char* c = reinterpret_cast<const char*>(m.cov + (size_t)s * 36);
    prefetch_l2(c); prefetch_l2(c + 128); prefetch_l2(c + 256);
    prefetch_l2(m.sgroup + s); prefetch_l2(m.born_scan + s); prefetch_l2(m.full_scan + s); prefetch_l2(m.ft + s);
    prefetch_l2(m.full_idx + s); prefetch_l2(m.evict_t + s); prefetch_l2(m.cnt + s); prefetch_l2(m.evn + s); prefetch_l2(m.seg_off + s);
    prefetch_l2(m.skey + s);
}

__global__ void __launch_bounds__(512) k_merge_rounds(DevMap m, DevCtl* ctl) {
    __shared__ ActiveSet as;
    __shared__ WarpScratch wsc[16];
    __shared__ int s_cnt, s_nready;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n0 = ctl->n_hot;                       // (voxel, first relevant event) pairs prepared by k_merge_prefilter
    if (n0 == 0) return;
    if (n0 > MERGE_CAP) { if (tid == 0) atomicOr(&ctl->err, E_MERGE_CAP); return; }
    const unsigned scan_id = ctl->scan_id;
    if (tid == 0) {
        long long x, y, z;
        unpack_key(m.skey[m.act_slot[0]], x, y, z);
        as.ox = (int)x; as.oy = (int)y; as.oz = (int)z;
        as.n = n0; s_nready = 0;
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
    // The rounds below are chains of dependent look-ups (neighbour key -> hash -> slot -> record -> covariance) by a
    // single CTA; after the fill kernels have streamed hundreds of MB through the L2 (200 k-point scans) every link of the
    // chain is a DRAM round trip.  One parallel pass first resolves the six neighbours of every active voxel and
    // prefetches what merge() reads about them, so that the serial part runs out of the L2.
    for (int k = tid; k < n0 * 8; k += blockDim.x) {
        const int j = k >> 3, d = k & 7;
        const int A = as.slot[j];
        if (d == 6) { prefetch_slot(m, A); continue; }
        if (d == 7) continue;
        bool ok;
        const unsigned long long nk = nbr_key(m.skey[A], d, ok);
        if (!ok) continue;
        for (int X = hash_find(m, nk); X >= 0; X = m.ghost[X]) prefetch_slot(m, X);
    }
    for (int round = 0; round < 100000; round++) {
        const int n = as.n;
        // readiness (one warp per entry, lanes over the others): no other entry with an earlier event within MERGE_R
        for (int j = wid; j < n; j += 16) {
            const int tj = as.t[j];
            if (tj == T_INF) continue;
            const int x = as.kx[j], y = as.ky[j], z = as.kz[j];
            bool conflict = false;
            for (int i = lane; i < n; i += 32)
                if (as.t[i] < tj && abs(as.kx[i] - x) + abs(as.ky[i] - y) + abs(as.kz[i] - z) <= MERGE_R) conflict = true;
            if (!__any_sync(0xffffffffu, conflict) && lane == 0) as.rlist[atomicAdd(&s_nready, 1)] = (short)j;
        }
        __syncthreads();
        const int nr = s_nready;
        for (int q = wid; q < nr; q += 16) process_event(m, ctl, as, wsc[wid], as.rlist[q], scan_id);   // one ready event per warp
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
            if (lane == 0) { as.n = w; s_cnt = w; s_nready = 0; atomicAdd(&ctl->dbg[2], 65536); }      // (rounds in the upper half of dbg[2])
        }
        __syncthreads();
        if (s_cnt == 0) break;
    }
}
