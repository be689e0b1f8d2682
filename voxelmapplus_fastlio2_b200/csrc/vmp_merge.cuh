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
struct CellRec {
    double mean[3], nrm[3];
    unsigned long long group;
    long long w6;                       // raw flags | n word of the hot record
    unsigned born_scan, full_scan;
    int ft, full_idx, evict_t, ghost;
    int slot, cnt, evn, pad;
};
constexpr int NCELL = 25;

// cell numbering: 0 = A; 1 + d = neighbour d (-x -y -z +x +y +z); 7 + d = two steps along d; 13 + 4 p + 2 [sa > 0] + [sb > 0] =
// one step along each axis of pair p (xy, xz, yz).  Offsets and the neighbour-of-a-cell relation as constant tables.
__constant__ signed char c_cell_off[NCELL][3] = {{0, 0, 0}, {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {-2, 0, 0}, {0, -2, 0}, {0, 0, -2}, {2, 0, 0}, {0, 2, 0}, {0, 0, 2}, {-1, -1, 0}, {-1, 1, 0}, {1, -1, 0}, {1, 1, 0}, {-1, 0, -1}, {-1, 0, 1}, {1, 0, -1}, {1, 0, 1}, {0, -1, -1}, {0, -1, 1}, {0, 1, -1}, {0, 1, 1}};
__constant__ signed char c_cell_nb[7][6] = {{1, 2, 3, 4, 5, 6}, {7, 13, 17, 0, 14, 18}, {13, 8, 21, 15, 0, 22}, {17, 21, 9, 19, 23, 0}, {0, 15, 19, 10, 16, 20}, {14, 0, 23, 16, 11, 24}, {18, 22, 0, 20, 24, 12}};
__device__ __forceinline__ void cell_offset(int c, int& dx, int& dy, int& dz) { dx = c_cell_off[c][0]; dy = c_cell_off[c][1]; dz = c_cell_off[c][2]; }
// neighbour d of cell c (c <= 6), as a cell
__device__ __forceinline__ int cell_neighbour(int c, int d) { return c_cell_nb[c][d]; }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// pull everything merge() will ask about slot s into L2 (the fields of VoxRec, the covariance, the per-scan lists)
__device__ __forceinline__ void prefetch_slot(const DevMap& m, int s, bool with_cov = true) {
    const char* h = reinterpret_cast<const char*>(m.hot + (size_t)s * 8);
    prefetch_l2(h); prefetch_l2(h + 32);
    if (with_cov) {
        const char* c = reinterpret_cast<const char*>(m.cov + (size_t)s * 36);
        prefetch_l2(c); prefetch_l2(c + 128); prefetch_l2(c + 256);
    }
    prefetch_l2(m.sgroup + s); prefetch_l2(m.born_scan + s); prefetch_l2(m.full_scan + s); prefetch_l2(m.ft + s);
    prefetch_l2(m.full_idx + s); prefetch_l2(m.evict_t + s); prefetch_l2(m.cnt + s); prefetch_l2(m.evn + s); prefetch_l2(m.seg_off + s);
    prefetch_l2(m.skey + s);
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
        int w = T_INF, A = -1, c = 0, off = 0;
        if (vi < V) {
            A = m.touched[vi];
            // everything about A in ONE round trip (the kernel is a chain of dependent look-ups per voxel: each one it does not wait for
            // separately is ~1 us off a launch that lasts 17): event count, key, record, segment
            const int evn = m.evn[A];
            const unsigned long long keyA = m.skey[A];
            const VoxRec ra = load_rec(m, A);
            c = m.cnt[A]; off = m.seg_off[A];
            if (evn > 0 && sub < 6) w = wake_dir(m, keyA, ra.mean, ra.nrm, ra.group, sub, event_floor(ra, scan_id), scan_id);
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) { const int y = __shfl_xor_sync(gmask, w, o); w = y < w ? y : w; }
        if (w == T_INF) continue;                                   // uniform within the 8-lane group
        // first merge() call of A after the earliest time one of its pairs can pass
        int best = T_INF;
        for (int q = sub; q < c; q += 8) { const int i = m.seg[off + q]; if (i > w && i < best) best = i; }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) { const int y = __shfl_xor_sync(gmask, best, o); best = y < best ? y : best; }
        if (best != T_INF && sub == 0) {
            const int k = atomicAdd(&ctl->n_hot, 1);
            m.act_slot[k] = A;
            m.act_t[k] = best;
        }
        if (best != T_INF) {
            // an active voxel: its event will look at everything within Manhattan distance 2 (vmp_merge.cuh, "diamond"); pull those
            // records (and the covariances merge() may blend) into the L2 now, many voxels at a time, instead of in the
            // dependent chain of the single CTA that simulates the events
            long long x, y, z;
            unpack_key(m.skey[A], x, y, z);
            for (int c = sub; c < NCELL; c += 8) {
                int dx, dy, dz;
                cell_offset(c, dx, dy, dz);
                const long long cx = x + dx, cy = y + dy, cz = z + dz;
                if (!(key_in_range(cx) && key_in_range(cy) && key_in_range(cz))) continue;
                for (int X = c == 0 ? A : hash_find(m, pack_key(cx, cy, cz)); X >= 0; X = m.ghost[X]) prefetch_slot(m, X, c <= 6);
            }
        }
    }
}

// ---- undo log of the merge phase: the state of a voxel before its modification (covariance, mean / normal / flags word, group), so
// that the parallel rounds can be taken back and the scan's merge() calls redone in strict event order (k_merge_rounds)
// An event reserves its UNDO_EVENT entries (its own voxel + six neighbours) with ONE atomic at its start - the index comes back while
// the event is still fetching its neighbourhood; a returning atomic per record was two dependent L2 round trips per merge - and marks
// the ones it did not use (slot -1).  Events that touch the same voxel run in different rounds, so their entries are in time order.
constexpr int UNDO_W = 44;              // doubles per entry: cov[36], mean[3], nrm[3], w6 bits, group bits
constexpr int UNDO_EVENT = 7;
__device__ __forceinline__ int undo_reserve(DevCtl* ctl) {                  // lane 0's value counts
    return (threadIdx.x & 31) == 0 ? atomicAdd(&ctl->n_undo, UNDO_EVENT) : 0;
}
__device__ __forceinline__ void undo_close(const DevMap& m, int base, int used) {
    const int lane = threadIdx.x & 31;
    if (lane >= used && lane < UNDO_EVENT && base + lane < m.undo_cap) m.undo_slot[base + lane] = -1;
}
__device__ __forceinline__ void undo_log(const DevMap& m, int e, int slot, double c0, double c1, const double* mean, const double* nrm, long long w6,
                                         unsigned long long group) {
    const int lane = threadIdx.x & 31;
    if (e >= m.undo_cap) return;                                            // (a redo would then be reported instead of performed)
    double* r = m.undo_rec + (size_t)e * UNDO_W;
    r[lane] = c0;
    if (lane < 4) r[32 + lane] = c1;
    if (lane < 3) { r[36 + lane] = mean[lane]; r[39 + lane] = nrm[lane]; }
    if (lane == 0) { r[42] = __longlong_as_double(w6); r[43] = __longlong_as_double((long long)group); m.undo_slot[e] = slot; }
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
// component i of a vector held in registers (v[i] with a run-time i would move the vector to local memory)
__device__ __forceinline__ double sel3(const V3& v, int i) { return i == 0 ? v[0] : i == 1 ? v[1] : v[2]; }

// VoxelGrid::merge() of voxel A at time t, executed by the whole warp.  Returns the number of
// successful pair merges; changed[0..n) (warp-uniform) are the neighbour slots that were modified.
__device__ int merge_at_warp(const DevMap& m, DevCtl* ctl, int A, int t, unsigned scan_id, int* changed) {
    const int lane = threadIdx.x & 31;
    // A's own record and covariance (issued first: these loads overlap the neighbour look-ups below);
    // lanes hold the 36 covariance entries (lane e: entry e, lanes 0..3 also entry 32+e)
    const int ubase_l0 = undo_reserve(ctl);
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
        {       // undo log: A before its first modification of this event, B before this one
            double m3[3], n3[3];
            const int ubase = __shfl_sync(0xffffffffu, ubase_l0, 0);
            if (nchg == 0) {
                for (int k = 0; k < 3; k++) { m3[k] = ra.mean[k]; n3[k] = ra.nrm[k]; }
                undo_log(m, ubase, A, a0, a1, m3, n3, ra.w6, gA);
            }
            for (int k = 0; k < 3; k++) { m3[k] = mb[k]; n3[k] = nb[k]; }
            undo_log(m, ubase + 1 + nchg, Bd, b0, b1, m3, n3, __shfl_sync(0xffffffffu, rb.w6, d), __shfl_sync(0xffffffffu, rb.group, d));
        }
        const double c0 = (b0 * w0 + a0 * w1) / den;
        a0 = c0; cb[lane] = c0;                                    // A's copy stays in registers until the end
        if (lane < 4) { const double c1 = (b1 * w0 + a1 * w1) / den; a1 = c1; cb[32 + lane] = c1; }
        if (-dot(nm, nn) < 0.0) nn = neg(nn);
        mA = nm; nA = nn;
        if (lane < 3) { m.hot[(size_t)Bd * 8 + lane] = sel3(nm, lane); m.hot[(size_t)Bd * 8 + 3 + lane] = sel3(nn, lane); }
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
        if (lane < 3) { m.hot[(size_t)A * 8 + lane] = sel3(mA, lane); m.hot[(size_t)A * 8 + 3 + lane] = sel3(nA, lane); }
        if (lane == 0) m.hot[(size_t)A * 8 + 6] = __longlong_as_double(ra.w6 | (long long)F_MERGED);
    }
    undo_close(m, __shfl_sync(0xffffffffu, ubase_l0, 0), nchg > 0 ? nchg + 1 : 0);
    __syncwarp();
    return nchg;
}

// ------------------------------------------------------------------------- event simulation in rounds
// Active set in shared memory.  One event touches its voxel and the six neighbours (writes) and looks at the
// neighbours' neighbours (reads): footprint radius 2, so two events commute when their voxels are more than
// MERGE_R = 4 apart (Manhattan).  An entry is READY when no other PENDING entry has an earlier event within MERGE_R
// of its voxel; ready entries are executed in parallel, one warp each.  That is the reference's sequential order as
// long as the set of pending events is complete - but an event can create new ones (a merge changes planes, after
// which the voxels around them are re-examined: as_activate).  A new event (Y, nt) comes too late if an event E with
// t_E > nt and |E - Y| <= MERGE_R has already been started (earlier round, or the same round on another warp): the
// reference would have run (Y, nt) first.  Every started event is therefore kept in a history, every activation is
// checked against it, and a hit - rare: an activation happens in one scan of three on the BASELINE workloads, and
// then it has to land next to a later event - takes the whole scan to the exact serial mode below (undo log).
// (Round 2 first used MERGE_R = 8 with a cascade-depth bound instead: 6.7 rounds per C2 scan; 4.5 with this rule.  The exact conflict
// radius is 3 - writes reach 1, reads 2 -: measured 3.8 rounds, but then one C2 scan in 68 meets a late activation, and a serial redo
// costs a hundred times what the saved round does; 4 keeps them at 0 of 68 C2 scans and 0 of 12 500 C3 scans.)
constexpr int MERGE_R = 4;
constexpr int MERGE_CAP = 2048;          // (DevMap::merge_cap <= MERGE_CAP)
constexpr int MERGE_HIST = 4096;         // started events per scan the parallel rounds can remember (more: serial mode)

struct ActiveSet {
    int slot[MERGE_CAP];        // voxel slot | depth << 28
    int t[MERGE_CAP];           // next event (point index), T_INF = retired
    short kx[MERGE_CAP], ky[MERGE_CAP], kz[MERGE_CAP];   // voxel coordinate relative to the first entry (clamped)
    short rlist[MERGE_CAP];     // entries that are ready in the current round
    int ht[MERGE_HIST];         // history of started events: time ...
    short hx[MERGE_HIST], hy[MERGE_HIST], hz[MERGE_HIST];      // ... and voxel
    int n;                      // entries (including retired ones until compaction)
    int hn;
    int ox, oy, oz;
};

__device__ __forceinline__ void as_rel_key(const ActiveSet& as, unsigned long long pk, short& rx, short& ry, short& rz) {
    long long x, y, z;
    unpack_key(pk, x, y, z);
    const long long dx = x - as.ox, dy = y - as.oy, dz = z - as.oz;
    // clamping only ever makes two voxels look closer -> more waiting, never less
    rx = (short)(dx < -30000 ? -30000 : dx > 30000 ? 30000 : dx);
    ry = (short)(dy < -30000 ? -30000 : dy > 30000 ? 30000 : dy);
    rz = (short)(dz < -30000 ? -30000 : dz > 30000 ? 30000 : dz);
}
__device__ __forceinline__ void as_set_key(ActiveSet& as, int k, unsigned long long pk) { as_rel_key(as, pk, as.kx[k], as.ky[k], as.kz[k]); }

// What the event code sees of the active set: the shared-memory arrays of the parallel rounds, or (exact serial mode, see
// k_merge_rounds) the global arrays the prefilter filled.  redo: set when the parallel rounds cannot guarantee the reference's
// order any more (a merge succeeded deeper in a cascade than max_depth, or the shared arrays are full).
struct SetView {
    int* slot;                  // voxel slot | depth << 28
    int* t;                     // next event (point index), T_INF = retired
    int* n;                     // entries
    int cap;
    ActiveSet* keys;            // voxel coordinates for the readiness test (null in serial mode)
    int* redo;
    int max_depth;
};

// insert voxel Y (first relevant event nt, cascade depth) or pull its pending event earlier; whole warp calls
__device__ void as_activate(const DevMap& m, DevCtl* ctl, const SetView& sv, int Y, int nt, int depth) {
    const int lane = threadIdx.x & 31;
    const int n = *sv.n;                                  // entries appended concurrently by other warps are never Y: concurrent
    int found = -1;                                       // events are > MERGE_R = 4 apart and activate voxels within 2 of themselves
    for (int k = lane; k < n; k += 32) if ((sv.slot[k] & 0x0FFFFFFF) == Y && sv.t[k] != T_INF) found = k;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const int y = __shfl_xor_sync(0xffffffffu, found, o); found = y > found ? y : found; }
    if (sv.keys) {                                        // parallel rounds: does (Y, nt) come too late for an event that has already been started?
        const ActiveSet& as = *sv.keys;
        short yx = 0, yy = 0, yz = 0;
        if (lane == 0) as_rel_key(as, m.skey[Y], yx, yy, yz);
        yx = (short)__shfl_sync(0xffffffffu, (int)yx, 0); yy = (short)__shfl_sync(0xffffffffu, (int)yy, 0); yz = (short)__shfl_sync(0xffffffffu, (int)yz, 0);
        const int hn = as.hn < MERGE_HIST ? as.hn : MERGE_HIST;
        bool late = false;
        for (int k = lane; k < hn; k += 32)
            if (as.ht[k] > nt && abs(as.hx[k] - yx) + abs(as.hy[k] - yy) + abs(as.hz[k] - yz) <= m.merge_r) late = true;
        if (late) *sv.redo = 1;
    }
    if (lane == 0) {
        if (found >= 0) {
            atomicMin(&sv.t[found], nt);
            const int d0 = sv.slot[found] >> 28;
            if (depth > d0) sv.slot[found] = Y | (depth << 28);
        } else {
            const int k = atomicAdd(sv.n, 1);
            if (k >= sv.cap) { *sv.redo = 1; atomicSub(sv.n, 1); }          // shared arrays full: the scan is redone in serial mode
            else {
                sv.slot[k] = Y | (depth << 28);
                if (sv.keys) as_set_key(*sv.keys, k, m.skey[Y]);
                __threadfence();                                             // the entry becomes visible (t != T_INF) after its voxel (to the whole cluster)
                *(volatile int*)&sv.t[k] = nt;
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
__device__ void process_event(const DevMap& m, DevCtl* ctl, const SetView& sv, WarpScratch& ws, int j, unsigned scan_id) {
    const int lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & 7;
    const int A = sv.slot[j] & 0x0FFFFFFF, depth = sv.slot[j] >> 28, t = sv.t[j];
    const int nchg = merge_at_warp(m, ctl, A, t, scan_id, ws.chg);
    if (nchg > 0 && depth > sv.max_depth && lane == 0) *sv.redo = 1;
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
        else if (nt != T_INF) as_activate(m, ctl, sv, Yq, nt, dn);
        else if (kind == 1) {
            // the partner has no pair left that can pass (typically: it now shares A's group): retire its
            // pending entry right away instead of spending a round on a no-op.  Its entry is not being
            // processed concurrently: it lies within 1 of A and has a later event.
            const int n = *sv.n;
            for (int k = lane; k < n; k += 32) if ((sv.slot[k] & 0x0FFFFFFF) == Yq) sv.t[k] = T_INF;
            __syncwarp();
        }
    }
    if (lane == 0) { sv.t[j] = ntA; atomicAdd(&ctl->dbg[1], 1); }
    __syncwarp();
}

// ------------------------------------------------------------------------- fast path of one event: the "diamond"
// Everything one merge() call and its follow-up work can look at lies within Manhattan distance 2 of the event's voxel A: A, its
// six neighbours (what merge() reads and writes) and the eighteen voxels next to those (re-examined after a success).  The slow
// path above walks that neighbourhood as ~13 dependent look-ups (hash probe -> slot -> record, one after the other); here the 25
// cells are resolved by 25 lanes at once - one round trip for the hash probes, one for the records - into shared memory, and
// the whole event (decisions in the reference's order, blend, re-examination) then runs on that copy.  Cells hold the CURRENT
// incarnation of a voxel; if any cell has an older incarnation still visible in this scan (a ghost: evicted and re-created
// within the scan), the event takes the slow path, which walks those chains.
// could merge(P) succeed against Q after time `after`?  -> time bound (events with t > bound may succeed) or T_INF.
// (pair_static + pair_window of the slow path on the shared copies; no older incarnations here.)
__device__ __forceinline__ int cell_pair_bound(const DevMap& m, const CellRec& P, const CellRec& Q, int after, unsigned scan_id) {
    if (Q.slot < 0) return T_INF;
    const uint32_t fq = (uint32_t)(Q.w6 & 0xFFFFFFFFll);
    if ((fq & F_UE) || !(fq & F_PLANE)) return T_INF;
    if (Q.group == P.group) return T_INF;
    if (!plane_thresholds(m, v3(P.mean[0], P.mean[1], P.mean[2]), v3(P.nrm[0], P.nrm[1], P.nrm[2]), v3(Q.mean[0], Q.mean[1], Q.mean[2]),
                          v3(Q.nrm[0], Q.nrm[1], Q.nrm[2]))) return T_INF;
    const int born = (Q.born_scan == scan_id) ? Q.ft : -1;
    const int full = (Q.full_scan == scan_id) ? Q.full_idx : -1;
    const int s0 = born > full ? born : full;
    const int bound = s0 > after ? s0 : after;
    return (bound + 1 < Q.evict_t) ? bound : T_INF;
}
__device__ __forceinline__ int cell_event_floor(const CellRec& A, unsigned scan_id) { return (A.full_scan == scan_id) ? A.full_idx : -1; }

// merge() of the voxel in cell[0] at time t on the shared copies (same decisions, same arithmetic and the same global writes as
// merge_at_warp).  changed[0..n): the neighbour CELLS that were modified.
__device__ int merge_at_cells(const DevMap& m, DevCtl* ctl, CellRec* cell, int t, unsigned scan_id, int* changed, int ubase_l0) {
    const int lane = threadIdx.x & 31;
    const int ubase = __shfl_sync(0xffffffffu, ubase_l0, 0);                 // (reserved at the start of the event: long since back)
    const int A = cell[0].slot;
    double* ca = m.cov + (size_t)A * 36;
    double a0 = ca[lane], a1 = lane < 4 ? ca[32 + lane] : 0.0;
    const unsigned long long gA = cell[0].group;
    V3 mA = v3(cell[0].mean[0], cell[0].mean[1], cell[0].mean[2]), nA = v3(cell[0].nrm[0], cell[0].nrm[1], cell[0].nrm[2]);
    int B = -1;
    bool elig = false;
    if (lane < 6) {
        const CellRec& r = cell[1 + lane];
        if (r.slot >= 0) {
            const int born = (r.born_scan == scan_id) ? r.ft : -1;
            if (born <= t && t < r.evict_t) {                         // alive at time t
                B = r.slot;
                const uint32_t fl = (uint32_t)(r.w6 & 0xFFFFFFFFll);
                const bool closed = !(fl & F_UE) && (r.full_scan != scan_id || r.full_idx < t);
                elig = closed && (fl & F_PLANE) && r.group != gA;
            }
        }
    }
    const unsigned elmask = __ballot_sync(0xffffffffu, elig) & 0x3Fu;
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
    int nchg = 0;
#pragma unroll
    for (int d = 0; d < 6; d++) {
        if (!((elmask >> d) & 1u)) continue;
        const int Bd = __shfl_sync(0xffffffffu, B, d);
        CellRec& rb = cell[1 + d];
        const V3 mb = v3(rb.mean[0], rb.mean[1], rb.mean[2]), nb = v3(rb.nrm[0], rb.nrm[1], rb.nrm[2]);
        if (!plane_thresholds(m, mA, nA, mb, nb)) continue;
        double* cb = m.cov + (size_t)Bd * 36;
        const double b0 = b0p[d], b1 = b1p[d];
        const double tn0 = shfl_d(a0, 0) + shfl_d(a0, 7) + shfl_d(a0, 14);
        const double tm0 = shfl_d(a0, 21) + shfl_d(a0, 28) + shfl_d(a1, 3);
        const double tn1 = shfl_d(b0, 0) + shfl_d(b0, 7) + shfl_d(b0, 14);
        const double tm1 = shfl_d(b0, 21) + shfl_d(b0, 28) + shfl_d(b1, 3);
        const double tc0 = tn0 + tm0, tc1 = tn1 + tm1;
        V3 nm, nn;
        {       // Q9: operator precedence exactly as in voxel_map.cpp:166-167 (six divisions on six lanes, broadcast)
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
        // undo log: A before its first modification of this event (cell[0] still holds its original record), B before this one
        if (nchg == 0) undo_log(m, ubase, A, a0, a1, cell[0].mean, cell[0].nrm, cell[0].w6, gA);
        undo_log(m, ubase + 1 + nchg, Bd, b0, b1, rb.mean, rb.nrm, rb.w6, rb.group);
        const double c0 = (b0 * w0 + a0 * w1) / den;
        a0 = c0; cb[lane] = c0;
        if (lane < 4) { const double c1 = (b1 * w0 + a1 * w1) / den; a1 = c1; cb[32 + lane] = c1; }
        if (-dot(nm, nn) < 0.0) nn = neg(nn);
        mA = nm; nA = nn;
        const long long w6b = rb.w6 | (long long)F_MERGED;
        __syncwarp();
        if (lane < 3) {
            const double nml = sel3(nm, lane), nnl = sel3(nn, lane);
            m.hot[(size_t)Bd * 8 + lane] = nml; m.hot[(size_t)Bd * 8 + 3 + lane] = nnl;
            rb.mean[lane] = nml; rb.nrm[lane] = nnl;
        }
        if (lane == 0) {
            m.sgroup[Bd] = gA;
            rb.group = gA;
            rb.w6 = w6b;
            m.hot[(size_t)Bd * 8 + 6] = __longlong_as_double(w6b);
            atomicAdd((unsigned long long*)&ctl->st.n_merge, 1ull);
            changed[nchg] = 1 + d;
        }
        __syncwarp();
        nchg++;
    }
    if (nchg > 0) {
        ca[lane] = a0;
        if (lane < 4) ca[32 + lane] = a1;
        if (lane < 3) {
            const double mal = sel3(mA, lane), nal = sel3(nA, lane);
            m.hot[(size_t)A * 8 + lane] = mal; m.hot[(size_t)A * 8 + 3 + lane] = nal;
            cell[0].mean[lane] = mal; cell[0].nrm[lane] = nal;
        }
        if (lane == 0) {
            cell[0].w6 |= (long long)F_MERGED;
            m.hot[(size_t)A * 8 + 6] = __longlong_as_double(cell[0].w6);
        }
    }
    undo_close(m, ubase, nchg > 0 ? nchg + 1 : 0);
    __syncwarp();
    return nchg;
}

// returns false (nothing touched) when the event needs the slow path
__device__ bool process_event_fast(const DevMap& m, DevCtl* ctl, const SetView& sv, WarpScratch& ws, CellRec* cell, int j, unsigned scan_id) {
    const int lane = threadIdx.x & 31;
    const int A = sv.slot[j] & 0x0FFFFFFF, depth = sv.slot[j] >> 28, t = sv.t[j];
    // ---- the 25 cells, one per lane: two dependent round trips (hash probe, record)
    bool ghosts = false;
    int ubase_l0 = 0;
    if (lane < NCELL) {
        CellRec r;
        r.slot = -1; r.ghost = -1; r.cnt = 0; r.evn = 0; r.pad = 0; r.group = 0; r.w6 = 0; r.born_scan = 0; r.full_scan = SCAN_NEVER;
        r.ft = T_INF; r.full_idx = T_INF; r.evict_t = T_INF;
        for (int k = 0; k < 3; k++) { r.mean[k] = 0.0; r.nrm[k] = 0.0; }
        int X = -1;
        if (lane == 0) X = A;
        else {
            long long x, y, z;
            unpack_key(m.skey[A], x, y, z);
            int dx, dy, dz;
            cell_offset(lane, dx, dy, dz);
            x += dx; y += dy; z += dz;
            if (key_in_range(x) && key_in_range(y) && key_in_range(z)) X = hash_find(m, pack_key(x, y, z));
        }
        if (X >= 0) {
            const VoxRec v = load_rec(m, X);
            r.slot = X; r.cnt = m.cnt[X]; r.evn = m.evn[X];
            for (int k = 0; k < 3; k++) { r.mean[k] = v.mean[k]; r.nrm[k] = v.nrm[k]; }
            r.group = v.group; r.w6 = v.w6; r.born_scan = v.born_scan; r.full_scan = v.full_scan;
            r.ft = v.ft; r.full_idx = v.full_idx; r.evict_t = v.evict_t; r.ghost = v.ghost;
            ghosts = v.ghost >= 0;
        }
        cell[lane] = r;
    }
    if (__any_sync(0xffffffffu, ghosts)) return false;
    ubase_l0 = undo_reserve(ctl);
    __syncwarp();
    int* chg = ws.chg;                                                   // modified neighbour CELLS (1..6)
    const int nchg = merge_at_cells(m, ctl, cell, t, scan_id, chg, ubase_l0);
    if (nchg > 0 && depth > sv.max_depth && lane == 0) *sv.redo = 1;
    const int dn = depth + 1 > 7 ? 7 : depth + 1;
    // ---- follow-up work on the shared copies.  X ranges over A and the modified neighbours:
    //   wake(X):      earliest time one of X's own pairs can pass (again)             [A: always; a neighbour: if merge() is called for it in this scan]
    //   cand(X, d):   the voxel Y next to X in direction d: can the pair (Y, X) pass now?  (only for Y that is neither A nor modified)
    int nq = 0;
    const int nx = nchg > 0 ? nchg + 1 : 1;
    int* wmin = ws.res;                                                  // per X: earliest bound of its own pairs (reused below as result array)
    if (lane < 8) wmin[lane] = T_INF;
    __syncwarp();
    for (int e0 = 0; e0 < nx * 6; e0 += 32) {
        const int e = e0 + lane;
        if (e < nx * 6) {
            const int xi = e / 6, d = e % 6;
            const int cx = xi == 0 ? 0 : chg[xi - 1];
            const CellRec& X = cell[cx];
            if (xi == 0 || (X.cnt != 0 && X.evn != 0)) {
                int after = t;
                if (xi > 0) { const int fl = cell_event_floor(X, scan_id); after = t > fl ? t : fl; }
                const int cy = cell_neighbour(cx, d);
                const int w = cy >= 0 ? cell_pair_bound(m, X, cell[cy], after, scan_id) : T_INF;
                if (w != T_INF) atomicMin(&wmin[xi], w);
            }
        }
    }
    __syncwarp();
    // queue: the wakes first (kind 0: A, kind 1: a modified neighbour with merge() calls in this scan) ...
    for (int xi = 0; xi < nx; xi++) {                                     // warp-uniform, nx <= 7
        const int cx = xi == 0 ? 0 : chg[xi - 1];
        if (!(xi == 0 || (cell[cx].cnt != 0 && cell[cx].evn != 0))) continue;
        if (lane == 0) { ws.tgt[nq] = cell[cx].slot; ws.aft[nq] = wmin[xi]; ws.kind[nq] = (unsigned char)(xi == 0 ? 0 : 1); }
        nq++;
    }
    __syncwarp();
    // ... then the candidates
    if (nchg > 0) {
        for (int e0 = 0; e0 < nx * 6; e0 += 32) {
            const int e = e0 + lane;
            int Y = -1, w = T_INF;
            if (e < nx * 6) {
                const int xi = e / 6, d = e % 6;
                const int cx = xi == 0 ? 0 : chg[xi - 1];
                const int cy = cell_neighbour(cx, d);
                bool skip = cy <= 0;                                        // A itself (or outside the diamond: cannot happen for cx <= 6)
                for (int q = 0; q < nchg && !skip; q++) if (chg[q] == cy) skip = true;      // handled as a wake
                if (!skip) {
                    const CellRec& Yc = cell[cy];
                    if (Yc.slot >= 0 && Yc.cnt != 0 && Yc.evn != 0) {     // merge() is called for Y in this scan
                        const int fl = cell_event_floor(Yc, scan_id);
                        w = cell_pair_bound(m, Yc, cell[cx], t > fl ? t : fl, scan_id);
                        if (w != T_INF) Y = Yc.slot;
                    }
                }
            }
            const bool q_me = Y >= 0;
            const unsigned bal = __ballot_sync(0xffffffffu, q_me);
            if (q_me) {
                const int pos = nq + __popc(bal & ((1u << lane) - 1));
                if (pos < FOLLOW_CAP) { ws.tgt[pos] = Y; ws.aft[pos] = w; ws.kind[pos] = 2; }
            }
            nq += __popc(bal);
        }
    }
    if (nq > FOLLOW_CAP) { if (lane == 0) atomicOr(&ctl->err, E_QUEUE); nq = FOLLOW_CAP; }
    __syncwarp();
    // first point index of each queued voxel after its bound: the whole warp scans the voxel's segment (coalesced)
    for (int q = 0; q < nq; q++) {
        const int Yq = ws.tgt[q], aq = ws.aft[q];
        int best = T_INF;
        if (aq != T_INF) {
            const int c = m.cnt[Yq], off = m.seg_off[Yq];
            for (int k = lane; k < c; k += 32) { const int i = m.seg[off + k]; if (i > aq && i < best) best = i; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const int y = __shfl_xor_sync(0xffffffffu, best, o); best = y < best ? y : best; }
        }
        __syncwarp();
        if (lane == 0) ws.res[q] = best;
    }
    __syncwarp();
    // apply: activations / retirements of the neighbourhood, then A's own next event (as in the slow path)
    int ntA = T_INF;
    for (int q = 0; q < nq; q++) {
        const int kind = ws.kind[q], Yq = ws.tgt[q], nt = ws.res[q];
        if (kind == 0) ntA = nt;
        else if (nt != T_INF) as_activate(m, ctl, sv, Yq, nt, dn);
        else if (kind == 1) {
            const int n = *sv.n;
            for (int k = lane; k < n; k += 32) if ((sv.slot[k] & 0x0FFFFFFF) == Yq) sv.t[k] = T_INF;
            __syncwarp();
        }
    }
    if (lane == 0) { sv.t[j] = ntA; atomicAdd(&ctl->dbg[1], 1); }
    __syncwarp();
    return true;
}

struct MergeShared {                    // dynamic shared memory of k_merge_rounds (opt-in size, map_configure_kernels)
    ActiveSet as;
    WarpScratch wsc[16];
    CellRec cells[16][NCELL];
    int s_cnt, s_nready, s_redo, s_n;
    unsigned long long s_best[16];
};

// The merge() calls of one scan.
//   parallel rounds (the normal case): the active set lives in shared memory; every round executes the events that have no
//     earlier pending event within MERGE_R, one per warp (footprint argument above: exact as long as no activation comes too late
//     and the set and the history fit).  Every modification is recorded in an undo log.
//   exact serial mode: if one of those conditions fails (forced by the tests through VMP_MERGE_MAX_DEPTH=-1, which declares every
//     successful merge a failure) the modifications of the parallel rounds are taken back from the undo log and the scan's events are
//     executed again one at a time in strict event order - the reference's own order, restricted to the events that can do
//     something - on the global arrays the prefilter filled (no capacity limit).  Scans whose active set does not fit the shared
//     arrays start in that mode.
// Launched as ONE thread-block cluster of MERGE_CLUSTER CTAs (16 warps each): the active set, the history and the round counters
// live in the shared memory of CTA rank 0 and are reached by the other CTAs through distributed shared memory; every CTA has its own
// per-warp scratch (WarpScratch, the 25-cell copies).  A round's ready events are dealt to the 16 x MERGE_CLUSTER warps of the
// cluster, so that a round is one event time long instead of ceil(ready / 16) (the first round of a C2 scan has ~40 ready events).
// Rank 0 alone does the bookkeeping between the rounds (readiness, compaction) and, if it comes to that, the serial redo.
constexpr int MERGE_CLUSTER = 4;

// The events write voxels (global memory) that events on the other SMs of the cluster read in later rounds.  cluster.sync() is
// barrier.cluster.arrive.release + wait.acquire: in SASS MEMBAR.ALL.GPU in front of UCGABAR_ARV and CCTL.IVALL (L1 invalidate) behind
// UCGABAR_WAIT (profiles/sass_r02_k_merge_rounds.txt) - a __threadfence() on either side of it only repeated that.
__device__ __forceinline__ void merge_cluster_sync(cg::cluster_group& cluster) { cluster.sync(); }

__global__ void __launch_bounds__(512) k_merge_rounds(DevMap m, DevCtl* ctl) {
    extern __shared__ __align__(16) unsigned char merge_smem[];
    MergeShared& S = *reinterpret_cast<MergeShared*>(merge_smem);
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), nrank = (int)cluster.num_blocks();
    MergeShared& S0 = *cluster.map_shared_rank(&S, 0);
    ActiveSet& as = S0.as;                           // rank 0's, for everybody
    WarpScratch* wsc = S.wsc;                        // this CTA's
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n0 = ctl->n_hot;                       // (voxel, first relevant event) pairs prepared by k_merge_prefilter
    if (n0 == 0) return;
    const unsigned scan_id = ctl->scan_id;
    bool serial = n0 > m.merge_cap;
    if (serial && rank != 0) return;
    if (rank == 0 && tid == 0) { S.s_redo = 0; ctl->dbg[0] = n0; }
    if (!serial) {
        if (rank == 0) {
            if (tid == 0) {
                long long x, y, z;
                unpack_key(m.skey[m.act_slot[0]], x, y, z);
                S.as.ox = (int)x; S.as.oy = (int)y; S.as.oz = (int)z;
                S.as.n = n0; S.as.hn = 0; S.s_nready = 0;
            }
            __syncthreads();
            for (int k = tid; k < n0; k += blockDim.x) {
                const int A = m.act_slot[k];
                S.as.slot[k] = A;
                S.as.t[k] = m.act_t[k];
                as_set_key(S.as, k, m.skey[A]);
            }
            // positions outside the live range always read as retired: a warp that scans the set while another one appends (n already
            // counted, entry not yet written) must not match whatever an earlier entry left there
            for (int k = n0 + tid; k < MERGE_CAP; k += blockDim.x) S.as.t[k] = T_INF;
            __syncthreads();
        }
        // (everything the events will look at was pulled into the L2 by k_merge_prefilter, 25 cells around every active voxel)
        SetView sv;
        sv.slot = as.slot; sv.t = as.t; sv.n = &as.n; sv.cap = m.merge_cap; sv.keys = &as; sv.redo = &S0.s_redo; sv.max_depth = m.merge_max_depth;
        // Two cluster barriers per round.  Rank 0 prepares the round (compaction of the previous one, readiness) while the other CTAs
        // wait at barrier A; everybody executes events up to barrier B.  The shared counters are written by rank 0 between B and the
        // next A only, and read by the others between A and B only.
        const int merge_r = m.merge_r;
        for (int round = 0; round < 100000; round++) {
            if (rank == 0) {
                if (round == 0) { if (tid == 0) S.s_cnt = n0; }
                else if (wid == 0) {
                    // compaction of retired entries (one warp), keeps the rest in place order
                    const int nn = S.as.n;
                    int w = 0;
                    for (int base = 0; base < nn; base += 32) {
                        const int k = base + lane;
                        const bool live = k < nn && S.as.t[k] != T_INF;
                        const int sl = live ? S.as.slot[k] : 0, tt = live ? S.as.t[k] : 0;
                        const short kx = live ? S.as.kx[k] : 0, ky = live ? S.as.ky[k] : 0, kz = live ? S.as.kz[k] : 0;
                        const unsigned bal = __ballot_sync(0xffffffffu, live);
                        const int pos = w + __popc(bal & ((1u << lane) - 1));
                        __syncwarp();
                        if (live) { S.as.slot[pos] = sl; S.as.t[pos] = tt; S.as.kx[pos] = kx; S.as.ky[pos] = ky; S.as.kz[pos] = kz; }
                        w += __popc(bal);
                        __syncwarp();
                    }
                    for (int k = w + lane; k < nn; k += 32) S.as.t[k] = T_INF;     // (the moved entries' old copies)
                    if (lane == 0) { S.as.n = w; S.s_cnt = w; S.s_nready = 0; atomicAdd(&ctl->dbg[2], 65536); }      // (rounds in the upper half of dbg[2])
                }
                __syncthreads();
                const int n = S.as.n;
                // readiness (one warp per entry, lanes over the others): no other entry with an earlier event within MERGE_R
                for (int j = wid; j < n; j += 16) {
                    const int tj = S.as.t[j];
                    if (tj == T_INF) continue;
                    const int x = S.as.kx[j], y = S.as.ky[j], z = S.as.kz[j];
                    bool conflict = false;
                    for (int i = lane; i < n; i += 32)
                        if (S.as.t[i] < tj && abs(S.as.kx[i] - x) + abs(S.as.ky[i] - y) + abs(S.as.kz[i] - z) <= merge_r) conflict = true;
                    if (!__any_sync(0xffffffffu, conflict) && lane == 0) {
                        S.as.rlist[atomicAdd(&S.s_nready, 1)] = (short)j;
                        const int h = atomicAdd(&S.as.hn, 1);                     // started: remembered for the activation check
                        if (h < MERGE_HIST) { S.as.ht[h] = tj; S.as.hx[h] = (short)x; S.as.hy[h] = (short)y; S.as.hz[h] = (short)z; }
                        else S.s_redo = 1;
                    }
                }
            }
            cluster.sync();                                                   // A (shared memory only)
            if (S0.s_cnt == 0) break;
            const int nr = S0.s_nready;
            for (int q = rank * 16 + wid; q < nr; q += 16 * nrank) {           // one ready event per warp of the cluster
                const int j = as.rlist[q];
                if (!process_event_fast(m, ctl, sv, wsc[wid], S.cells[wid], j, scan_id)) process_event(m, ctl, sv, wsc[wid], j, scan_id);
            }
            merge_cluster_sync(cluster);                                      // B (the events' global writes)
            if (S0.s_redo) break;
        }
        const int redo = S0.s_redo;
        cluster.sync();                              // nobody leaves while another CTA may still be reading rank 0's shared memory
        if (rank != 0 || !redo) return;
        // ---- take the parallel rounds back (undo log in reverse order: the oldest record of a voxel is restored last)
        __threadfence();
        __syncthreads();
        const int nu = ctl->n_undo;
        if (nu > m.undo_cap) { if (tid == 0) atomicOr(&ctl->err, E_MERGE_DEPTH); return; }      // log overflowed: cannot be redone exactly, reported
        if (wid == 0) {
            for (int e = nu - 1; e >= 0; e--) {
                const int slot = m.undo_slot[e];
                if (slot < 0) continue;                                     // (reserved by an event, not used)
                const double* r = m.undo_rec + (size_t)e * UNDO_W;
                m.cov[(size_t)slot * 36 + lane] = r[lane];
                if (lane < 4) m.cov[(size_t)slot * 36 + 32 + lane] = r[32 + lane];
                if (lane < 6) m.hot[(size_t)slot * 8 + lane] = r[36 + lane];
                if (lane == 0) { m.hot[(size_t)slot * 8 + 6] = r[42]; m.sgroup[slot] = (unsigned long long)__double_as_longlong(r[43]); }
                __syncwarp();
            }
            if (lane == 0) { ctl->st.n_merge = 0; ctl->dbg[1] = 0; ctl->dbg[2] = 1 << 30; }     // (dbg[2] bit 30: the scan was redone serially)
        }
        __threadfence();
        __syncthreads();
        serial = true;
    }
    // ---- exact serial mode on the prefilter's global arrays
    if (tid == 0) S.s_n = n0;
    __syncthreads();
    SetView sv;
    sv.slot = m.act_slot; sv.t = m.act_t; sv.n = &S.s_n; sv.cap = m.nmax; sv.keys = nullptr; sv.redo = &S.s_redo; sv.max_depth = INT_MAX;
    for (int step = 0; step < 100000000; step++) {
        const int n = S.s_n;
        unsigned long long best = ~0ull;                                     // (time << 32 | entry): the earliest pending event
        for (int k = tid; k < n; k += blockDim.x) {
            const int t = m.act_t[k];
            if (t != T_INF) { const unsigned long long v = ((unsigned long long)(unsigned)t << 32) | (unsigned)k; best = v < best ? v : best; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long y = __shfl_xor_sync(0xffffffffu, best, o); best = y < best ? y : best; }
        if (lane == 0) S.s_best[wid] = best;
        __syncthreads();
        best = S.s_best[0];
        for (int w = 1; w < 16; w++) best = S.s_best[w] < best ? S.s_best[w] : best;
        if (best == ~0ull) break;
        const int j = (int)(best & 0xFFFFFFFFull);
        if (wid == 0) {
            if (!process_event_fast(m, ctl, sv, wsc[0], S.cells[0], j, scan_id)) process_event(m, ctl, sv, wsc[0], j, scan_id);
        }
        __threadfence();
        __syncthreads();
    }
}
