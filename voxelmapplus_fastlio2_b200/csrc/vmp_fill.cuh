// vmp_fill.cuh — VoxelGrid::pushPoint / addToPlane / updatePlane (voxel_map.cpp:29-136) and the per-voxel
// part of VoxelMap::build (voxel_map.cpp:200-230).  Included by vmp_map.cu inside namespace vmp.
//
// The pushPoint state machine is strictly sequential per voxel, but its CONTROL FLOW never depends on
// the result of a refit (is_plane only selects between two identical branches while the voxel is
// filling), so the fill is split into three kernels:
//   k_fill_state   one warp per touched voxel: select the (at most max_point_thresh) points the voxel
//                  consumes, in point order; running mean / sum p p^T; append to the stored points;
//                  emit one REFIT JOB (snapshot of n, mean, sum pp^T, number of stored points) wherever
//                  the reference calls updatePlane() past its early return
//   k_fill_refit   one warp per (job, 32 stored points): 3x3 symmetric eigen solve (restated
//                  SelfAdjointEigenSolver), is_plane test, and the 6x6 contributions J Sigma J^T of its
//                  points, staged in HBM
//   k_fill_acc     one warp per voxel with jobs: plane->cov += contributions, jobs in refit order and
//                  points in stored order (Q7: accumulates, never reset), final normal / centre / flags
// All refits of a scan therefore run concurrently (a voxel that goes from empty to full in one scan has
// 10 of them over 550 points), and the summation order is exactly the reference's.
#pragma once

constexpr int SEL_MAX = 256;            // points one voxel can consume per scan on the fast path (>= max_point_thresh)

// ascending in-place sort of a[0..c) by one warp (distinct values); a may be shared or global
__device__ void warp_sort(int* a, int c) {
    const int lane = threadIdx.x & 31;
    if (c <= 1) return;
    if (c <= 32) {
        const int v = lane < c ? a[lane] : INT_MAX;
        int rank = 0;
#pragma unroll
        for (int l = 0; l < 32; l++) { const int o = __shfl_sync(0xffffffffu, v, l); rank += (o < v) ? 1 : 0; }
        __syncwarp();
        if (lane < c) a[rank] = v;
        __syncwarp();
        return;
    }
    int npow = 64;
    while (npow < c) npow <<= 1;
    const int half = npow >> 1;
    for (int k = 2; k <= npow; k <<= 1) {
        const int hk = k >> 1;
        for (int t = lane; t < half; t += 32) {                 // flip stage: all comparators ascending
            const int blk = t / hk, o = t % hk;
            const int lo = blk * k + o, hi = blk * k + k - 1 - o;
            if (hi < c) { const int x = a[lo], y = a[hi]; if (x > y) { a[lo] = y; a[hi] = x; } }
        }
        __syncwarp();
        for (int j = k >> 2; j >= 1; j >>= 1) {
            for (int t = lane; t < half; t += 32) {
                const int lo = (t / j) * 2 * j + (t % j), hi = lo + j;
                if (hi < c) { const int x = a[lo], y = a[hi]; if (x > y) { a[lo] = y; a[hi] = x; } }
            }
            __syncwarp();
        }
    }
}

// the K smallest of seg[0..c) (distinct point indices in [0, n)), ascending, into sel[0..K) (shared)
__device__ void warp_select_sorted(const int* seg, int c, int K, int n, int* sel) {
    const int lane = threadIdx.x & 31;
    if (K >= c || c <= 32) {                                     // small segment: sort all of it, the caller takes the first K
        for (int q = lane; q < c; q += 32) sel[q] = seg[q];
        __syncwarp();
        warp_sort(sel, c);
        return;
    }
    // K-th smallest value by bisection on the value range (counts are warp reductions over the segment)
    int lo = 0, hi = n - 1;                                      // invariant: count(<= hi) >= K, count(<= lo - 1) < K
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        int cntl = 0;
        for (int q = lane; q < c; q += 32) cntl += (seg[q] <= mid) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cntl += __shfl_xor_sync(0xffffffffu, cntl, o);
        if (cntl >= K) hi = mid; else lo = mid + 1;
    }
    int w = 0;
    for (int base = 0; base < c; base += 32) {
        const int q = base + lane;
        const int v = q < c ? seg[q] : INT_MAX;
        const bool take = v <= lo;
        const unsigned bal = __ballot_sync(0xffffffffu, take);
        if (take) sel[w + __popc(bal & ((1u << lane) - 1))] = v;
        w += __popc(bal);
    }
    __syncwarp();
    warp_sort(sel, K);
}

// J Sigma J^T of one stored point (voxel_map.cpp:115-129)
__device__ __forceinline__ void plane_contrib(const V3& p, const M3& S, const V3& mean, int n, const double* evals,
                                              const M3& evecs, const V3& nrm, double* out /*36*/) {
    M3 F = zeros<3, 3>();
#pragma unroll
    for (int mm = 1; mm < 3; mm++) {
        const V3 vm = v3(evecs(0, mm), evecs(1, mm), evecs(2, mm));
        const Mat<1, 3> lhs = divs(tr(sub(p, mean)), n * (evals[0] - evals[mm]));
        const M3 Sm = add(outer(vm, nrm), outer(nrm, vm));
        const Mat<1, 3> Fm = mul(lhs, Sm);
        F(mm, 0) = Fm[0]; F(mm, 1) = Fm[1]; F(mm, 2) = Fm[2];
    }
    Mat<6, 3> J;
    set_block(J, 0, 0, mul(evecs, F));
    set_block(J, 3, 0, divs(eye<3>(), (double)n));
    const Mat<6, 6> C = mul(mul(J, S), tr(J));
#pragma unroll
    for (int k = 0; k < 36; k++) out[k] = C.a[k];
}

constexpr int SNAP_MAX = 40;            // refit snapshots buffered per voxel before they are emitted as jobs
constexpr int SNAP_W = 12;              // doubles per snapshot: n, nt, mean[3], ppt[6], pad

// Emit the buffered snapshots of one voxel as refit jobs: ONE atomic per counter for the whole group
// (whole warp calls).  Returns the id of the first job (-1 on overflow); chains them after prev_job.
__device__ int emit_jobs(const DevMap& m, DevCtl* ctl, int slot, const double* snap, int nj, int src_off, int prev_job) {
    const int lane = threadIdx.x & 31;
    int j0 = 0, b0 = 0;
    long long off0 = 0;
    int tot_nt = 0, tot_nb = 0;
    for (int k = 0; k < nj; k++) { const int nt = (int)snap[k * SNAP_W + 1]; tot_nt += nt; tot_nb += (nt + 31) >> 5; }
    if (lane == 0) {
        j0 = atomicAdd(&ctl->n_jobs, nj);
        off0 = (long long)atomicAdd(&ctl->contrib_top, (unsigned long long)tot_nt);
        b0 = atomicAdd(&ctl->n_batches, tot_nb);
        if (j0 + nj > m.job_cap || off0 + tot_nt > m.contrib_cap || b0 + tot_nb > m.bat_cap) { atomicOr(&ctl->err, E_FILL_CAP); j0 = -1; }
        else if (prev_job >= 0) m.job_next[prev_job] = j0;
    }
    j0 = __shfl_sync(0xffffffffu, j0, 0);
    b0 = __shfl_sync(0xffffffffu, b0, 0);
    off0 = __shfl_sync(0xffffffffu, off0, 0);
    if (j0 < 0) return -1;
    int acc_nt = 0, acc_nb = 0;
    for (int k = 0; k < nj; k++) {                                    // nj is small (<= SNAP_MAX), warp-uniform loop
        const double* sn = snap + k * SNAP_W;
        const int nt = (int)sn[1], nb = (nt + 31) >> 5, j = j0 + k;
        if (lane == 0) {
            m.job_slot[j] = slot; m.job_n[j] = (int)sn[0]; m.job_nt[j] = nt; m.job_off[j] = off0 + acc_nt; m.job_src[j] = src_off;
            m.job_next[j] = (k + 1 < nj) ? j + 1 : -1; m.job_plane[j] = 0;
        }
        if (lane < 3) m.job_mean[3 * (size_t)j + lane] = sn[2 + lane];
        if (lane < 6) m.job_ppt[6 * (size_t)j + lane] = sn[5 + lane];
        for (int b = lane; b < nb; b += 32) { m.bat_job[b0 + acc_nb + b] = j; m.bat_idx[b0 + acc_nb + b] = b; }
        acc_nt += nt; acc_nb += nb;
    }
    __syncwarp();                                                     // the caller may overwrite the snapshots now
    return j0;
}

__global__ void __launch_bounds__(128) k_fill_state(DevMap m, DevScan s, DevCtl* ctl, int build) {
    __shared__ int sel_all[4][SEL_MAX];
    __shared__ double spt_all[4][SEL_MAX * 3];
    __shared__ double snap_all[4][SNAP_MAX * SNAP_W];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int* sel = sel_all[wib];
    double* spt = spt_all[wib];
    double* snap = snap_all[wib];
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
    const int V = ctl->n_touched, npts = ctl->n;
    const unsigned scan_id = ctl->scan_id;
    long long c_ins = 0, c_full = 0, c_probe = 0, c_pvox = 0;
    for (int vi = wg; vi < V; vi += nW) {
        const int slot = m.touched[vi];
        const int c = m.cnt[slot], off = m.seg_off[slot];
        uint32_t flags; int n;
        hot_get_fn(m.hot, slot, flags, n);
        int events = 0, first_job = -1;
        if (!(flags & F_UE) && !build) {
            events = c;                                     // full before this scan: merge() or nothing per point
        } else {
            int nt = m.n_temp[slot], nw = m.newly[slot];
            const int nt0 = nt;
            // addToPlane (voxel_map.cpp:29-34) is a dependent chain per point; its nine scalar updates run on nine
            // lanes' worth of registers instead of one lane's: lane l < 3 owns mean[l], lane l < 6 owns ppt[l]
            // (xx, yx, yy, zx, zy, zz).  A lone warp pays per instruction, and the three divisions become one.
            const int cm = lane < 3 ? lane : 0;
            const int ia = lane == 0 ? 0 : lane < 3 ? 1 : lane < 6 ? 2 : 0;
            const int ib = (lane == 2 || lane == 4) ? 1 : lane == 5 ? 2 : 0;
            double mean_l = m.hot[(size_t)slot * 8 + cm];
            double ppt_l = m.ppt[(size_t)slot * 6 + (lane < 6 ? lane : 0)];
            unsigned full_scan = SCAN_NEVER; int full_idx = T_INF;
            double* tp = m.tp + (size_t)slot * 12 * m.maxpt;
            int consumed = 0;
            if (build) {
                // build() has no cap (Q18): every point in order, one updatePlane() at the end.  Runs once per map.
                int* order = m.seg + off;
                warp_sort(order, c);
                for (int j = 0; j < c; j++) {
                    const size_t i3 = 3 * (size_t)order[j];
                    const double pm = s.pw[i3 + cm], pa = s.pw[i3 + ia], pb = s.pw[i3 + ib];
                    mean_l = mean_l + (pm - mean_l) / (n + 1.0);
                    ppt_l += pa * pb;
                    n += 1;
                }
                for (int q = lane; q < c && nt0 + q < m.maxpt; q += 32) {
                    const int i = order[q];
#pragma unroll
                    for (int k = 0; k < 3; k++) tp[(size_t)k * m.maxpt + nt0 + q] = s.pw[3 * (size_t)i + k];
#pragma unroll
                    for (int k = 0; k < 9; k++) tp[(size_t)(3 + k) * m.maxpt + nt0 + q] = s.pcov[9 * (size_t)i + k];
                }
                nt += c; consumed = c;
                if (n >= m.upt) {
                    flags |= F_INIT;
                    if (lane == 0) { snap[0] = n; snap[1] = nt; }
                    if (lane < 3) snap[2 + lane] = mean_l;
                    if (lane < 6) snap[5 + lane] = ppt_l;
                    __syncwarp();
                    first_job = emit_jobs(m, ctl, slot, snap, 1, off, -1);
                }
            } else {
                // the voxel consumes at most K points before it closes (never more than its free room, at least one)
                const int room = m.maxpt - nt;
                int K = room < 1 ? 1 : room;
                if (K > c) K = c;
                warp_select_sorted(m.seg + off, c, K, npts, sel);              // K <= max_point_thresh <= SEL_MAX
                for (int q = lane; q < K; q += 32) {                           // gather the K points once, in parallel
                    const int i = sel[q];
                    spt[3 * q] = s.pw[3 * (size_t)i]; spt[3 * q + 1] = s.pw[3 * (size_t)i + 1]; spt[3 * q + 2] = s.pw[3 * (size_t)i + 2];
                }
                __syncwarp();
                // pushPoint's control flow (voxel_map.cpp:42-95) depends on the counters only, never on the points: which steps
                // refit, which step closes the voxel and the final counters have a closed form in (n, n_temp, newly_add_point, K)
                // (checked against the step-by-step state machine on 6.2e6 parameter combinations).  The per-point loop below is
                // left with the two recurrences that ARE sequential (running mean, sum p p^T) and the refit snapshots.
                const bool init0 = (flags & F_INIT) != 0;
                const int n0 = n, nw0 = nw;
                const int jA = init0 ? -1 : (m.upt - 1 - n0 > 0 ? m.upt - 1 - n0 : 0);   // first updatePlane() past the early return (voxel not yet initialised)
                const int j_init = init0 ? 0 : jA + 1;                                   // first step that sees is_init == true
                int jc = m.maxpt - nt0 - 1;                                              // closing step: is_init before it, temp_points.size() >= max_point_thresh after it
                if (jc < j_init) jc = j_init;
                const bool closes = jc <= K - 1;
                consumed = closes ? jc + 1 : K;
                int nsnap = 0, prev_job = -1;
                int next_refit = init0 ? m.upt - nw0 - 1 : jA;
                for (int j = 0; j < consumed; j++) {                           // point order
                    const double pm = spt[3 * j + cm], pa = spt[3 * j + ia], pb = spt[3 * j + ib];
                    mean_l = mean_l + (pm - mean_l) / (double)(n0 + j + 1);
                    ppt_l += pa * pb;
                    if (j == next_refit) {
                        next_refit = (!init0 && j == jA) ? j_init + (m.upt - nw0) - 1 : next_refit + m.upt;
                        if (nsnap == SNAP_MAX) {
                            __syncwarp();
                            const int jb = emit_jobs(m, ctl, slot, snap, nsnap, -1, prev_job);
                            if (first_job < 0) first_job = jb;
                            prev_job = jb < 0 ? -1 : jb + nsnap - 1;
                            nsnap = 0;
                        }
                        double* sn = snap + nsnap * SNAP_W;
                        if (lane == 0) { sn[0] = n0 + j + 1; sn[1] = nt0 + j + 1; }
                        if (lane < 3) sn[2 + lane] = mean_l;
                        if (lane < 6) sn[5 + lane] = ppt_l;
                        nsnap++;
                    }
                }
                n = n0 + consumed;
                nt = closes ? 0 : nt0 + consumed;                              // closing frees temp_points
                { const int sdone = consumed - j_init; nw = (nw0 + (sdone > 0 ? sdone : 0)) % m.upt; }
                if (!init0 && consumed - 1 >= jA) flags |= F_INIT;
                if (closes) { flags &= ~F_UE; full_scan = scan_id; full_idx = sel[jc]; }
                if (!closes && K < c && lane == 0) atomicOr(&ctl->err, E_QUEUE);          // cannot happen: K points always close the voxel
                __syncwarp();
                if (nsnap > 0) {
                    const int jb = emit_jobs(m, ctl, slot, snap, nsnap, -1, prev_job);
                    if (first_job < 0) first_job = jb;
                }
                // store the consumed points (xyz + cov) behind the nt0 already stored ones, one point per lane
                for (int q = lane; q < consumed && nt0 + q < m.maxpt; q += 32) {
                    const int i = sel[q];
#pragma unroll
                    for (int k = 0; k < 3; k++) tp[(size_t)k * m.maxpt + nt0 + q] = spt[3 * q + k];
#pragma unroll
                    for (int k = 0; k < 9; k++) tp[(size_t)(3 + k) * m.maxpt + nt0 + q] = s.pcov[9 * (size_t)i + k];
                }
            }
            c_ins += consumed;
            events = c - consumed;
            __syncwarp();
            if (lane < 3) m.hot[(size_t)slot * 8 + lane] = mean_l;
            if (lane < 6) m.ppt[(size_t)slot * 6 + lane] = ppt_l;
            if (lane == 0) {
                hot_set_fn(m.hot, slot, flags, n);
                m.n_temp[slot] = nt; m.newly[slot] = nw;
                if (full_scan != SCAN_NEVER) { m.full_scan[slot] = full_scan; m.full_idx[slot] = full_idx; }
            }
        }
        c_full += events;
        if (first_job < 0) {
            // no refit in this scan: is_plane is final, settle the merge() bookkeeping here
            if (!(flags & F_UE) && (flags & F_PLANE)) { c_probe += events; c_pvox += events > 0 ? 1 : 0; } else events = 0;
        }
        if (lane == 0) { m.evn[slot] = events; m.vox_job[vi] = first_job; }
        __syncwarp();
    }
    if (lane == 0) {
        if (c_ins) atomicAdd((unsigned long long*)&ctl->st.n_ins, (unsigned long long)c_ins);
        if (c_full) atomicAdd((unsigned long long*)&ctl->st.n_full, (unsigned long long)c_full);
        if (c_probe) atomicAdd((unsigned long long*)&ctl->st.n_mergeprobe, (unsigned long long)c_probe);
        if (c_pvox) atomicAdd((unsigned long long*)&ctl->st.n_mergevox, (unsigned long long)c_pvox);
    }
}

// updatePlane() body (voxel_map.cpp:100-135) for one batch of 32 stored points of one job
__global__ void __launch_bounds__(128) k_fill_refit(DevMap m, DevScan s, DevCtl* ctl) {
    constexpr int TILE_LD = 37;
    __shared__ double tile_all[4][32 * TILE_LD];
    double* tile = tile_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
    const int NB = ctl->n_batches < m.bat_cap ? ctl->n_batches : m.bat_cap;
    for (int b = wg; b < NB; b += nW) {
        const int j = m.bat_job[b], bi = m.bat_idx[b];
        const int slot = m.job_slot[j], n = m.job_n[j], nt = m.job_nt[j], src = m.job_src[j];
        const V3 mean = v3(m.job_mean[3 * (size_t)j], m.job_mean[3 * (size_t)j + 1], m.job_mean[3 * (size_t)j + 2]);
        const double* pp = m.job_ppt + 6 * (size_t)j;
        const double nd = (double)n;
        const double c00 = pp[0] / nd - mean[0] * mean[0];
        const double c10 = pp[1] / nd - mean[1] * mean[0];
        const double c11 = pp[2] / nd - mean[1] * mean[1];
        const double c20 = pp[3] / nd - mean[2] * mean[0];
        const double c21 = pp[4] / nd - mean[2] * mean[1];
        const double c22 = pp[5] / nd - mean[2] * mean[2];
        double evals[3];
        M3 evecs;
        eig3_sym(c00, c10, c11, c20, c21, c22, evals, evecs);
        const bool plane = !(evals[0] > m.plane_thresh);                  // Q13: otherwise norm / cov stay
        V3 nrm = v3(evecs(0, 0), evecs(1, 0), evecs(2, 0));
        if (bi == 0 && lane == 0) {
            m.job_plane[j] = plane ? 1 : 0;
            if (plane) {
                V3 ns = nrm;
                if (-dot(mean, nrm) < 0.0) ns = neg(nrm);
                for (int k = 0; k < 3; k++) m.job_norm[3 * (size_t)j + k] = ns[k];
                atomicAdd((unsigned long long*)&ctl->st.refit_points, (unsigned long long)nt);
                if (src < 0 && nt > m.maxpt) atomicOr(&ctl->err, E_REFIT_OVERFLOW);
            }
        }
        if (!plane) continue;
        const int q = bi * 32 + lane;
        if (q < nt) {
            V3 p; M3 S;
            if (src >= 0) {                                                // build(): through the sorted segment
                const int i = m.seg[src + q];
                p = v3(s.pw[3 * (size_t)i], s.pw[3 * (size_t)i + 1], s.pw[3 * (size_t)i + 2]);
#pragma unroll
                for (int k = 0; k < 9; k++) S.a[k] = s.pcov[9 * (size_t)i + k];
            } else {
                const double* tp = m.tp + (size_t)slot * 12 * m.maxpt;
                const int qq = q < m.maxpt ? q : m.maxpt - 1;
                p = v3(tp[qq], tp[m.maxpt + qq], tp[2 * m.maxpt + qq]);
#pragma unroll
                for (int k = 0; k < 9; k++) S.a[k] = tp[(size_t)(3 + k) * m.maxpt + qq];
            }
            double out[36];
            plane_contrib(p, S, mean, n, evals, evecs, nrm, out);
#pragma unroll
            for (int k = 0; k < 36; k++) tile[lane * TILE_LD + k] = out[k];
        }
        // the batch's contributions are one contiguous block of <= 32 x 36 doubles: written cooperatively, fully
        // coalesced (one 288-byte record per lane would touch 36 separate sectors per store instruction)
        __syncwarp();
        const int np = (nt - bi * 32) < 32 ? (nt - bi * 32) : 32;
        double* dst = m.contrib + ((size_t)m.job_off[j] + (size_t)bi * 32) * 36;
        for (int e = lane; e < np * 36; e += 32) dst[e] = tile[(e / 36) * TILE_LD + e % 36];
        __syncwarp();
    }
}

// plane->cov += J Sigma J^T, jobs in refit order, points in stored order; final normal / centre / is_plane.
// The adds of one entry form one dependent chain (that IS the reference's order); the loads do not: the
// contributions are streamed through shared memory in chunks of 32 points, the next chunk's loads are in
// flight while the current chunk is being added.
constexpr int ACC_CHUNK = 32;
__global__ void __launch_bounds__(64) k_fill_acc(DevMap m, DevCtl* ctl) {
    __shared__ double buf_all[2][2][ACC_CHUNK * 36];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
    const int V = ctl->n_touched;
    long long c_probe = 0, c_pvox = 0;
    for (int vi = wg; vi < V; vi += nW) {
        int j = m.vox_job[vi];
        if (j < 0) continue;
        const int slot = m.touched[vi];
        double* cv = m.cov + (size_t)slot * 36;
        double acc0 = cv[lane];
        double acc1 = lane < 4 ? cv[32 + lane] : 0.0;
        int plane_final = 0, last_plane = -1;
        for (; j >= 0; j = m.job_next[j]) {
            plane_final = m.job_plane[j];
            if (!plane_final) continue;
            last_plane = j;
            const double* src = m.contrib + (size_t)m.job_off[j] * 36;
            const int nt = m.job_nt[j];
            const int nch = (nt + ACC_CHUNK - 1) / ACC_CHUNK;
            double r[36];
            // prologue: chunk 0 -> registers (36 coalesced loads per lane)
            {
                const int tot = (nt < ACC_CHUNK ? nt : ACC_CHUNK) * 36;
#pragma unroll
                for (int k = 0; k < 36; k++) { const int e = k * 32 + lane; r[k] = e < tot ? src[e] : 0.0; }
            }
            for (int ch = 0; ch < nch; ch++) {
                double* buf = buf_all[wib][ch & 1];
#pragma unroll
                for (int k = 0; k < 36; k++) buf[k * 32 + lane] = r[k];
                __syncwarp();
                if (ch + 1 < nch) {                                        // next chunk's loads fly during the adds
                    const int base = (ch + 1) * ACC_CHUNK;
                    const int tot = ((nt - base) < ACC_CHUNK ? (nt - base) : ACC_CHUNK) * 36;
                    const double* sc = src + (size_t)base * 36;
#pragma unroll
                    for (int k = 0; k < 36; k++) { const int e = k * 32 + lane; r[k] = e < tot ? sc[e] : 0.0; }
                }
                const int cb = (nt - ch * ACC_CHUNK) < ACC_CHUNK ? (nt - ch * ACC_CHUNK) : ACC_CHUNK;
                for (int q = 0; q < cb; q++) {                             // strictly in stored-point order (Q7)
                    acc0 += buf[q * 36 + lane];
                    if (lane < 4) acc1 += buf[q * 36 + 32 + lane];
                }
            }
            __syncwarp();
        }
        cv[lane] = acc0;
        if (lane < 4) cv[32 + lane] = acc1;
        if (last_plane >= 0 && lane < 3) {
            m.hot[(size_t)slot * 8 + 3 + lane] = m.job_norm[3 * (size_t)last_plane + lane];
            m.center[(size_t)slot * 3 + lane] = m.job_mean[3 * (size_t)last_plane + lane];
        }
        if (lane == 0) {
            uint32_t flags; int n;
            hot_get_fn(m.hot, slot, flags, n);
            flags = plane_final ? (flags | F_PLANE) : (flags & ~F_PLANE);
            hot_set_fn(m.hot, slot, flags, n);
            const int events = m.evn[slot];
            if (!(flags & F_UE) && (flags & F_PLANE)) { c_probe += events; c_pvox += events > 0 ? 1 : 0; } else m.evn[slot] = 0;
        }
    }
    if (lane == 0 && c_probe) atomicAdd((unsigned long long*)&ctl->st.n_mergeprobe, (unsigned long long)c_probe);
    if (lane == 0 && c_pvox) atomicAdd((unsigned long long*)&ctl->st.n_mergevox, (unsigned long long)c_pvox);
}
