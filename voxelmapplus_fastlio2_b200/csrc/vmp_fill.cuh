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

// updatePlane() reached its body: record the job (whole warp calls; returns the job id in every lane)
__device__ int emit_job(const DevMap& m, DevCtl* ctl, int slot, int n, int nt, const V3& mean, const double* ppt,
                        int src_off, int prev_job) {
    const int lane = threadIdx.x & 31;
    int j = 0, b0 = 0;
    const int nb = (nt + 31) >> 5;
    if (lane == 0) {
        j = atomicAdd(&ctl->n_jobs, 1);
        const unsigned long long off = atomicAdd(&ctl->contrib_top, (unsigned long long)nt);
        b0 = atomicAdd(&ctl->n_batches, nb);
        if (j >= m.job_cap || off + (unsigned long long)nt > (unsigned long long)m.contrib_cap || b0 + nb > m.bat_cap) {
            atomicOr(&ctl->err, E_QUEUE);
            j = -1;
        } else {
            m.job_slot[j] = slot; m.job_n[j] = n; m.job_nt[j] = nt; m.job_off[j] = (long long)off; m.job_src[j] = src_off;
            m.job_next[j] = -1; m.job_plane[j] = 0;
            for (int k = 0; k < 3; k++) m.job_mean[3 * (size_t)j + k] = mean[k];
            for (int k = 0; k < 6; k++) m.job_ppt[6 * (size_t)j + k] = ppt[k];
            if (prev_job >= 0) m.job_next[prev_job] = j;
        }
    }
    j = __shfl_sync(0xffffffffu, j, 0);
    b0 = __shfl_sync(0xffffffffu, b0, 0);
    if (j >= 0) for (int b = lane; b < nb; b += 32) { m.bat_job[b0 + b] = j; m.bat_idx[b0 + b] = b; }
    return j;
}

__global__ void __launch_bounds__(128) k_fill_state(DevMap m, DevScan s, DevCtl* ctl, int build) {
    __shared__ int sel_all[4][SEL_MAX];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int* sel = sel_all[wib];
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
    const int V = ctl->n_touched, npts = ctl->n;
    const unsigned scan_id = ctl->scan_id;
    long long c_ins = 0, c_full = 0, c_probe = 0;
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
            // the voxel consumes at most K points before it closes (never more than its free room, at least one)
            int K = c;
            const int* order;
            if (build) {
                warp_sort(m.seg + off, c);                  // build() has no cap (Q18): every point, in order
                order = m.seg + off;
            } else {
                const int room = m.maxpt - nt;
                K = room < 1 ? 1 : room;
                if (K > c) K = c;
                if (K <= SEL_MAX) { warp_select_sorted(m.seg + off, c, K, npts, sel); order = sel; }
                else { warp_sort(m.seg + off, c); order = m.seg + off; }
            }
            const double* h = m.hot + (size_t)slot * 8;
            V3 mean = v3(h[0], h[1], h[2]);
            double ppt[6];
#pragma unroll
            for (int k = 0; k < 6; k++) ppt[k] = m.ppt[(size_t)slot * 6 + k];
            unsigned full_scan = SCAN_NEVER; int full_idx = T_INF;
            double* tp = m.tp + (size_t)slot * 12 * m.maxpt;
            int prev_job = -1, j = 0;
            for (; j < K; j++) {
                if (!(flags & F_UE) && !build) break;
                const int i = order[j];
                const V3 p = v3(s.pw[3 * (size_t)i], s.pw[3 * (size_t)i + 1], s.pw[3 * (size_t)i + 2]);
                // addToPlane (voxel_map.cpp:29-34)
                mean = add(mean, divs(sub(p, mean), n + 1.0));
                ppt[0] += p[0] * p[0]; ppt[1] += p[1] * p[0]; ppt[2] += p[1] * p[1];
                ppt[3] += p[2] * p[0]; ppt[4] += p[2] * p[1]; ppt[5] += p[2] * p[2];
                n += 1;
                // temp_points.push_back
                if (nt < m.maxpt) {
                    if (lane < 3) tp[(size_t)lane * m.maxpt + nt] = p[lane];
                    else if (lane < 12) tp[(size_t)lane * m.maxpt + nt] = s.pcov[9 * (size_t)i + (lane - 3)];
                }
                nt += 1;
                c_ins++;
                if (build) continue;                        // addPoint (voxel_map.cpp:36-40)
                bool refit = false;
                if (!(flags & F_INIT)) {
                    refit = n >= m.upt;                     // updatePlane() every point, early return while n < thresh
                } else {
                    nw += 1;
                    if (nw >= m.upt) { refit = true; nw = 0; }
                }
                const bool was_init = (flags & F_INIT) != 0;
                if (refit) {
                    flags |= F_INIT;
                    const int jb = emit_job(m, ctl, slot, n, nt, mean, ppt, -1, prev_job);
                    if (first_job < 0) first_job = jb;
                    prev_job = jb;
                }
                if (was_init && nt >= m.maxpt) {            // update_enable = false; temp_points freed
                    flags &= ~F_UE; full_scan = scan_id; full_idx = i; nt = 0;
                }
            }
            events = c - j;
            if (!build && j == K && K < c && (flags & F_UE) && lane == 0) atomicOr(&ctl->err, E_QUEUE);   // cannot happen: K points always close the voxel
            if (build && n >= m.upt) {                       // one updatePlane() per voxel at the end of build()
                flags |= F_INIT;
                first_job = emit_job(m, ctl, slot, n, nt, mean, ppt, off, -1);
            }
            __syncwarp();
            if (lane == 0) {
                double* hw = m.hot + (size_t)slot * 8;
                hw[0] = mean[0]; hw[1] = mean[1]; hw[2] = mean[2];
                hot_set_fn(m.hot, slot, flags, n);
                for (int k = 0; k < 6; k++) m.ppt[(size_t)slot * 6 + k] = ppt[k];
                m.n_temp[slot] = nt; m.newly[slot] = nw;
                if (full_scan != SCAN_NEVER) { m.full_scan[slot] = full_scan; m.full_idx[slot] = full_idx; }
            }
        }
        c_full += events;
        if (first_job < 0) {
            // no refit in this scan: is_plane is final, settle the merge() bookkeeping here
            if (!(flags & F_UE) && (flags & F_PLANE)) c_probe += events; else events = 0;
        }
        if (lane == 0) { m.evn[slot] = events; m.vox_job[vi] = first_job; }
    }
    if (lane == 0) {
        if (c_ins) atomicAdd((unsigned long long*)&ctl->st.n_ins, (unsigned long long)c_ins);
        if (c_full) atomicAdd((unsigned long long*)&ctl->st.n_full, (unsigned long long)c_full);
        if (c_probe) atomicAdd((unsigned long long*)&ctl->st.n_mergeprobe, (unsigned long long)c_probe);
    }
}

// updatePlane() body (voxel_map.cpp:100-135) for one batch of 32 stored points of one job
__global__ void __launch_bounds__(128) k_fill_refit(DevMap m, DevScan s, DevCtl* ctl) {
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
            double* dst = m.contrib + ((size_t)m.job_off[j] + q) * 36;
#pragma unroll
            for (int k = 0; k < 36; k++) dst[k] = out[k];
        }
    }
}

// plane->cov += J Sigma J^T, jobs in refit order, points in stored order; final normal / centre / is_plane
__global__ void __launch_bounds__(128) k_fill_acc(DevMap m, DevCtl* ctl) {
    const int lane = threadIdx.x & 31;
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
    const int V = ctl->n_touched;
    long long c_probe = 0;
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
            int q = 0;
            for (; q + 4 <= nt; q += 4) {                                  // loads batched, adds strictly in order
                const double a0 = src[(size_t)q * 36 + lane], a1 = src[(size_t)(q + 1) * 36 + lane];
                const double a2 = src[(size_t)(q + 2) * 36 + lane], a3 = src[(size_t)(q + 3) * 36 + lane];
                double b0 = 0, b1 = 0, b2 = 0, b3 = 0;
                if (lane < 4) { b0 = src[(size_t)q * 36 + 32 + lane]; b1 = src[(size_t)(q + 1) * 36 + 32 + lane];
                                b2 = src[(size_t)(q + 2) * 36 + 32 + lane]; b3 = src[(size_t)(q + 3) * 36 + 32 + lane]; }
                acc0 += a0; acc0 += a1; acc0 += a2; acc0 += a3;
                if (lane < 4) { acc1 += b0; acc1 += b1; acc1 += b2; acc1 += b3; }
            }
            for (; q < nt; q++) {
                acc0 += src[(size_t)q * 36 + lane];
                if (lane < 4) acc1 += src[(size_t)q * 36 + 32 + lane];
            }
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
            if (!(flags & F_UE) && (flags & F_PLANE)) c_probe += events; else m.evn[slot] = 0;
        }
    }
    if (lane == 0 && c_probe) atomicAdd((unsigned long long*)&ctl->st.n_mergeprobe, (unsigned long long)c_probe);
}
