// vmp_map.cu — GPU-resident voxel map update with the reference's strictly sequential
// semantics (VoxelMap::update / build, voxel_map.cpp:200-256; VoxelGrid::pushPoint /
// addToPlane / updatePlane / merge, voxel_map.cpp:29-186), decomposed into parallel phases:
//
//   k_map_insert     key (VoxelMap::index) + find-or-insert in the open-addressing hash
//   k_map_count      per-voxel point count, first/last touching point index, touched list
//   k_seg_scan       exclusive scan of the counts -> per-voxel segments
//   k_seg_fill       point indices grouped by voxel (+ per-block counts for ordered compaction)
//   k_lru_evict      exact LRU victims in creation order (cache.back() semantics, Q17)
//   k_map_fill       one warp per touched voxel: sort its points by index, run the pushPoint
//                    state machine in order, warp-parallel refit (3x3 eigen solve + ordered
//                    accumulation of J Sigma J^T over the stored points)
//   k_merge_prefilter / k_merge_serial
//                    merge(): parallel static candidate filter, then an ordered event
//                    simulation over the (few) voxels whose merge can succeed
//   k_log_append     LRU log append in last-touch order, new stamps
//   k_map_finalize   apply evictions (tombstones, free list), reset per-scan scratch
//
// Why this is exact: a voxel's fill phase depends only on its own points in order; merge()
// only involves voxels that are full (update_enable == false) and never refit again, so an
// event (t, A) is a no-op unless some neighbour pair passes the thresholds on the CURRENT
// planes; the planes change only through successful merges, after which every voxel adjacent
// to a changed one is re-examined.  LRU recency changes only through insertion (Q17), so the
// victim of the j-th over-capacity creation is the oldest log entry whose voxel has not been
// touched earlier in the same scan.
#include "vmp_device.cuh"
#include "vmp_kernels.h"

namespace vmp {

// ------------------------------------------------------------------------- helpers
__device__ __forceinline__ int block_excl_scan(int v, int* total, int* sh /*>=33 ints*/) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) sh[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int w = lane < nw ? sh[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
        sh[lane] = w;      // inclusive over warps
    }
    __syncthreads();
    const int warp_base = wid == 0 ? 0 : sh[wid - 1];
    *total = sh[nw - 1];
    const int r = warp_base + x - v;
    __syncthreads();
    return r;
}

__device__ __forceinline__ void slot_init_fresh(const DevMap& m, DevCtl* ctl, int slot, unsigned long long pk) {
    double* h = m.hot + (size_t)slot * 8;
#pragma unroll
    for (int k = 0; k < 8; k++) h[k] = 0.0;
    hot_set_fn(m.hot, slot, F_UE, 0);
    for (int k = 0; k < 6; k++) m.ppt[(size_t)slot * 6 + k] = 0.0;
    for (int k = 0; k < 36; k++) m.cov[(size_t)slot * 36 + k] = 0.0;      // make_shared<Plane>() value-initialises (Q7)
    for (int k = 0; k < 3; k++) m.center[(size_t)slot * 3 + k] = 0.0;
    m.skey[slot] = pk;
    m.sgroup[slot] = atomicAdd(&ctl->group_counter, 1ull);                 // VoxelGrid::count++ (only equality matters, Q22)
    m.stamp[slot] = 0;
    m.n_temp[slot] = 0;
    m.newly[slot] = 0;
    m.born_scan[slot] = ctl->scan_id;
    m.full_scan[slot] = SCAN_NEVER;
    m.full_idx[slot] = T_INF;
    m.evict_t[slot] = T_INF;
    m.ghost[slot] = -1;
}

__device__ __forceinline__ int pop_free(const DevMap& m, DevCtl* ctl) {
    const int top = atomicSub(&ctl->free_top, 1);
    if (top <= 0) { atomicOr(&ctl->err, E_POOL); atomicAdd(&ctl->free_top, 1); return -1; }
    return m.free_slots[top - 1];
}

// ------------------------------------------------------------------------- init / begin / end
__global__ void k_map_init(DevMap m, DevCtl* ctl) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i <= m.hmask; i += nth) { m.tkey[i] = KEY_EMPTY; m.tval[i] = -1; }
    for (size_t i = tid; i < (size_t)m.pool; i += nth) {
        m.free_slots[i] = m.pool - 1 - (int)i;
        m.stamp[i] = 0; m.cnt[i] = 0; m.cursor[i] = 0; m.ft[i] = T_INF; m.lt[i] = -1; m.seg_off[i] = 0;
        m.evict_t[i] = T_INF; m.ghost[i] = -1; m.evn[i] = 0;
    }
    if (tid == 0) {
        ctl->n = 0; ctl->scan_id = 1; ctl->stamp_base = 1; ctl->n_live = 0;
        ctl->n_touched = ctl->n_new = ctl->n_evict = ctl->n_hot = ctl->n_ghost = 0;
        ctl->free_top = m.pool; ctl->tombstones = 0; ctl->need_rehash = 0; ctl->need_log_compact = 0;
        ctl->log_sel = 0; ctl->log_head = 0; ctl->log_tail = 0; ctl->group_counter = 0; ctl->err = 0;
        ctl->iter = 0; ctl->done = 0; ctl->converged = 0;
    }
}
void launch_map_init(cudaStream_t st, const DevMap& m, DevCtl* ctl) { k_map_init<<<592, 256, 0, st>>>(m, ctl); }

__global__ void k_map_begin(DevCtl* ctl) {
    DevStats z = {};
    ctl->st = z;
    ctl->n_touched = 0; ctl->n_new = 0; ctl->n_evict = 0; ctl->n_hot = 0; ctl->n_ghost = 0;
    for (int q = 0; q < 3; q++) ctl->dbg[q] = 0;          // [3..7] belong to the solve kernel
}

__global__ void k_map_end(DevMap m, DevCtl* ctl) {
    ctl->st.n_points = ctl->n;
    ctl->st.n_touch = ctl->n_touched;
    ctl->st.map_size = ctl->n_live;
    ctl->log_tail += ctl->n_touched;                  // exactly one last-touch entry per touched voxel
    ctl->stamp_base += (unsigned long long)ctl->n;
    ctl->scan_id += 1;
    ctl->need_rehash = (ctl->tombstones > (int)((m.hmask + 1) / 8)) ? 1 : 0;
    ctl->need_log_compact = (ctl->log_tail + 2ll * m.nmax + 2 > m.log_cap) ? 1 : 0;
}

// ------------------------------------------------------------------------- rehash (tombstone purge)
__global__ void k_rehash_clear(DevMap m, const DevCtl* ctl) {
    if (!ctl->need_rehash) return;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i <= m.hmask; i += nth) { m.tkey[i] = KEY_EMPTY; m.tval[i] = -1; }
}
__global__ void k_rehash_insert(DevMap m, DevCtl* ctl) {
    if (!ctl->need_rehash) return;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (size_t s = tid; s < (size_t)m.pool; s += nth) {
        if (m.stamp[s] == 0) continue;
        const unsigned long long pk = m.skey[s];
        unsigned h = hash_key(pk) & m.hmask;
        while (true) {
            const unsigned long long old = atomicCAS(&m.tkey[h], KEY_EMPTY, pk);
            if (old == KEY_EMPTY) { m.tval[h] = (int)s; break; }
            h = (h + 1) & m.hmask;
        }
    }
    if (tid == 0) ctl->tombstones = 0;
}

// ------------------------------------------------------------------------- LRU log compaction
__global__ void __launch_bounds__(1024) k_logc_count(DevMap m, const DevCtl* ctl) {
    if (!ctl->need_log_compact) return;
    const int sel = ctl->log_sel;
    const long long len = ctl->log_tail - ctl->log_head;
    for (long long b = blockIdx.x; b * 1024 < len; b += gridDim.x) {
        const long long p = ctl->log_head + b * 1024 + threadIdx.x;
        int f = 0;
        if (p < ctl->log_tail) { const int s = m.log_slot[sel][p]; f = (m.stamp[s] == m.log_stamp[sel][p]) ? 1 : 0; }
        const int c = __syncthreads_count(f);
        if (threadIdx.x == 0) m.log_blk[b] = c;
    }
}
__global__ void __launch_bounds__(1024) k_logc_scatter(DevMap m, const DevCtl* ctl) {
    if (!ctl->need_log_compact) return;
    __shared__ int sh[34];
    __shared__ long long base_sh;
    const int sel = ctl->log_sel;
    const long long len = ctl->log_tail - ctl->log_head;
    for (long long b = blockIdx.x; b * 1024 < len; b += gridDim.x) {
        if (threadIdx.x == 0) { long long s = 0; for (long long q = 0; q < b; q++) s += m.log_blk[q]; base_sh = s; }
        const long long p = ctl->log_head + b * 1024 + threadIdx.x;
        int f = 0, sl = 0; unsigned long long stp = 0;
        if (p < ctl->log_tail) { sl = m.log_slot[sel][p]; stp = m.log_stamp[sel][p]; f = (m.stamp[sl] == stp) ? 1 : 0; }
        int total;
        const int r = block_excl_scan(f, &total, sh);
        if (f) { m.log_slot[sel ^ 1][base_sh + r] = sl; m.log_stamp[sel ^ 1][base_sh + r] = stp; }
        __syncthreads();
    }
}
__global__ void k_logc_end(DevMap m, DevCtl* ctl) {
    if (!ctl->need_log_compact) return;
    const long long len = ctl->log_tail - ctl->log_head;
    long long s = 0;
    for (long long b = 0; b * 1024 < len; b++) s += m.log_blk[b];
    ctl->log_sel ^= 1; ctl->log_head = 0; ctl->log_tail = s; ctl->need_log_compact = 0;
}

// ------------------------------------------------------------------------- M1a: hash find-or-insert
__global__ void __launch_bounds__(256) k_map_insert(DevMap m, DevScan s, DevCtl* ctl) {
    const int n = ctl->n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned long long pk;
        if (!voxel_index(s.pw[3 * (size_t)i], s.pw[3 * (size_t)i + 1], s.pw[3 * (size_t)i + 2], m.voxel_size, pk)) {
            atomicOr(&ctl->err, E_KEY_RANGE);
            m.tpos[i] = 0xFFFFFFFFu;
            continue;
        }
        unsigned h = hash_key(pk) & m.hmask;
        bool won = false, ok = false;
        for (unsigned probe = 0; probe <= m.hmask; probe++) {
            const unsigned long long cur = __ldcg(&m.tkey[h]);
            if (cur == pk) { ok = true; break; }
            if (cur == KEY_EMPTY) {
                const unsigned long long old = atomicCAS(&m.tkey[h], KEY_EMPTY, pk);
                if (old == KEY_EMPTY) { won = true; ok = true; break; }
                if (old == pk) { ok = true; break; }
            }
            h = (h + 1) & m.hmask;
        }
        if (!ok) { atomicOr(&ctl->err, E_HASH_FULL); m.tpos[i] = 0xFFFFFFFFu; continue; }
        m.tpos[i] = h;
        if (won) {
            const int slot = pop_free(m, ctl);
            if (slot >= 0) slot_init_fresh(m, ctl, slot, pk);
            m.tval[h] = slot;
            atomicAdd(&ctl->n_new, 1);
        }
    }
}

// ------------------------------------------------------------------------- M1b: counts / first / last touch
__global__ void __launch_bounds__(256) k_map_count(DevMap m, DevCtl* ctl) {
    const int n = ctl->n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned h = m.tpos[i];
        int slot = -1;
        if (h != 0xFFFFFFFFu) slot = m.tval[h];
        m.pslot[i] = slot;
        if (slot < 0) continue;
        const int c = atomicAdd(&m.cnt[slot], 1);
        if (c == 0) m.touched[atomicAdd(&ctl->n_touched, 1)] = slot;
        atomicMin(&m.ft[slot], i);
        atomicMax(&m.lt[slot], i);
    }
}

// ------------------------------------------------------------------------- M3a: segment offsets
__global__ void __launch_bounds__(1024) k_seg_scan(DevMap m, const DevCtl* ctl) {
    __shared__ int sh[34];
    const int V = ctl->n_touched;
    int carry = 0;
    for (int base = 0; base < V; base += 1024) {
        const int idx = base + threadIdx.x;
        const int slot = idx < V ? m.touched[idx] : -1;
        const int v = slot >= 0 ? m.cnt[slot] : 0;
        int total;
        const int r = block_excl_scan(v, &total, sh);
        if (slot >= 0) m.seg_off[slot] = carry + r;
        carry += total;
    }
}

// ------------------------------------------------------------------------- M3b: fill segments
__global__ void __launch_bounds__(1024) k_seg_fill(DevMap m, DevCtl* ctl) {
    const int n = ctl->n;
    const unsigned scan_id = ctl->scan_id;
    for (int b = blockIdx.x; b * PT_BLOCK < n; b += gridDim.x) {
        const int i = b * PT_BLOCK + threadIdx.x;
        int is_last = 0, is_new = 0;
        if (i < n) {
            const int slot = m.pslot[i];
            if (slot >= 0) {
                const int pos = m.seg_off[slot] + atomicAdd(&m.cursor[slot], 1);
                m.seg[pos] = i;
                is_last = (m.lt[slot] == i);
                is_new = (m.ft[slot] == i && m.born_scan[slot] == scan_id);
            }
        }
        const int cl = __syncthreads_count(is_last);
        const int cn = __syncthreads_count(is_new);
        if (threadIdx.x == 0) { m.blk_last[b] = cl; m.blk_new[b] = cn; }
    }
}

// ------------------------------------------------------------------------- M2: exact LRU eviction
// cache.push_front / splice / "if (cache.size() > capacity) erase(cache.back())"
// (voxel_map.cpp:242-253).  One CTA: ordered compaction of the creation times, then a
// short serial walk over the head of the LRU log.
__global__ void __launch_bounds__(1024) k_lru_evict(DevMap m, DevCtl* ctl) {
    __shared__ int sh[34];
    __shared__ int rq[64];
    const int n = ctl->n;
    const int n_live0 = ctl->n_live, n_new = ctl->n_new;
    if (n_live0 + n_new <= m.capacity) {
        if (threadIdx.x == 0) { ctl->n_live = n_live0 + n_new; ctl->st.n_created = n_new; }
        return;
    }
    const unsigned scan_id = ctl->scan_id;
    int base = 0;
    for (int b = 0; b * PT_BLOCK < n; b++) {
        const int nb = m.blk_new[b];
        if (nb == 0) continue;
        const int i = b * PT_BLOCK + threadIdx.x;
        int f = 0;
        if (i < n) { const int slot = m.pslot[i]; if (slot >= 0) f = (m.ft[slot] == i && m.born_scan[slot] == scan_id); }
        int total;
        const int r = block_excl_scan(f, &total, sh);
        if (f) m.ct[base + r] = i;
        base += nb;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    int size = n_live0, ne = 0, rqn = 0, ci = 0, created = n_new;
    long long h = ctl->log_head;
    const long long tail = ctl->log_tail;
    const int sel = ctl->log_sel;
    while (ci < n_new || rqn > 0) {
        int t;
        if (rqn > 0 && (ci >= n_new || rq[0] < m.ct[ci])) { t = rq[0]; for (int q = 1; q < rqn; q++) rq[q - 1] = rq[q]; rqn--; }
        else t = m.ct[ci++];
        size += 1;
        if (size <= m.capacity) continue;
        int victim = -1;
        while (h < tail) {
            const int sl = m.log_slot[sel][h];
            const unsigned long long stp = m.log_stamp[sel][h];
            h++;
            if (m.stamp[sl] != stp) continue;                       // stale entry (lazy deletion)
            if (m.cnt[sl] > 0 && m.ft[sl] < t) continue;            // spliced to the front earlier in this scan
            victim = sl;
            break;
        }
        if (victim < 0) { atomicOr(&ctl->err, E_LRU_EXHAUSTED); break; }
        size -= 1;
        m.ev_time[ne] = t;
        m.ev_key[ne] = m.skey[victim];
        if (m.cnt[victim] > 0) {
            // evicted now, touched again later in this scan: the key is re-created fresh at
            // ft[victim]; the old incarnation stays visible to merge() until t as a ghost slot
            const int g = pop_free(m, ctl);
            if (g < 0) break;
            for (int k = 0; k < 8; k++) m.hot[(size_t)g * 8 + k] = m.hot[(size_t)victim * 8 + k];
            for (int k = 0; k < 36; k++) m.cov[(size_t)g * 36 + k] = m.cov[(size_t)victim * 36 + k];
            m.skey[g] = m.skey[victim]; m.sgroup[g] = m.sgroup[victim];
            m.full_scan[g] = m.full_scan[victim]; m.full_idx[g] = m.full_idx[victim];
            m.born_scan[g] = m.born_scan[victim]; m.stamp[g] = 0; m.evict_t[g] = t; m.ghost[g] = m.ghost[victim];
            slot_init_fresh(m, ctl, victim, m.skey[g]);
            m.ghost[victim] = g;
            m.ev_slot[ne] = g | 0x40000000;
            if (rqn >= 64) { atomicOr(&ctl->err, E_QUEUE); break; }
            int q = rqn++;
            const int tt = m.ft[victim];
            while (q > 0 && rq[q - 1] > tt) { rq[q] = rq[q - 1]; q--; }
            rq[q] = tt;
            created++;
        } else {
            m.evict_t[victim] = t;
            m.ev_slot[ne] = victim;
        }
        ne++;
    }
    ctl->n_evict = ne; ctl->n_live = size; ctl->log_head = h;
    ctl->st.n_evicted = ne; ctl->st.n_created = created;
}

// ------------------------------------------------------------------------- M4: per-voxel fill + refit
// ascending in-place sort of a[0..c) by one warp (point indices are distinct)
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

// J Sigma J^T of one stored point (voxel_map.cpp:115-129)
__device__ __forceinline__ void plane_contrib(const V3& p, const M3& S, const V3& mean, int n, const double* evals,
                                              const M3& evecs, const V3& nrm, double* out /*36, stride 1*/) {
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

constexpr int CONTRIB_STRIDE = 37;

struct VoxelRun {            // warp-uniform running state of the voxel being filled
    V3 mean; double ppt[6]; int n; uint32_t flags; int nt, nw;
    unsigned full_scan; int full_idx;
};

// updatePlane() body after the "n >= update_point_thresh" test (voxel_map.cpp:100-135).
// idx != nullptr: read the points through seg indices from the scan arrays (build()),
// else from the voxel's stored points.
__device__ void warp_refit(const DevMap& m, const DevScan& s, DevCtl* ctl, int slot, VoxelRun& v, const int* idx,
                           double* shc /*32*CONTRIB_STRIDE*/, long long& c_refit, long long& c_refit_pts) {
    const int lane = threadIdx.x & 31;
    v.flags |= F_INIT;
    const double nd = (double)v.n;
    const double c00 = v.ppt[0] / nd - v.mean[0] * v.mean[0];
    const double c10 = v.ppt[1] / nd - v.mean[1] * v.mean[0];
    const double c11 = v.ppt[2] / nd - v.mean[1] * v.mean[1];
    const double c20 = v.ppt[3] / nd - v.mean[2] * v.mean[0];
    const double c21 = v.ppt[4] / nd - v.mean[2] * v.mean[1];
    const double c22 = v.ppt[5] / nd - v.mean[2] * v.mean[2];
    double evals[3];
    M3 evecs;
    eig3_sym(c00, c10, c11, c20, c21, c22, evals, evecs);
    c_refit++;
    if (evals[0] > m.plane_thresh) { v.flags &= ~F_PLANE; return; }     // Q13: norm / cov stay
    v.flags |= F_PLANE;
    V3 nrm = v3(evecs(0, 0), evecs(1, 0), evecs(2, 0));
    const int np = v.nt;
    c_refit_pts += np;
    if (!idx && np > m.maxpt) { if (lane == 0) atomicOr(&ctl->err, E_REFIT_OVERFLOW); }
    double* cv = m.cov + (size_t)slot * 36;
    double acc0 = cv[lane];
    double acc1 = lane < 4 ? cv[32 + lane] : 0.0;
    const double* tp = m.tp + (size_t)slot * 12 * m.maxpt;
    for (int base = 0; base < np; base += 32) {
        const int j = base + lane;
        if (j < np) {
            V3 p; M3 S;
            if (idx) {
                const int i = idx[j];
                p = v3(s.pw[3 * (size_t)i], s.pw[3 * (size_t)i + 1], s.pw[3 * (size_t)i + 2]);
#pragma unroll
                for (int k = 0; k < 9; k++) S.a[k] = s.pcov[9 * (size_t)i + k];
            } else {
                const int jj = j < m.maxpt ? j : m.maxpt - 1;
                p = v3(tp[jj], tp[m.maxpt + jj], tp[2 * m.maxpt + jj]);
#pragma unroll
                for (int k = 0; k < 9; k++) S.a[k] = tp[(size_t)(3 + k) * m.maxpt + jj];
            }
            plane_contrib(p, S, v.mean, v.n, evals, evecs, nrm, shc + lane * CONTRIB_STRIDE);
        }
        __syncwarp();
        const int cb = np - base < 32 ? np - base : 32;
        for (int q = 0; q < cb; q++) {                      // ordered accumulation (Q7: never reset)
            acc0 += shc[q * CONTRIB_STRIDE + lane];
            if (lane < 4) acc1 += shc[q * CONTRIB_STRIDE + 32 + lane];
        }
        __syncwarp();
    }
    cv[lane] = acc0;
    if (lane < 4) cv[32 + lane] = acc1;
    const double axis_distance = -dot(v.mean, nrm);
    if (axis_distance < 0.0) nrm = neg(nrm);
    if (lane < 3) {
        m.hot[(size_t)slot * 8 + 3 + lane] = nrm[lane];
        m.center[(size_t)slot * 3 + lane] = v.mean[lane];
    }
}

__global__ void __launch_bounds__(128) k_map_fill(DevMap m, DevScan s, DevCtl* ctl, int build) {
    __shared__ double shc_all[4][32 * CONTRIB_STRIDE];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* shc = shc_all[wib];
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
    const int V = ctl->n_touched;
    const unsigned scan_id = ctl->scan_id;
    long long c_ins = 0, c_refit = 0, c_refit_pts = 0, c_full = 0, c_probe = 0;
    for (int vi = wg; vi < V; vi += nW) {
        const int slot = m.touched[vi];
        const int c = m.cnt[slot], off = m.seg_off[slot];
        VoxelRun v;
        hot_get_fn(m.hot, slot, v.flags, v.n);
        int events = 0;
        if (!(v.flags & F_UE)) {
            events = c;                                     // full before this scan: merge() or nothing per point
        } else {
            int* sg = m.seg + off;
            warp_sort(sg, c);
            const double* h = m.hot + (size_t)slot * 8;
            v.mean = v3(h[0], h[1], h[2]);
#pragma unroll
            for (int k = 0; k < 6; k++) v.ppt[k] = m.ppt[(size_t)slot * 6 + k];
            v.nt = m.n_temp[slot]; v.nw = m.newly[slot];
            v.full_scan = SCAN_NEVER; v.full_idx = T_INF;
            double* tp = m.tp + (size_t)slot * 12 * m.maxpt;
            int j = 0;
            for (; j < c; j++) {
                if (!(v.flags & F_UE)) break;
                const int i = sg[j];
                const V3 p = v3(s.pw[3 * (size_t)i], s.pw[3 * (size_t)i + 1], s.pw[3 * (size_t)i + 2]);
                // addToPlane (voxel_map.cpp:29-34)
                v.mean = add(v.mean, divs(sub(p, v.mean), v.n + 1.0));
                v.ppt[0] += p[0] * p[0]; v.ppt[1] += p[1] * p[0]; v.ppt[2] += p[1] * p[1];
                v.ppt[3] += p[2] * p[0]; v.ppt[4] += p[2] * p[1]; v.ppt[5] += p[2] * p[2];
                v.n += 1;
                // temp_points.push_back
                if (v.nt < m.maxpt) {
                    if (lane < 3) tp[(size_t)lane * m.maxpt + v.nt] = p[lane];
                    else if (lane < 12) tp[(size_t)lane * m.maxpt + v.nt] = s.pcov[9 * (size_t)i + (lane - 3)];
                }
                v.nt += 1;
                c_ins++;
                if (build) continue;                        // addPoint (voxel_map.cpp:36-40)
                __syncwarp();
                if (!(v.flags & F_INIT)) {
                    if (v.n >= m.upt) warp_refit(m, s, ctl, slot, v, nullptr, shc, c_refit, c_refit_pts);
                } else {
                    v.nw += 1;
                    if (v.nw >= m.upt) { warp_refit(m, s, ctl, slot, v, nullptr, shc, c_refit, c_refit_pts); v.nw = 0; }
                    if (v.nt >= m.maxpt) {                   // update_enable = false; temp_points freed
                        v.flags &= ~F_UE; v.full_scan = scan_id; v.full_idx = i; v.nt = 0;
                    }
                }
            }
            events = c - j;
            if (build && v.n >= m.upt) { __syncwarp(); warp_refit(m, s, ctl, slot, v, sg, shc, c_refit, c_refit_pts); }
            if (lane == 0) {
                double* hw = m.hot + (size_t)slot * 8;
                hw[0] = v.mean[0]; hw[1] = v.mean[1]; hw[2] = v.mean[2];
                hot_set_fn(m.hot, slot, v.flags, v.n);
                for (int k = 0; k < 6; k++) m.ppt[(size_t)slot * 6 + k] = v.ppt[k];
                m.n_temp[slot] = v.nt; m.newly[slot] = v.nw;
                if (v.full_scan != SCAN_NEVER) { m.full_scan[slot] = v.full_scan; m.full_idx[slot] = v.full_idx; }
            }
        }
        c_full += events;
        if (!(v.flags & F_UE) && (v.flags & F_PLANE)) c_probe += events; else events = 0;
        if (lane == 0) m.evn[slot] = events;                // merge() invocations of this voxel in this scan
    }
    if (lane == 0) {
        if (c_ins) atomicAdd((unsigned long long*)&ctl->st.n_ins, (unsigned long long)c_ins);
        if (c_refit) atomicAdd((unsigned long long*)&ctl->st.n_refit, (unsigned long long)c_refit);
        if (c_refit_pts) atomicAdd((unsigned long long*)&ctl->st.refit_points, (unsigned long long)c_refit_pts);
        if (c_full) atomicAdd((unsigned long long*)&ctl->st.n_full, (unsigned long long)c_full);
        if (c_probe) atomicAdd((unsigned long long*)&ctl->st.n_mergeprobe, (unsigned long long)c_probe);
    }
}

#include "vmp_merge.cuh"

// ------------------------------------------------------------------------- M7a: LRU log append
__global__ void __launch_bounds__(1024) k_log_append(DevMap m, DevCtl* ctl) {
    __shared__ int sh[34];
    __shared__ int base_sh;
    const int n = ctl->n;
    const int sel = ctl->log_sel;
    const long long tail = ctl->log_tail;
    const unsigned long long sb = ctl->stamp_base;
    for (int b = blockIdx.x; b * PT_BLOCK < n; b += gridDim.x) {
        if (threadIdx.x == 0) { int sacc = 0; for (int q = 0; q < b; q++) sacc += m.blk_last[q]; base_sh = sacc; }
        const int i = b * PT_BLOCK + threadIdx.x;
        int f = 0, slot = -1;
        if (i < n) { slot = m.pslot[i]; if (slot >= 0) f = (m.lt[slot] == i); }
        int total;
        const int r = block_excl_scan(f, &total, sh);
        if (f) {
            const long long pos = tail + base_sh + r;
            if (pos < m.log_cap) { m.log_slot[sel][pos] = slot; m.log_stamp[sel][pos] = sb + (unsigned long long)i; }
            else atomicOr(&ctl->err, E_QUEUE);
            m.stamp[slot] = sb + (unsigned long long)i;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------- M7b: finalize
__global__ void __launch_bounds__(256) k_map_finalize(DevMap m, DevCtl* ctl) {
    const int V = ctl->n_touched, E = ctl->n_evict;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int vi = tid; vi < V; vi += nth) {
        const int s = m.touched[vi];
        m.cnt[s] = 0; m.cursor[s] = 0; m.ft[s] = T_INF; m.lt[s] = -1; m.evn[s] = 0; m.seg_off[s] = 0;
    }
    for (int e = tid; e < E; e += nth) {
        const int es = m.ev_slot[e];
        const bool is_ghost = (es & 0x40000000) != 0;
        const int slot = es & 0x3FFFFFFF;
        const unsigned long long pk = m.ev_key[e];
        unsigned h = hash_key(pk) & m.hmask;
        for (unsigned probe = 0; probe <= m.hmask; probe++) {
            const unsigned long long cur = m.tkey[h];
            if (cur == pk) break;
            if (cur == KEY_EMPTY) { h = 0xFFFFFFFFu; break; }
            h = (h + 1) & m.hmask;
        }
        if (is_ghost) {
            if (h != 0xFFFFFFFFu) { const int owner = m.tval[h]; if (owner >= 0 && m.ghost[owner] == slot) m.ghost[owner] = -1; }
        } else {
            if (h != 0xFFFFFFFFu) { m.tkey[h] = KEY_TOMB; m.tval[h] = -1; atomicAdd(&ctl->tombstones, 1); }
        }
        m.stamp[slot] = 0; m.evict_t[slot] = T_INF; m.ghost[slot] = -1;
        m.free_slots[atomicAdd(&ctl->free_top, 1)] = slot;
    }
}

// ------------------------------------------------------------------------- launcher
int launch_map_update(cudaStream_t st, const DevMap& m, const DevScan& s, DevCtl* ctl, int sm_count, bool build, const Marker* mk) {
    const int gpt = (m.nmax + PT_BLOCK - 1) / PT_BLOCK;               // order-preserving passes: 1024 points / block
    const int gstride = sm_count * 2;
    int launches = 0;
    k_map_begin<<<1, 1, 0, st>>>(ctl); launches++; mark(mk, VMP_K_MAP_BEGIN);
    k_map_insert<<<gstride, 256, 0, st>>>(m, s, ctl); launches++; mark(mk, VMP_K_MAP_INSERT);
    k_map_count<<<gstride, 256, 0, st>>>(m, ctl); launches++; mark(mk, VMP_K_MAP_COUNT);
    k_seg_scan<<<1, 1024, 0, st>>>(m, ctl); launches++; mark(mk, VMP_K_SEG_SCAN);
    k_seg_fill<<<gpt, 1024, 0, st>>>(m, ctl); launches++; mark(mk, VMP_K_SEG_FILL);
    k_lru_evict<<<1, 1024, 0, st>>>(m, ctl); launches++; mark(mk, VMP_K_LRU_EVICT);
    k_map_fill<<<sm_count * 4, 128, 0, st>>>(m, s, ctl, build ? 1 : 0); launches++; mark(mk, VMP_K_MAP_FILL);
    if (!build) {
        k_merge_prefilter<<<sm_count, 128, 0, st>>>(m, ctl); launches++; mark(mk, VMP_K_MERGE_PREFILTER);
        k_merge_rounds<<<1, 512, 0, st>>>(m, ctl); launches++; mark(mk, VMP_K_MERGE_SERIAL);
    }
    k_log_append<<<gpt, 1024, 0, st>>>(m, ctl); launches++; mark(mk, VMP_K_LOG_APPEND);
    k_map_finalize<<<sm_count, 256, 0, st>>>(m, ctl); launches++; mark(mk, VMP_K_MAP_FINALIZE);
    k_map_end<<<1, 1, 0, st>>>(m, ctl); launches++; mark(mk, VMP_K_MAP_END);
    return launches;
}

// Rare maintenance, launched by the host between scans only when the previous scan asked for it
// (tombstone purge of the hash, compaction of the LRU log); every kernel re-checks its device flag.
int launch_map_maintenance(cudaStream_t st, const DevMap& m, DevCtl* ctl, int sm_count, int what) {
    int launches = 0;
    if (what & 1) {
        k_rehash_clear<<<sm_count * 2, 256, 0, st>>>(m, ctl); launches++;
        k_rehash_insert<<<sm_count * 2, 256, 0, st>>>(m, ctl); launches++;
    }
    if (what & 2) {
        k_logc_count<<<sm_count, 1024, 0, st>>>(m, ctl); launches++;
        k_logc_scatter<<<sm_count, 1024, 0, st>>>(m, ctl); launches++;
        k_logc_end<<<1, 1, 0, st>>>(m, ctl); launches++;
    }
    return launches;
}

}  // namespace vmp
