// vmp_map.cu — GPU-resident voxel map update with the reference's strictly sequential
// semantics (VoxelMap::update / build, voxel_map.cpp:200-256; VoxelGrid::pushPoint /
// addToPlane / updatePlane / merge, voxel_map.cpp:29-186), decomposed into parallel phases:
//
//   k_world_insert_count   world point + covariance of every scan point (lio_builder.cpp:233-245), key (VoxelMap::index) +
//                    find-or-insert in the open-addressing hash, per-voxel point count, first / last touching point index,
//                    touched list; its last CTA lays out the per-voxel segments (disjoint, unordered: no scan kernel)
//   k_seg_fill       which touched voxels take the CTA path of k_fill (+ L2 prefetch of what k_fill will read), then the point
//                    indices grouped by voxel (+ per-block counts for ordered compaction)
//   k_lru_evict      exact LRU victims in creation order (cache.back() semantics, Q17); side branch of the graph
//   k_fill<false> / k_fill<true>   (vmp_fill.cuh) pushPoint state machine per touched voxel and the refits it triggers (3x3
//                    eigen-solves one per lane, J Sigma J^T per stored point, ordered accumulation): a warp per voxel, and a
//                    CTA per voxel for the voxels whose refits loop over many points; two launches side by side
//   k_merge_prefilter / k_merge_rounds          (vmp_merge.cuh) merge(): parallel static candidate filter, then an event
//                    simulation in rounds of spatially independent events on a 4-CTA cluster, exact serial redo behind it
//   k_log_append     LRU log append in last-touch order, new stamps; side branch
//   k_map_finalize   apply evictions (tombstones, free list), reset per-scan scratch; its last CTA closes the update
//                    (counters, maintenance requests, host mailbox)
//
// Why this is exact: a voxel's fill phase depends only on its own points in order; merge()
// only involves voxels that are full (update_enable == false) and never refit again, so an
// event (t, A) is a no-op unless some neighbour pair passes the thresholds on the CURRENT
// planes; the planes change only through successful merges, after which every voxel adjacent
// to a changed one is re-examined.  LRU recency changes only through insertion (Q17), so the
// victim of the j-th over-capacity creation is the oldest log entry whose voxel has not been
// touched earlier in the same scan.
#include <cooperative_groups.h>

#include "vmp_device.cuh"
#include "vmp_kernels.h"
#include "vmp_state.cuh"

namespace cg = cooperative_groups;

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
cudaError_t map_configure_kernels(const DevMap& m);


// end of VoxelMap::update (single thread, the last CTA of k_map_finalize): counters, maintenance requests, host mailbox
__device__ void map_end(const DevMap& m, DevCtl* ctl, MapOut* out) {
    ctl->st.n_points = ctl->n;
    ctl->st.n_touch = ctl->n_touched;
    ctl->st.map_size = ctl->n_live;
    ctl->log_tail += ctl->n_touched;                  // exactly one last-touch entry per touched voxel
    ctl->stamp_base += (unsigned long long)ctl->n;
    ctl->scan_id += 1;
    // requested early enough that they may be served one update late (pipelined mode): an update adds at most nmax
    // tombstones and nmax log entries
    ctl->need_rehash = (atomicAdd(&ctl->tombstones, 0) > (int)((m.hmask + 1) / 8)) ? 1 : 0;     // (other CTAs' atomics)
    ctl->need_log_compact = (ctl->log_tail + 4ll * m.nmax + 4 > m.log_cap) ? 1 : 0;
    if (out) {
        out += ctl->seq & 1ull;                       // double-buffered by scan parity: a pipelined host reads scan k-1's while scan k runs
        out->st = ctl->st;
        out->err = ctl->err;
        out->need_maint = (ctl->need_rehash ? 1 : 0) | (ctl->need_log_compact ? 2 : 0);
        for (int q = 0; q < 8; q++) out->dbg[q] = ctl->dbg[q];
        __threadfence_system();
        *(volatile unsigned long long*)&out->seq = ctl->seq;
    }
    ctl->err = 0;                                     // error bits are per update: reported once, then cleared
    map_counters_reset(ctl);                          // the next update starts from clean counters (there is no "begin" kernel)
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
        if (threadIdx.x < 32) {
            long long s = 0;
            for (long long q = threadIdx.x; q < b; q += 32) s += m.log_blk[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (threadIdx.x == 0) base_sh = s;
        }
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

// ------------------------------------------------------------------------- M1: world points, hash find-or-insert, counts, segment offsets
// ONE pass over the points of the scan (round 1: k_world_points, k_map_insert, k_map_count):
//   * lidarToWorld + pv_list loop (lio_builder.cpp:155-163, 231-245): float32 world transform with the association of PCL's SSE
//     Transformer::se3 (x' = m00 x + (m01 y + (m02 z + tx)), separate mul / add, Q15), widened to fp64, and pv.cov with the posterior
//     R, P (world_mode 0: a scan; 1: the first scan, calcBodyCov on a local copy, lio_builder.cpp:196-198; 2: pv_list given by the caller)
//   * VoxelMap::index + featmap find-or-create (voxel_map.cpp:234-247): the lanes of a warp that hit the same voxel are grouped
//     (match.any on the packed key; consecutive points of a scan share voxels) and their leader does ONE hash operation and ONE
//     count / first-touch / last-touch update for the group.  A leader that finds a key whose creator (another warp) has not
//     published the slot yet waits for it; the creator never waits for anybody.
//   * its last CTA hands every touched voxel its segment of the per-voxel point lists (segments only have to be disjoint - they are
//     selected / sorted by point index later - so there is no ordered scan: warp prefix + one shared atomic per warp).
struct WorldState { M3 r_wl, Prr, Ppp; V3 p_wl; float mf[12]; };

__device__ __forceinline__ int find_or_insert(const DevMap& m, DevCtl* ctl, unsigned long long pk) {
    unsigned h = hash_key(pk) & m.hmask;
    for (unsigned probe = 0; probe <= m.hmask; probe++) {
        unsigned long long cur = __ldcg(&m.tkey[h]);
        if (cur == KEY_EMPTY) {
            cur = atomicCAS(&m.tkey[h], KEY_EMPTY, pk);
            if (cur == KEY_EMPTY) {                                 // created (voxel_map.cpp:236-241)
                const int slot = pop_free(m, ctl);
                if (slot >= 0) {
                    slot_init_fresh(m, ctl, slot, pk);
                    m.newlist[atomicAdd(&ctl->n_new, 1)] = slot;
                    __threadfence();                                // the slot is initialised before it becomes visible
                }
                *(volatile int*)&m.tval[h] = slot >= 0 ? slot : -2;
                return slot;
            }
        }
        if (cur == pk) {
            int v;
            while ((v = *(volatile int*)&m.tval[h]) == -1) __nanosleep(20);      // being created by another warp right now
            return v >= 0 ? v : -1;
        }
        h = (h + 1) & m.hmask;
    }
    atomicOr(&ctl->err, E_HASH_FULL);
    return -1;
}

__global__ void __launch_bounds__(256) k_world_insert_count(DevMap m, DevScan s, const DevFilter* __restrict__ f, DevCtl* ctl, int world_mode) {
    __shared__ WorldState ws;
    __shared__ int s_top, s_last;
    if (world_mode != 2) {
        // one entry per thread (a single thread doing all of it was a ~2 us chain in front of every CTA); same evaluation order as mul():
        // s = a0 b0; s += a1 b1; s += a2 b2
        const int t = threadIdx.x;
        const double* x = f->x;                    // pos3 rot9 rot_ext9 pos_ext3 ...
        if (t < 9) {
            const int i = t / 3, j = t % 3;
            double sacc = x[3 + 3 * i] * x[12 + j];
            sacc += x[3 + 3 * i + 1] * x[15 + j];
            sacc += x[3 + 3 * i + 2] * x[18 + j];
            ws.r_wl.a[t] = sacc;
            ws.mf[i * 4 + j] = (float)sacc;
        } else if (t < 12) {
            const int i = t - 9;
            double sacc = x[3 + 3 * i] * x[21];
            sacc += x[3 + 3 * i + 1] * x[22];
            sacc += x[3 + 3 * i + 2] * x[23];
            const double pw = sacc + x[i];
            ws.p_wl[i] = pw;
            ws.mf[i * 4 + 3] = (float)pw;
        } else if (t >= 32 && t < 41) { const int e = t - 32; ws.Prr.a[e] = f->P[(3 + e / 3) * 23 + 3 + e % 3]; }
        else if (t >= 64 && t < 73) { const int e = t - 64; ws.Ppp.a[e] = f->P[(e / 3) * 23 + e % 3]; }
        __syncthreads();
    }
    const int n = ctl->n;
    const size_t NM = (size_t)s.nmax;
    const int lane0 = threadIdx.x & 31;
    unsigned long long skipped = 0;
    for (int i0 = blockIdx.x * blockDim.x + threadIdx.x - lane0; i0 < n; i0 += gridDim.x * blockDim.x) {    // warp-uniform trip count
        const int i = i0 + lane0;
        unsigned long long pk = KEY_EMPTY;
        bool have = false;
        if (i < n) {
            double w[3];
            if (world_mode != 2) {
                const float x = s.raw[3 * (size_t)i], y = s.raw[3 * (size_t)i + 1], z = s.raw[3 * (size_t)i + 2];
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const float p0 = __fmul_rn(ws.mf[r * 4 + 0], x), p1 = __fmul_rn(ws.mf[r * 4 + 1], y), p2 = __fmul_rn(ws.mf[r * 4 + 2], z);
                    const float o = __fadd_rn(p0, __fadd_rn(p1, __fadd_rn(p2, ws.mf[r * 4 + 3])));
                    w[r] = (double)o;
                    s.pw[3 * (size_t)i + r] = w[r];
                }
                V3 pl; M3 cl;
                if (world_mode == 1) {
                    pl = v3((double)x, (double)y, (double)z);
                    calc_body_cov(pl, s.range_var, s.sn2, cl);
                } else {
                    pl = v3(s.pl[i], s.pl[NM + i], s.pl[2 * NM + i]);
#pragma unroll
                    for (int k = 0; k < 9; k++) cl.a[k] = s.cl[(size_t)k * NM + i];
                }
                const M3 cw = world_cov(ws.r_wl, cl, pl, ws.Prr, ws.Ppp);
#pragma unroll
                for (int k = 0; k < 9; k++) s.pcov[9 * (size_t)i + k] = cw.a[k];
            } else {
                w[0] = s.pw[3 * (size_t)i]; w[1] = s.pw[3 * (size_t)i + 1]; w[2] = s.pw[3 * (size_t)i + 2];
            }
            have = voxel_index(w[0], w[1], w[2], m.voxel_size, pk, m.voxel_inv);
            if (!have) skipped++;                                   // counted skip (see voxel_index), not an error
        }
        // keys are 63-bit, so the all-ones prefix makes the lanes without a voxel distinct from every key and from each other
        const unsigned grp = __match_any_sync(0xffffffffu, have ? pk : (0xFFFFFFFF00000000ull | (unsigned long long)lane0));
        const int leader = __ffs(grp) - 1;
        int slot = -1;
        if (have && lane0 == leader) {
            slot = find_or_insert(m, ctl, pk);
            if (slot >= 0) {
                const int c = atomicAdd(&m.cnt[slot], __popc(grp));
                if (c == 0) m.touched[atomicAdd(&ctl->n_touched, 1)] = slot;
                atomicMin(&m.ft[slot], i);                         // the leader is the group's lowest lane = lowest point index
                atomicMax(&m.lt[slot], i0 + 31 - __clz(grp));
            }
        }
        slot = __shfl_sync(0xffffffffu, slot, leader);
        if (i < n) m.pslot[i] = have ? slot : -1;
    }
    if (skipped) atomicAdd((unsigned long long*)&ctl->st.n_skipped, skipped);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&ctl->cnt_ticket, 1u);
        s_last = (t == gridDim.x - 1) ? 1 : 0;
        if (s_last) { ctl->cnt_ticket = 0; s_top = 0; }
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int V = atomicAdd(&ctl->n_touched, 0);
    const int lane = threadIdx.x & 31;
    constexpr int U = 16;                                          // voxels per thread per round: independent L2 loads
    for (int base = 0; base < V; base += blockDim.x * U) {
        int slot[U], c[U], sum = 0;
#pragma unroll
        for (int u = 0; u < U; u++) { const int idx = base + threadIdx.x * U + u; slot[u] = idx < V ? __ldcg(&m.touched[idx]) : -1; }
#pragma unroll
        for (int u = 0; u < U; u++) { c[u] = slot[u] >= 0 ? __ldcg(&m.cnt[slot[u]]) : 0; sum += c[u]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        int wbase = 0;
        if (lane == 31) wbase = atomicAdd(&s_top, incl);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        int off = wbase + incl - sum;
#pragma unroll
        for (int u = 0; u < U; u++) { if (slot[u] >= 0) m.seg_off[slot[u]] = off; off += c[u]; }
    }
}

__device__ void fill_classify(const DevMap& m, DevCtl* ctl, int vi);      // vmp_fill.cuh

// ------------------------------------------------------------------------- M3b: fill segments
__global__ void __launch_bounds__(1024) k_seg_fill(DevMap m, DevCtl* ctl, int classify) {
    const int n = ctl->n;
    // which touched voxels go to the CTA path of k_fill (vmp_fill.cuh): one voxel per thread, dealt round-robin to the CTAs and done
    // first (a few thousand voxels on the first CTAs alone were the tail of this kernel; the L2 prefetches they issue for k_fill
    // also get the time of the segment pass to land)
    if (classify) {
        const int V = ctl->n_touched;
        for (int vi = blockIdx.x + gridDim.x * threadIdx.x; vi < V; vi += gridDim.x * blockDim.x) fill_classify(m, ctl, vi);
    }
    for (int b = blockIdx.x; b * PT_BLOCK < n; b += gridDim.x) {
        const int i = b * PT_BLOCK + threadIdx.x;
        const int lane = threadIdx.x & 31;
        int is_last = 0;
        int slot = -1;
        if (i < n) slot = m.pslot[i];
        // one cursor update per (warp, voxel) group; the order inside a segment is irrelevant (it is selected / sorted later)
        const unsigned grp = __match_any_sync(0xffffffffu, slot >= 0 ? slot : -1 - lane);
        if (slot >= 0) {
            const int leader = __ffs(grp) - 1;
            int base = 0;
            if (lane == leader) base = m.seg_off[slot] + atomicAdd(&m.cursor[slot], __popc(grp));
            base = __shfl_sync(grp, base, leader);
            m.seg[base + __popc(grp & ((1u << lane) - 1u))] = i;
            is_last = (m.lt[slot] == i);
        }
        const int cl = __syncthreads_count(is_last);
        if (threadIdx.x == 0) m.blk_last[b] = cl;
    }
}

// ------------------------------------------------------------------------- M2: exact LRU eviction
// cache.push_front / splice / "if (cache.size() > capacity) erase(cache.back())" (voxel_map.cpp:242-253).
// The k-th over-capacity creation (at point index t_k) evicts the oldest LRU-log entry that is neither
// stale nor was spliced to the front earlier in this scan (touched with first touch < t_k).
//   fast path (one CTA, parallel): creation times of the new voxels in ascending order (bitmap over the point indices +
//     prefix popcount); rank the valid untouched
//     entries of the log head with block scans: the r-th of them is the victim of eviction r, PROVIDED every
//     touched entry met on the way was indeed touched before the eviction it was examined for - checked
//     in the same pass.
//   slow path (thread 0, serial walk): when that check fails, i.e. a victim is touched again later in the
//     same scan (it is then re-created fresh; its old incarnation stays visible to merge() as a ghost slot).

// serial walk (one thread): creation times ct[0..n_new) ascending
__device__ void lru_serial_walk(const DevMap& m, DevCtl* ctl, const int* ct, int n_live0, int n_new) {
    // re-creation times of victims that are touched again later in the scan, ascending.  In global scratch (act_t is not
    // used before the merge prefilter): when a drive turns back into the oldest part of the map, a single scan can re-create
    // hundreds of the voxels it evicts (a 64-entry local queue overflowed on the C3 city drive)
    int* rq = m.act_t;
    const int rq_cap = m.nmax;
    int size = n_live0, ne = 0, rqh = 0, rqn = 0, ci = 0, created = n_new;     // queue = rq[rqh .. rqn)
    long long h = ctl->log_head;
    const long long tail = ctl->log_tail;
    const int sel = ctl->log_sel;
    while (ci < n_new || rqn > rqh) {
        int t;
        if (rqn > rqh && (ci >= n_new || rq[rqh] < ct[ci])) t = rq[rqh++];
        else t = ct[ci++];
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
            if (rqn >= rq_cap) { atomicOr(&ctl->err, E_QUEUE); break; }
            int q = rqn++;
            const int tt = m.ft[victim];
            while (q > rqh && rq[q - 1] > tt) { rq[q] = rq[q - 1]; q--; }
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

constexpr int LRU_BITMAP_WORDS = 8192;            // creation-time bitmap in shared memory: scans of up to 262 144 points

__global__ void __launch_bounds__(1024) k_lru_evict(DevMap m, DevCtl* ctl) {
    __shared__ unsigned bm[LRU_BITMAP_WORDS];
    __shared__ int sh[34];
    __shared__ int s_bad;
    __shared__ long long s_headpos;
    const int tid = threadIdx.x;
    const int n = ctl->n;
    const int n_live0 = ctl->n_live, n_new = ctl->n_new;
    if (n_live0 + n_new <= m.capacity) {
        if (tid == 0) { ctl->n_live = n_live0 + n_new; ctl->st.n_created = n_new; ctl->n_evict = 0; }
        return;
    }
    const unsigned scan_id = ctl->scan_id;
    if (n > LRU_BITMAP_WORDS * 32) {
        // scans beyond the bitmap: creation times by ordered compaction over the points, then the serial walk
        int base = 0;
        for (int b = 0; b * PT_BLOCK < n; b++) {
            const int i = b * PT_BLOCK + tid;
            int f = 0;
            if (i < n) { const int slot = m.pslot[i]; if (slot >= 0) f = (m.ft[slot] == i && m.born_scan[slot] == scan_id); }
            int total;
            const int r = block_excl_scan(f, &total, sh);
            if (f) m.ct[base + r] = i;
            base += total;
        }
        __syncthreads();
        if (tid == 0) lru_serial_walk(m, ctl, m.ct, n_live0, n_new);
        return;
    }
    // creation times = first touches of the voxels created by k_map_insert, ascending.  They are distinct point indices
    // below n, so "sorting" them is a bitmap over the point indices + a prefix popcount (a shared-memory bitonic sort of up
    // to 8192 keys was 0.19 ms per 200 k-point batch at C4, a quarter of the whole map update)
    const int nw = (n + 31) >> 5;
    for (int q = tid; q < nw; q += blockDim.x) bm[q] = 0u;
    if (tid == 0) { s_bad = 0; s_headpos = ctl->log_head; }
    __syncthreads();
    for (int q = tid; q < n_new; q += blockDim.x) { const int t = m.ft[m.newlist[q]]; atomicOr(&bm[t >> 5], 1u << (t & 31)); }
    __syncthreads();
    {
        const int W = (nw + (int)blockDim.x - 1) / (int)blockDim.x, w0 = tid * W;
        int c = 0;
        for (int k = 0; k < W; k++) if (w0 + k < nw) c += __popc(bm[w0 + k]);
        int total;
        int pos = block_excl_scan(c, &total, sh);
        for (int k = 0; k < W; k++) {
            if (w0 + k >= nw) break;
            for (unsigned word = bm[w0 + k]; word; word &= word - 1) m.ct[pos++] = (w0 + k) * 32 + (__ffs(word) - 1);
        }
    }
    __syncthreads();
    const int* sct = m.ct;                                  // sorted creation times (global scratch, read back through L2)
    const int slack = m.capacity - n_live0;                 // creations that still fit
    const int E = n_new - slack;                            // evictions if no victim is re-created
    const int sel = ctl->log_sel;
    const long long tail = ctl->log_tail;
    // The log is append-only with lazy deletion: a voxel touched in k scans has k entries of which only the last is live, so
    // the walk from the head passes many stale entries per victim (C4: ~50 per victim).  Every thread takes EPT consecutive
    // entries per step (independent loads, one block scan per blockDim.x * EPT entries).
    constexpr int EPT = 8;
    int base_u = 0;
    for (long long pos = ctl->log_head; base_u < E && pos < tail; pos += (long long)blockDim.x * EPT) {
        const long long p0 = pos + (long long)tid * EPT;
        int sl[EPT];
        unsigned long long lst[EPT];
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            const long long p = p0 + e;
            sl[e] = p < tail ? m.log_slot[sel][p] : -1;
            lst[e] = p < tail ? m.log_stamp[sel][p] : 0ull;
        }
        unsigned umask = 0, tmask = 0;
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            if (sl[e] < 0) continue;
            const bool live = m.stamp[sl[e]] == lst[e];
            const bool touched = live && m.cnt[sl[e]] > 0;
            if (live && !touched) umask |= 1u << e;
            if (touched) tmask |= 1u << e;
        }
        int total;
        int r = base_u + block_excl_scan(__popc(umask), &total, sh);       // victim index the thread's first entry is examined for
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            if (r < E) {
                const int t = sct[slack + r];
                if ((umask >> e) & 1u) {
                    m.ev_slot[r] = sl[e]; m.ev_time[r] = t; m.ev_key[r] = m.skey[sl[e]];
                    if (r == E - 1) s_headpos = p0 + e + 1;
                } else if (((tmask >> e) & 1u) && !(m.ft[sl[e]] < t)) {
                    s_bad = 1;                              // this entry would be evicted and re-created later in the scan
                }
            }
            r += (umask >> e) & 1u;
        }
        base_u += total;
        __syncthreads();
    }
    __syncthreads();
    if (s_bad) {
        if (tid == 0) lru_serial_walk(m, ctl, m.ct, n_live0, n_new);
        return;
    }
    if (base_u < E) { if (tid == 0) atomicOr(&ctl->err, E_LRU_EXHAUSTED); return; }
    for (int r = tid; r < E; r += blockDim.x) m.evict_t[m.ev_slot[r]] = m.ev_time[r];
    if (tid == 0) {
        ctl->n_evict = E; ctl->n_live = m.capacity; ctl->log_head = s_headpos;
        ctl->st.n_evicted = E; ctl->st.n_created = n_new;
    }
}

#include "vmp_fill.cuh"

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
        if (threadIdx.x < 32) {                             // last-touch points in the blocks before this one (one warp, strided)
            int sacc = 0;
            for (int q = threadIdx.x; q < b; q += 32) sacc += m.blk_last[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
            if (threadIdx.x == 0) base_sh = sacc;
        }
        const int i = b * PT_BLOCK + threadIdx.x;
        int f = 0, slot = -1;
        if (i < n) { slot = m.pslot[i]; if (slot >= 0) f = (m.lt[slot] == i); }
        int total;
        const int r = block_excl_scan(f, &total, sh);
        if (f) {
            const long long pos = tail + base_sh + r;
            if (pos < m.log_cap) { m.log_slot[sel][pos] = slot; m.log_stamp[sel][pos] = sb + (unsigned long long)i; }
            else atomicOr(&ctl->err, E_LOG_CAP);
            m.stamp[slot] = sb + (unsigned long long)i;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------- M7b: finalize
__global__ void __launch_bounds__(256) k_map_finalize(DevMap m, DevCtl* ctl, MapOut* out) {
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
    // the last CTA to get here closes the update
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&ctl->fin_ticket, 1u);
        if (t == gridDim.x - 1) {
            ctl->fin_ticket = 0;
            __threadfence();
            map_end(m, ctl, out);
        }
    }
}

// ------------------------------------------------------------------------- launcher
int launch_map_update(cudaStream_t st, const DevMap& m, const DevScan& s, const DevFilter* f, DevCtl* ctl, int sm_count, bool build, int world_mode, MapOut* out,
                      const Marker* mk, const SideStream* side) {
    const int gpt = (m.nmax + PT_BLOCK - 1) / PT_BLOCK;               // order-preserving passes: 1024 points / block
    const int gstride = sm_count * 2;               // (3 or 4 CTAs per SM measured the same: 42 us at 200 000 points either way)
    int launches = 0;
    k_world_insert_count<<<gstride, 256, 0, st>>>(m, s, f, ctl, world_mode); launches++; mark(mk, VMP_K_WORLD_POINTS);
    // the LRU eviction (one CTA) is independent of the segment build: side branch of the graph
    const bool fork = side != nullptr && mk == nullptr;
    if (fork) {
        cudaEventRecord(side->ev[0], st); cudaStreamWaitEvent(side->st, side->ev[0], 0);
        k_lru_evict<<<1, 1024, 0, side->st>>>(m, ctl); launches++;
        cudaEventRecord(side->ev[1], side->st);
    }
    k_seg_fill<<<gpt, 1024, 0, st>>>(m, ctl, build ? 0 : 1); launches++; mark(mk, VMP_K_SEG_FILL);
    if (fork) cudaStreamWaitEvent(st, side->ev[1], 0);
    else { k_lru_evict<<<1, 1024, 0, st>>>(m, ctl); launches++; mark(mk, VMP_K_LRU_EVICT); }
    if (build) { k_fill_build<<<sm_count * 4, 128, 0, st>>>(m, s, ctl); launches++; mark(mk, VMP_K_MAP_FILL); }
    else {
        // resident warps take voxels dynamically; shared memory per CTA = FILL_WARPS x fill_warp_bytes (opt-in size, set at create)
        // the heavy voxels (CTA path) and the rest (a warp per voxel) are two launches of k_fill on two streams: different shared-memory
        // footprints, disjoint voxels
        const size_t smem = FILL_WARPS * fill_warp_bytes(m.maxpt), smem_h = fill_heavy_bytes(m.maxpt);
        auto per_sm = [](size_t b) { const int q = (int)((227 * 1024) / (b + 1024)); return q < 1 ? 1 : q > 8 ? 8 : q; };
        if (fork) {
            cudaEventRecord(side->ev[4], st); cudaStreamWaitEvent(side->st, side->ev[4], 0);
            k_fill<true><<<sm_count * per_sm(smem_h), HEAVY_WARPS * 32, smem_h, side->st>>>(m, s, ctl); launches++;
            cudaEventRecord(side->ev[5], side->st);
        } else {
            k_fill<true><<<sm_count * per_sm(smem_h), HEAVY_WARPS * 32, smem_h, st>>>(m, s, ctl); launches++; mark(mk, VMP_K_FILL_ACC);
        }
        k_fill<false><<<sm_count * per_sm(smem), FILL_WARPS * 32, smem, st>>>(m, s, ctl); launches++; mark(mk, VMP_K_MAP_FILL);
        if (fork) cudaStreamWaitEvent(st, side->ev[5], 0);
    }
    // the LRU-log append only needs the last-touch times: side branch next to the merge simulation
    if (fork && !build) {
        cudaEventRecord(side->ev[2], st); cudaStreamWaitEvent(side->st, side->ev[2], 0);
        k_log_append<<<gpt, 1024, 0, side->st>>>(m, ctl); launches++;
        cudaEventRecord(side->ev[3], side->st);
    }
    if (!build) {
        k_merge_prefilter<<<sm_count * 4, 128, 0, st>>>(m, ctl); launches++; mark(mk, VMP_K_MERGE_PREFILTER);
        {   // one thread-block cluster (vmp_merge.cuh)
            cudaLaunchConfig_t lc{};
            lc.gridDim = dim3(MERGE_CLUSTER); lc.blockDim = dim3(512); lc.dynamicSmemBytes = sizeof(MergeShared); lc.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = MERGE_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            lc.attrs = at; lc.numAttrs = 1;
            cudaLaunchKernelEx(&lc, k_merge_rounds, m, ctl);
        }
        launches++; mark(mk, VMP_K_MERGE_SERIAL);
    }
    if (fork && !build) cudaStreamWaitEvent(st, side->ev[3], 0);
    else { k_log_append<<<gpt, 1024, 0, st>>>(m, ctl); launches++; mark(mk, VMP_K_LOG_APPEND); }
    k_map_finalize<<<sm_count, 256, 0, st>>>(m, ctl, out); launches++; mark(mk, VMP_K_MAP_FINALIZE);
    return launches;
}

// opt-in shared memory size of k_fill (once per handle)
cudaError_t map_configure_kernels(const DevMap& m) {
    cudaError_t e = cudaFuncSetAttribute(k_merge_rounds, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MergeShared));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_fill<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fill_heavy_bytes(m.maxpt));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_fill<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(FILL_WARPS * fill_warp_bytes(m.maxpt)));
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
