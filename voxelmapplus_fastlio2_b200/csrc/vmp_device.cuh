// vmp_device.cuh — device-resident data layout of one trajectory (map + filter + scan).
//
// Everything lives in HBM for the lifetime of the handle; nothing is re-uploaded per
// scan except the raw scan (N x 12 B) and, when the caller keeps the filter on the host,
// the state/P (4.5 KB).  Layout:
//
//   voxel hash      open addressing, linear probing; tkey[] packed 3x21-bit voxel
//                   coordinate (uint64), tval[] slot id.  Replaces
//                   std::unordered_map<VoxelKey, shared_ptr<VoxelGrid>> (voxel_map.h:105).
//   voxel slots     SoA indexed by slot id (pool = map_capacity + max_points_per_scan + slack):
//                     hot[slot][8]   64-byte record {mean xyz, norm xyz, (flags,n), pad}: the
//                                    only thing the IEKF measurement kernel gathers per point
//                     ppt[slot][6]   sum p p^T (exactly symmetric -> 6 values)
//                     cov[slot][36]  Plane::cov (6x6, accumulated, not exactly symmetric)
//                     tp[slot][12][MAXPT] stored points of a filling voxel, component-major
//                     skey, sgroup, stamp (LRU), n_temp, newly, born/full bookkeeping
//   per-scan scratch  cnt/cursor/ft/lt/seg_off per slot, tpos/pslot/seg per point
//   LRU log         append-only (slot, stamp) pairs in last-touch order with lazy deletion;
//                   replaces std::list<VoxelKey> cache (voxel_map.h:126)
//   scan buffers    persistent lio::ResidualData (voxel_map.h:53-65) as SoA
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include "vmp_math.cuh"

namespace vmp {

// ---- flags (same bit values as VMP_F_* in include/vmp_b200.h) ----
constexpr uint32_t F_INIT = 1u, F_PLANE = 2u, F_UE = 4u, F_MERGED = 8u;

constexpr unsigned long long KEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned long long KEY_TOMB = 0xFFFFFFFFFFFFFFFEull;
constexpr int T_INF = INT_MAX;          // "never" for per-scan times (point indices)
constexpr unsigned SCAN_NEVER = 0xFFFFFFFFu;

// error bits reported through DevCtl::err
constexpr int E_KEY_RANGE = 1;          // (no longer raised: points outside +-2^20 voxels / non-finite points are counted skips, n_skipped)
constexpr int E_POOL = 2;               // slot pool exhausted
constexpr int E_LRU_EXHAUSTED = 4;      // eviction would have to take a voxel touched in the same scan
constexpr int E_REFIT_OVERFLOW = 8;     // refit of a voxel holding more than max_point_thresh points (build overflow + thresh 1)
constexpr int E_QUEUE = 16;             // internal queue overflow in the serial merge / eviction kernels
constexpr int E_HASH_FULL = 32;
constexpr int E_MERGE_DEPTH = 64;       // the merge phase had to be redone serially AND its undo log had overflowed (> 16384 modifications in one scan)
constexpr int E_MERGE_CAP = 128;        // (unused: a large active set runs in the exact serial mode)
constexpr int E_LOG_CAP = 256;          // LRU log full (compaction was not served in time)
constexpr int E_FILL_CAP = 512;         // (unused since the refits run out of shared memory)
constexpr int E_UPLOAD = 1024;          // a chunk of a streamed scan upload did not arrive within 2 s (the host side of the call failed half-way)

struct DevStats {                       // == vmp_update_stats order
    long long n_points, n_ins, n_touch, n_created, n_refit, refit_points, n_full, n_mergeprobe, n_merge, n_evicted, map_size, n_mergevox, n_skipped;
};

struct DevCtl {
    // scan
    int n;                              // points in the current scan
    unsigned scan_id;                   // increments per map update
    unsigned long long stamp_base;      // stamp of point i of this scan = stamp_base + i (>= 1)
    // map bookkeeping
    int n_live;                         // live voxels
    int n_touched, n_new, n_evict, n_hot, n_ghost;
    int free_top;                       // free-slot stack height
    int tombstones;
    int need_rehash;
    int need_log_compact;
    int log_sel;                        // which of the two log buffers is current
    long long log_head, log_tail;       // positions in the current log buffer
    unsigned long long group_counter;
    int err;
    int pad0;
    unsigned ticket;                    // arrival counter of k_measure's measurement CTAs (the solver CTA waits on it)
    unsigned long long iter_pub;        // k_iekf_loop: 2 * (8 * scan sequence number + executed iterations) + stop flag, released by the solver CTA
    int fill_next;                      // next touched voxel to be handed to a warp of k_fill
    int n_undo;                         // entries of the merge undo log of this update
    int n_heavy, heavy_next;            // voxels of this update that go to the CTA path of k_fill, and the next one to be handed out
    int dbg_it;                         // which IEKF iteration's solver phase cycles go to dbg[3..6]
    int dbg[8];                         // debug counters of the last map update: [0] active set after the prefilter, [1] merge events simulated, [2] re-examinations that activated a voxel
    DevStats st;
    // IEKF
    int iter;                           // executed iterations
    int done;                           // loop left (eps test or max_iter)
    int converged;
    int effect[8];
    int max_iter;
    unsigned long long seq;             // sequence number of the current scan (echoed in StateOut / MapOut)
    unsigned fin_ticket;                // arrival counter of k_map_finalize's CTAs (the last one closes the update)
    unsigned cnt_ticket;                // arrival counter of k_map_count's CTAs (the last one lays out the segments)
    unsigned long long up_pub;          // streamed upload (vmp_scan with a pageable pointer): 64 * scan sequence number + chunks of points that have
                                        // arrived after the first one; written IN STREAM ORDER behind each chunk's DMA copy, read by the first measurement pass
};

// per-update counters (single thread).  Called at the END of every update (map_end) so that the next one starts clean without a
// "begin" kernel; n_evict / the eviction list stay readable until then (vmp_dump_evicted)
__device__ __forceinline__ void map_counters_reset(DevCtl* ctl) {
    DevStats z = {};
    ctl->st = z;
    ctl->n_touched = 0; ctl->n_new = 0; ctl->n_hot = 0; ctl->n_ghost = 0;
    ctl->fill_next = 0; ctl->n_heavy = 0; ctl->heavy_next = 0; ctl->n_undo = 0;
    for (int q = 0; q < 3; q++) ctl->dbg[q] = 0;          // [3..7] belong to the solver CTA
}

// ---- per-scan mailboxes.  ScanIn travels with the points in one DMA copy; StateOut / MapOut live in PINNED, MAPPED
//      host memory and are written by the kernels directly (posted PCIe writes), so no copy-back operation is needed ----
constexpr int SCAN_STATE_HDR = 1;       // ScanIn::x / P hold the prior (vmp_scan)
constexpr int SCAN_STATE_DEV = 2;       // ScanIn::prior points at 36 + 529 doubles in device memory (vmp_scan_dev)
constexpr int SCAN_BEGIN_UPDATE = 4;    // start of IESKF::update: predict_x = x_, iteration counters (ieskf.cpp:127-130)
constexpr int SCAN_PREDICT = 8;         // the prior and the IMU poses are produced on the device (k_predict) from the filter state resident there
struct ScanIn {
    const float* pts;                   // n points, `stride` floats apart (3: xyz, 4: xyz + time offset), device memory
    const double* prior;
    unsigned long long seq;
    int n, mode;
    int n_poses, stride;                // raw scans (vmp_scan_raw): IMU poses for the motion compensation follow the header
    int gate_pts, pad_gate;             // > 0: the points arrive in chunks of gate_pts (a multiple of 32) WHILE the first measurement pass runs: point i may be
                                        // read once DevCtl::up_pub >= 64 * seq + i / gate_pts (chunk 0 is there before the launch)
    double x[36];
    double P[529];
};
// lio::Pose (commons.h:30-43) as uploaded for the motion compensation: 22 doubles
struct DevPose { double offset, acc[3], gyro[3], vel[3], pos[3], rot[9]; };
constexpr int MAX_POSES = 64;
struct StateOut {                       // written by the solver CTA when the IEKF loop ends
    double x[36];
    double P[529];
    int iter, converged, effect[8];
    unsigned long long seq;             // written last, after a system-scope fence
};
struct MapOut {                         // written by the last CTA of k_map_finalize
    DevStats st;
    int err, need_maint;                // need_maint: bit0 rehash, bit1 LRU-log compaction
    int dbg[8];
    unsigned long long seq;
};

struct DevFilter {
    double x[36];                       // vmp_state layout: pos3 rot9 rot_ext9 pos_ext3 vel3 bg3 ba3 g3
    double xpred[36];
    double P[529];
    double last[6];                     // LIODataGroup::last_acc / last_gyro (lio_builder.h:48-49), kept across scans by k_predict
};
// one propagation step of LIOBuilder::undistortCloud's IMU loop (lio_builder.cpp:89-111): averaged, rescaled IMU input, step length,
// and the time offset of the pose it ends at (-1e300: the closing step to the end of the scan, :113-114, which records no pose)
struct DevImuStep { double acc[3], gyro[3], dt, offset; };
struct DevPredictIn {
    DevImuStep steps[MAX_POSES];
    double Q[144];                      // process noise (lio_builder.cpp:9-12)
    double last[6];                     // last_acc / last_gyro to start from when use_last != 0 (first device-propagated scan)
    int n_steps, use_last;
};

struct DevMap {
    // hash
    unsigned long long* tkey;
    int* tval;
    unsigned hmask;
    // parameters
    int pool, maxpt, upt, capacity;
    int merge_cap;                      // active-set size up to which the merge phase runs its parallel rounds (<= MERGE_CAP); beyond: serial mode
    int merge_r;                        // conflict radius of the merge rounds (MERGE_R = 4; test knob VMP_MERGE_R = 3 is the exact radius and makes late activations happen)
    int merge_max_depth;                // test knob (VMP_MERGE_MAX_DEPTH): a merge deeper than this in a cascade sends the scan to the serial redo
    int* undo_slot; double* undo_rec; int undo_cap;      // undo log of the merge phase (vmp_merge.cuh)
    double voxel_inv;                   // 1 / voxel_size when voxel_size is a power of two (exact), else 0 (voxel_index divides)
    int heavy_points;                   // k_fill: voxels whose refits of a scan loop over at least this many stored points take the CTA path (0: never)
    double plane_thresh, voxel_size, th_angle, th_dist;
    // slots
    double* hot;                        // [pool][8]
    double* ppt;                        // [pool][6]
    double* cov;                        // [pool][36]
    double* center;                     // [pool][3]
    double* tp;                         // [pool][12][maxpt]
    unsigned long long* skey;
    unsigned long long* sgroup;
    unsigned long long* stamp;
    int* n_temp;
    int* newly;
    unsigned* born_scan;
    unsigned* full_scan;
    int* full_idx;
    // per-slot per-scan scratch
    int* cnt; int* cursor; int* ft; int* lt; int* seg_off; int* evict_t; int* ghost; int* evn;
    // free list
    int* free_slots;
    // per-point per-scan
    int* pslot; int* seg;
    // lists
    int* touched;                       // [nmax]
    int* newlist;                       // [nmax] slots created by this map update
    int* hotlist;                       // [nmax] voxels for the CTA path of k_fill
    int* vox_cls;                       // [nmax] per touched voxel: 1 = CTA path
    int* ev_slot; int* ev_time; unsigned long long* ev_key;   // eviction list [nmax]
    int* ct;                            // creation times (sorted) [nmax]
    int* blk_last; int* blk_new;        // per 1024-point block counts
    int* act_slot; int* act_t;          // serial merge active set [nmax]
    // LRU log (two buffers for compaction)
    int* log_slot[2];
    unsigned long long* log_stamp[2];
    long long log_cap;
    int* log_blk;                       // block counts for compaction
    int nmax;
};

struct DevScan {                        // persistent residual buffer, SoA over point index
    int nmax;
    double* pl;                         // [3][nmax] point_lidar (after calcBodyCov's z edit)
    double* cl;                         // [9][nmax] cov_lidar
    double* rnorm;                      // [3][nmax] plane_norm (persistent, Q2)
    double* rmean;                      // [3][nmax] plane_mean
    double* rres;                       // [nmax] residual
    uint8_t* rvalid;                    // [nmax] is_valid
    uint8_t* rstatus;                   // [nmax] bit0 found bit1 plane bit2 valid
    unsigned long long* rkey;           // [nmax] packed key of point_world
    int* rslot;                         // [nmax] voxel slot found for rkey (valid between the iterations of one scan: the map does not change)
    double* pw;                         // [nmax][3] pv.point  (float32 world widened), AoS: gathered per voxel
    double* pcov;                       // [nmax][9] pv.cov
    float* raw;                         // [nmax][3] staged raw scan
    double range_var, sn2;
};

// scratch of the pcl::VoxelGrid downsample (vmp_downsample.cu), sized for max_points_per_scan
struct DevDown {
    unsigned* mm;                       // [8] order-encoded bounding box + finite count (zero between scans)
    unsigned* keys[2]; int* vals[2];    // (leaf idx, point index) before / after the stable sort
    int* head; int* rank;               // leaf starts of the sorted list and their inclusive scan
    float4* out;                        // filtered cloud (centroid xyz, mean curvature), ascending leaf idx
    int* m;                             // number of leaves
    void* temp; size_t temp_bytes;      // CUB scratch
};

// scratch of the device-side map read-back (vmp_readback.cu), allocated by the first vmp_dump_map
struct DevDump {
    unsigned long long* keys[2]; int* vals[2];      // (LRU stamp, slot) of the live voxels before / after the sort
    int* count;
    unsigned long long* records;                    // [pool][61] assembled vmp_plane records
    void* temp; size_t temp_bytes;
};

// ---- key packing / hashing -----------------------------------------------------------
__host__ __device__ __forceinline__ bool key_in_range(long long k) { return k >= -(1ll << 20) && k < (1ll << 20); }
__host__ __device__ __forceinline__ unsigned long long pack_key(long long x, long long y, long long z) {
    return ((unsigned long long)(x + (1ll << 20)) << 42) | ((unsigned long long)(y + (1ll << 20)) << 21) |
           (unsigned long long)(z + (1ll << 20));
}
__host__ __device__ __forceinline__ void unpack_key(unsigned long long k, long long& x, long long& y, long long& z) {
    x = (long long)((k >> 42) & 0x1FFFFF) - (1ll << 20);
    y = (long long)((k >> 21) & 0x1FFFFF) - (1ll << 20);
    z = (long long)(k & 0x1FFFFF) - (1ll << 20);
}
// any hash is allowed: bucket order is never observable in the reference
__host__ __device__ __forceinline__ unsigned hash_key(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned)k;
}

// VoxelMap::index (voxel_map.cpp:194-198): true division, floor, cast
// Points whose voxel coordinate cannot be packed (|k| >= 2^20, i.e. 262 km at 0.25 m voxels) and non-finite points have no
// voxel here: the filter treats them like "voxel not in the map" (Q2), the map update skips and counts them (the reference
// would give them isolated far-away voxels; (long long)floor(NaN) is INT64_MIN on x86 but 0 on the device, hence the
// explicit range test on the floating-point value).
// inv != 0: voxel_size is a power of two (0.5 m, 0.25 m ... the BASELINE configurations) and inv = 1 / voxel_size, itself a power of
// two: x * inv IS x / voxel_size, bit for bit (both are exact scalings), and three division sequences per point and iteration - a tenth
// of the measurement pass's instructions - become three multiplications.  Any other voxel size takes the division (inv = 0).
__device__ __forceinline__ bool voxel_index(double x, double y, double z, double vs, unsigned long long& pk, double inv = 0.0) {
    double fx, fy, fz;
    if (inv != 0.0) { fx = floor(x * inv); fy = floor(y * inv); fz = floor(z * inv); }
    else { fx = floor(x / vs); fy = floor(y / vs); fz = floor(z / vs); }
    const double lim = 1048576.0;
    if (!(fx >= -lim && fx < lim && fy >= -lim && fy < lim && fz >= -lim && fz < lim)) { pk = KEY_EMPTY; return false; }   // false for NaN
    const long long kx = (long long)fx, ky = (long long)fy, kz = (long long)fz;
    pk = pack_key(kx, ky, kz);
    return true;
}

// featmap.find(k): slot id or -1.  Linear probing examines consecutive entries, so a window of four keys and values is
// fetched with independent loads and scanned in probe order: one memory latency per four probes (every caller is a
// chain of dependent look-ups, and probes past tombstones / to the terminating EMPTY are common).
__device__ __forceinline__ int hash_find(const DevMap& m, unsigned long long pk) {
    unsigned h = hash_key(pk) & m.hmask;
    for (unsigned probe = 0; probe <= m.hmask; probe += 4) {
        unsigned long long k[4];
        int v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { const unsigned hq = (h + q) & m.hmask; k[q] = m.tkey[hq]; v[q] = m.tval[hq]; }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (k[q] == pk) return v[q];
            if (k[q] == KEY_EMPTY) return -1;
        }
        h = (h + 4) & m.hmask;
    }
    return -1;
}

// hot record access
__device__ __forceinline__ void hot_get_fn(const double* hot, int slot, uint32_t& flags, int& n) {
    const long long w = __double_as_longlong(hot[(size_t)slot * 8 + 6]);
    flags = (uint32_t)(w & 0xFFFFFFFFll);
    n = (int)(w >> 32);
}
__device__ __forceinline__ void hot_set_fn(double* hot, int slot, uint32_t flags, int n) {
    const long long w = ((long long)n << 32) | (long long)flags;
    hot[(size_t)slot * 8 + 6] = __longlong_as_double(w);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

#define VMP_CUDA_CHECK(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) { vmp::set_error("%s failed: %s", #expr, cudaGetErrorString(_e)); return VMP_ERR_CUDA; } \
    } while (0)

void set_error(const char* fmt, ...);

}  // namespace vmp
