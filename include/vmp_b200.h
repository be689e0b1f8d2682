/*
 * vmp_b200.h — C ABI of the B200-native voxel_plus hot path.
 *
 * One handle = one trajectory = one voxel map + one filter state, resident on ONE
 * B200, driven by one CUDA stream / one CUDA graph per scan.  Handles are not
 * thread-safe; different handles are independent (no process-wide statics, unlike
 * the reference's VoxelGrid::merge_thresh_* / VoxelGrid::count, voxel_map.cpp:6-9).
 *
 * Every entry point names the reference interface it replaces
 * (paths relative to the reference repo, voxel_plus/src/map_builder/...).
 * All pointers are HOST pointers unless the name says `_dev`.  All matrices are
 * row-major fp64.  Return value: 0 = ok, negative = vmp_status; the message of the
 * last failure on this thread is available through vmp_last_error().
 * No C++ types, no torch types, no exceptions cross this boundary.
 */
#ifndef VMP_B200_H
#define VMP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum vmp_status {
    VMP_OK = 0,
    VMP_ERR_INVALID_ARG = -1,   /* null pointer, n < 0, n > max_points_per_scan, bad config */
    VMP_ERR_NO_DEVICE = -2,     /* no CUDA device / wrong architecture: there is NO CPU fallback */
    VMP_ERR_CUDA = -3,          /* a CUDA runtime call failed; see vmp_last_error() */
    VMP_ERR_CAPACITY = -4,      /* slot pool / hash / LRU log exhausted, map_capacity below one scan's voxels.  Reported by the update that hit it;
                                   error bits are per update (a later update starts clean).  Points that cannot be keyed (non-finite, or voxel
                                   coordinate outside +-2^20) are NOT errors: they are skipped and counted in vmp_update_stats.n_skipped */
    VMP_ERR_STATE = -5          /* call sequence error (e.g. scan before build) */
} vmp_status;

/* mirrors lio::LIOConfig (lio_builder.h:17-42) field for field; the last block is
 * what a device implementation additionally has to know. */
typedef struct vmp_config {
    int    opti_max_iter;            /* 5. The reference reads it but never applies it (always 5, ieskf.h:101); we honour it. */
    double na, ng, nba, nbg;         /* 0.01 0.01 1e-4 1e-4 */
    int    imu_init_num;             /* 20 */
    double r_il[9];                  /* identity */
    double p_il[3];                  /* zero */
    int    gravity_align;            /* 1 */
    int    estimate_ext;             /* 0 */
    double scan_resolution;          /* 0.1; <=0 means "no downsample" (lio_builder.cpp:215-223); > 0: pcl::VoxelGrid leaf size,
                                        applied by vmp_scan_raw / vmp_lio_process on the device (vmp_downsample) */
    double voxel_size;               /* 0.5 */
    int    update_size_thresh;       /* 10 */
    int    max_point_thresh;         /* 100 */
    double plane_thresh;             /* 0.01 */
    double ranging_cov;              /* 0.04 */
    double angle_cov;                /* 0.1 */
    double merge_thresh_for_angle;   /* 0.1 */
    double merge_thresh_for_distance;/* 0.04 */
    int    map_capacity;             /* 100000 */
    /* ---- device-side additions ---- */
    int    max_points_per_scan;      /* size of the persistent residual buffer (reference: hard 10000, lio_builder.cpp:25); must cover RAW scans too:
                                        the first scan and the input of the scan filter are staged unfiltered */
    int    device;                   /* CUDA device ordinal */
} vmp_config;

/* mirrors kf::State (ieskf.h:31-62); rot / rot_ext row-major 3x3 */
typedef struct vmp_state {
    double pos[3];
    double rot[9];
    double rot_ext[9];
    double pos_ext[3];
    double vel[3];
    double bg[3];
    double ba[3];
    double g[3];
} vmp_state;

/* mirrors lio::Plane (voxel_map.h:43-51) + the VoxelGrid fields that are observable
 * from outside (voxel_map.h:84-100; utils.cpp:161-195) */
#define VMP_F_INIT          1u   /* VoxelGrid::is_init        */
#define VMP_F_PLANE         2u   /* VoxelGrid::is_plane       */
#define VMP_F_UPDATE_ENABLE 4u   /* VoxelGrid::update_enable  */
#define VMP_F_MERGED        8u   /* VoxelGrid::merged         */
typedef struct vmp_plane {
    int64_t  key[3];             /* VoxelGrid::position */
    double   mean[3];
    double   ppt[9];
    double   norm[3];
    double   cov[36];
    double   center[3];
    int32_t  n;                  /* Plane::n */
    int32_t  n_temp;             /* temp_points.size() */
    int32_t  newly_add_point;
    uint32_t flags;
    uint64_t group;              /* group_id: only equality between voxels is meaningful (Q22) */
    uint64_t lru_rank;           /* 0 = front of VoxelMap::cache (most recently inserted-into) */
} vmp_plane;

/* counters of one map update: the terms of the algorithmic-byte model */
typedef struct vmp_update_stats {
    int64_t n_points;      /* points offered                              */
    int64_t n_ins;         /* points appended to a filling voxel          */
    int64_t n_touch;       /* distinct voxels touched                     */
    int64_t n_created;     /* voxels created (incl. re-creations)         */
    int64_t n_refit;       /* updatePlane() calls that reached the eigen solve */
    int64_t refit_points;  /* sum over those of the number of stored points looped */
    int64_t n_full;        /* points landing in full voxels (merge() or nothing)   */
    int64_t n_mergeprobe;  /* merge() invocations                         */
    int64_t n_merge;       /* successful pair merges                      */
    int64_t n_evicted;     /* LRU victims                                 */
    int64_t map_size;      /* live voxels afterwards                      */
    int64_t n_mergevox;    /* distinct full plane voxels for which merge() ran (the N_mergeprobe of the byte model) */
    int64_t n_skipped;     /* points NOT inserted: non-finite, or voxel coordinate outside +-2^20 (the reference would create far-away voxels) */
} vmp_update_stats;

typedef struct vmp_scan_stats {
    int32_t iters;              /* IEKF iterations executed (<= opti_max_iter)          */
    int32_t effect_num[8];      /* valid correspondences per executed iteration         */
    int32_t converged;          /* 1 if the loop left through the eps test (ieskf.cpp:148) */
    vmp_update_stats map;
    float   gpu_ms;             /* device time of the call: CUDA events on the launching stream around copies + graph */
    float   host_ms;            /* wall clock of the call measured inside the library (entry to return)           */
} vmp_scan_stats;

/* kernel classes of one scan, for vmp_profile_* (order = launch order inside a scan) */
typedef enum vmp_kernel_id {
    VMP_K_SCAN_IN = 0, VMP_K_SET_SCAN, VMP_K_UPDATE_BEGIN, VMP_K_MEASURE, VMP_K_SOLVE, VMP_K_WORLD_POINTS,
    VMP_K_MAP_BEGIN, VMP_K_MAP_INSERT, VMP_K_MAP_COUNT, VMP_K_SEG_SCAN, VMP_K_SEG_FILL, VMP_K_LRU_EVICT,
    VMP_K_MAP_FILL, VMP_K_MERGE_PREFILTER, VMP_K_MERGE_SERIAL, VMP_K_LOG_APPEND, VMP_K_MAP_FINALIZE,
    VMP_K_MAP_END, VMP_K_REHASH, VMP_K_LOG_COMPACT, VMP_K_SCAN_OUT, VMP_K_FILL_REFIT, VMP_K_FILL_ACC, VMP_K_UNDISTORT, VMP_K_DOWNSAMPLE, VMP_K_COUNT
} vmp_kernel_id;

typedef struct vmp_handle_t* vmp_handle;

/* thread-local description of the last failure */
const char* vmp_last_error(void);
/* library / build identification, e.g. "vmp_b200 0.1 sm_100a" */
const char* vmp_version(void);

/* lio::LIOConfig default member initialisers (lio_builder.h:19-41) */
void vmp_config_default(vmp_config* cfg);

/* LIOBuilder::loadConfig (lio_builder.cpp:5-26): builds the VoxelMap, sizes the residual buffer */
int vmp_create(const vmp_config* cfg, vmp_handle* out);
int vmp_destroy(vmp_handle h);

/* VoxelMap::build (voxel_map.cpp:200-230).  pts_world N x 3, cov N x 9 (PointWithCov, voxel_map.h:36-41) */
int vmp_map_build(vmp_handle h, const double* pts_world, const double* cov, int n, vmp_update_stats* stats);
/* VoxelMap::update (voxel_map.cpp:232-256) */
int vmp_map_update(vmp_handle h, const double* pts_world, const double* cov, int n, vmp_update_stats* stats);

/* One call of the measurement plug-in kf::measure_func = LIOBuilder::sharedUpdateFunc
 * (ieskf.h:74, lio_builder.cpp:250-311) on the scan last given to vmp_set_scan().
 * P_prior is the full 23x23 kf.P(); only its [0:6,0:6] corner is read (lio_builder.cpp:266-267).
 * H 12x12, b 12 as kf::SharedState (ieskf.h:22-29). */
int vmp_measure(vmp_handle h, const vmp_state* x, const double* P_prior,
                double* H, double* b, int* effect_num);

/* lio_builder.cpp:224-229: copies the scan (N x 3 float32 lidar-frame xyz) into the
 * persistent residual buffer and evaluates calcBodyCov (commons.cpp:18-45) per point. */
int vmp_set_scan(vmp_handle h, const float* pts_lidar, int n);

/* The whole timed region lio_builder.cpp:224-246 as ONE CUDA graph launch:
 * set_scan + IESKF::update (ieskf.cpp:125-156, all iterations, measurement model
 * on device) + lidarToWorld/pv_list (lio_builder.cpp:231-245) + VoxelMap::update.
 * x / P: prior in, posterior out. */
int vmp_scan(vmp_handle h, vmp_state* x_inout, double* P_inout,
             const float* pts_lidar, int n, vmp_scan_stats* stats);

/* vmp_scan without the intermediate copy: vmp_scan_buffer returns the handle's PINNED staging area (room for
 * max_points_per_scan x 3 float32); the caller writes the filter input there — in the loop that extracts x y z from its own
 * point type, e.g. pcl::PointXYZINormal in lio_builder.cpp:224-229 — and vmp_scan_staged uploads it from there (one DMA copy
 * with the header and the prior) and runs the scan.  The buffer may be refilled as soon as the call has returned
 * (also in pipelined mode). */
float* vmp_scan_buffer(vmp_handle h);
/* That extraction loop done by the handle's staging helpers: x y z of n records that lie stride_floats apart at src (>= 3; 4 for
 * x y z t, 12 for the 48-byte pcl::PointXYZINormal read as floats) go into the staging area; vmp_scan_staged follows.
 * (One host core moves a 200 000-point scan in ~0.3 ms - as long as the device side of the scan.) */
int vmp_scan_buffer_fill(vmp_handle h, const float* src, int stride_floats, int n);
int vmp_scan_staged(vmp_handle h, vmp_state* x_inout, double* P_inout, int n, vmp_scan_stats* stats);

/* Same work with every input already resident in device memory: pts_lidar_dev = device
 * pointer to N x 3 float32; prior_dev = device pointer to 36 + 529 doubles (vmp_state
 * followed by the 23x23 P), or NULL to continue from the state left on the device by the
 * previous call.  The posterior stays on the device (read it with vmp_get_state).  Used
 * for the "inputs resident in HBM" throughput figure. */
int vmp_scan_dev(vmp_handle h, const float* pts_lidar_dev, const double* prior_dev, int n, vmp_scan_stats* stats);

/* Pipelined mode (off by default).  The reference publishes the pose of a scan and then updates the map
 * (lio_builder.cpp:224-246 is one synchronous block only because it is single-threaded).  With pipelining on,
 * vmp_scan / vmp_scan_dev return as soon as the IEKF posterior of the scan is out; the map update of that scan keeps
 * running on the device while the caller prepares the next scan (IMU propagation, undistortion).  Results are
 * identical; what changes is when they are reported: stats->map and stats->gpu_ms describe the PREVIOUS scan, and a
 * capacity error raised by a map update is returned by the next call on the handle.  Every other entry point (and
 * vmp_sync) first waits for the pending map update. */
int vmp_set_pipelined(vmp_handle h, int on);
int vmp_sync(vmp_handle h);

/* lio::Pose (commons.h:30-43): IMU pose at `offset` seconds after the start of the scan, world-frame acceleration and
 * body rate valid up to that pose; produced by the IMU propagation loop of LIOBuilder::undistortCloud (lio_builder.cpp:78-112). */
typedef struct vmp_pose { double offset; double acc[3], gyro[3], vel[3], pos[3], rot[9]; } vmp_pose;
/* the motion compensation of LIOBuilder::undistortCloud (lio_builder.cpp:75, 127-152) in front of
 * vmp_scan, on the device, in the same graph.  cloud_xyzt: N x 4 float32 (x, y, z, time offset in ms = the `curvature`
 * field of pcl::PointXYZINormal), edited in place like the reference edits package.cloud (sorted by time, compensated to
 * the end-of-scan pose = x, the propagated prior).  poses: 2 <= n_poses <= 64.  Then exactly vmp_scan. */
int vmp_scan_raw(vmp_handle h, vmp_state* x_inout, double* P_inout, float* cloud_xyzt, int n,
                 const vmp_pose* poses, int n_poses, vmp_scan_stats* stats);
/* the IMU propagation IESKF::predict (ieskf.cpp:101-123, called at lio_builder.cpp:106,116) on the device as
 * well.  The filter state and covariance stay RESIDENT on the device from scan to scan (after vmp_first_scan / a previous scan);
 * the host only supplies the IMU input of every propagation step of LIOBuilder::undistortCloud (lio_builder.cpp:89-114): averaged
 * gyro / acceleration (the latter rescaled by 9.81 / gravity_norm, :96), the step length dt, and the time offset of the IMU pose the
 * step ends at (tail.timestamp - cloud_start_time; VMP_NO_POSE for the closing step to the end of the scan, which records no pose).
 * Q: 12 x 12 process noise (lio_builder.cpp:9-12).  last_acc_gyro: LIODataGroup::last_acc, last_gyro (6 doubles) for the FIRST
 * device-propagated scan (they come from the host propagation of the MAP_INIT scan); NULL afterwards (the device keeps them).
 * Then exactly vmp_scan_raw: motion compensation, update; x_out / P_out receive the posterior. */
#define VMP_NO_POSE (-1.0e300)
typedef struct vmp_imu_step { double acc[3], gyro[3], dt, offset; } vmp_imu_step;
int vmp_scan_raw_predict(vmp_handle h, vmp_state* x_out, double* P_out, float* cloud_xyzt, int n,
                         const vmp_imu_step* steps, int n_steps, const double* Q, const double* last_acc_gyro, vmp_scan_stats* stats);
/* The copy of the compensated cloud back into the caller's buffer (the reference edits package.cloud in place,
 * lio_builder.cpp:145-147) is on by default; on = 0 leaves the caller's buffer untouched - the compensated (and, with
 * scan_resolution > 0, filtered) cloud of the last raw scan is always available through vmp_get_lidar_cloud, which is what
 * the reference's consumers read (LIOBuilder::lidar_cloud, lio_node.cpp:185-228). */
int vmp_set_raw_writeback(vmp_handle h, int on);
/* SURVEY.md 8(f) row 1, second half: scan_filter.filter() = pcl::VoxelGrid<PointXYZINormal>::filter with leaf size
 * (leaf, leaf, leaf) (lio_builder.cpp:13-14, 215-219) on the device.  cloud_xyzc: N x 4 float32 (x, y, z, curvature).
 * out_xyzc: up to cap leaf centroids (x, y, z, mean curvature) in ascending leaf-index order, *m = number of leaves.
 * Points inside a leaf are summed in their original order (PCL's std::sort leaves that order unspecified).  With
 * cfg.scan_resolution > 0, vmp_scan_raw runs this between the motion compensation and the update, and the filter sees
 * the leaf centroids; vmp_get_lidar_cloud returns them (LIOBuilder::lidar_cloud, lio_builder.h:83). */
int vmp_downsample(vmp_handle h, const float* cloud_xyzc, int n, double leaf, float* out_xyzc, int cap, int* m);
int vmp_get_lidar_cloud(vmp_handle h, float* out_xyzc, int cap, int* m);
/* the prior (x, P after the IMU propagation) of the last scan as the device saw it: what vmp_scan / vmp_scan_raw uploaded, or what
 * the device-side propagation of vmp_scan_raw_predict produced */
int vmp_get_prior(vmp_handle h, vmp_state* x, double* P);
int vmp_set_state(vmp_handle h, const vmp_state* x, const double* P);
int vmp_get_state(vmp_handle h, vmp_state* x, double* P);

/* MAP_INIT branch, lio_builder.cpp:185-211: float32 world transform + covariances of
 * the raw cloud with the current state/P, then VoxelMap::build. */
int vmp_first_scan(vmp_handle h, const vmp_state* x, const double* P,
                   const float* pts_lidar, int n, vmp_update_stats* stats);

/* ---- observation / parity hooks ---- */
/* ResidualData records after the last measure call: keys N x 3 (VoxelMap::index of
 * point_world), status bit0 = voxel found, bit1 = is_plane, bit2 = is_valid (after Q2
 * staleness), residual N, plane_norm N x 3. Any pointer may be NULL. */
int vmp_dump_correspondences(vmp_handle h, int64_t* keys, uint8_t* status, double* residual, double* plane_norm, int n);
/* pv_list of the last vmp_scan / vmp_first_scan: points N x 3 (float32 world widened), cov N x 9 */
int vmp_dump_world_points(vmp_handle h, double* pts_world, double* cov, int n);
/* All live voxels in LRU order (front first), as utils.cpp:161-195 walks map->cache. */
int vmp_dump_map(vmp_handle h, vmp_plane* out, int cap, int* count);
/* keys (M x 3) evicted by the last build/update/scan, in eviction order */
int vmp_dump_evicted(vmp_handle h, int64_t* keys, int cap, int* count);
int vmp_map_size(vmp_handle h, int* count);
/* number of kernel launches (graph kernel nodes included) issued by this handle so far */
int64_t vmp_launch_count(vmp_handle h);
/* internal counters of the last map update (8 ints): [0] merge active set after the prefilter,
 * [1] merge events simulated, [2] voxels activated by re-examination */
int vmp_debug_counters(vmp_handle h, int* out8);

/* Per-kernel device timing.  With profiling on, vmp_scan / vmp_scan_dev launch the very same
 * kernels one by one on the handle's stream with a CUDA event after each (instead of the graph)
 * and accumulate the elapsed time per kernel class.  vmp_profile_read returns, per vmp_kernel_id,
 * the accumulated milliseconds and the number of launches since the last reset. */
int vmp_profile_enable(vmp_handle h, int on);
int vmp_profile_reset(vmp_handle h);
int vmp_profile_read(vmp_handle h, double* ms /*VMP_K_COUNT*/, int64_t* launches /*VMP_K_COUNT*/);
const char* vmp_kernel_name(int id);

/* ---- the voxelisation front end of the loop-closure package ----
 * STDManager::buildVoxels (std_matcher/src/std_manager/descriptor.cpp:70-122) on the device: a stateless pass over one cloud
 * (N x 4 float32: x y z intensity, pcl::PointXYZI).  One record per voxel (VoxelKey::index with `voxel_size`, descriptor.cpp:40-45), in the
 * order of the voxels' first points (the reference iterates an unordered_map: its order is unobservable).  sum / ppt are accumulated in
 * point order like the reference's (bit-exact); voxels with more than voxel_min_point points are VALID-sized and get mean, and, if the
 * smallest eigenvalue of their covariance is below voxel_plane_thresh, PLANE with lamdas = (min, mid, max) and norms[3 k + r] =
 * component r of the unit eigenvector k (the reference's VoxelNode::norms.col(k); signs are those of the solver, i.e. undefined).
 * Not bound to a map handle; needs a B200 like everything else. */
#define VMP_STD_F_VALID 1u   /* cloud->size() > voxel_min_point (descriptor.cpp:93) */
#define VMP_STD_F_PLANE 2u   /* VoxelNode::is_plane */
typedef struct vmp_std_voxel {
    int64_t  key[3];
    int32_t  count;
    uint32_t flags;
    double   sum[3];
    double   ppt[9];
    double   mean[3];
    double   lamdas[3];
    double   norms[9];
} vmp_std_voxel;
int vmp_std_build_voxels(const float* cloud_xyzi, int n, double voxel_size, int voxel_min_point, double voxel_plane_thresh,
                         vmp_std_voxel* out, int cap, int* count);
/* device time (ms, CUDA events around the kernels; copies excluded) of this thread's last vmp_std_build_voxels call */
double vmp_std_last_device_ms(void);

/* ---- host-side LIOBuilder (C++ class lio::LIOBuilder in vmp_lio.hpp) through C ---- */
typedef struct vmp_lio_t* vmp_lio;
typedef struct vmp_imu { double acc[3]; double gyro[3]; double timestamp; } vmp_imu;   /* lio::IMUData, commons.h:12-20 */
/* LIOBuilder::loadConfig */
int vmp_lio_create(const vmp_config* cfg, vmp_lio* out);
int vmp_lio_destroy(vmp_lio l);
/* LIOBuilder::process(SyncPackage&) (lio_builder.cpp:175-248). cloud: N x 4 float32
 * (x,y,z,curvature[ms]); sorted and undistorted IN PLACE like the reference. */
int vmp_lio_process(vmp_lio l, const vmp_imu* imus, int n_imu, float* cloud_xyzc, int n,
                    double cloud_start_time, double cloud_end_time, vmp_scan_stats* stats);
/* kf.x(), kf.P(), status (0 IMU_INIT, 1 MAP_INIT, 2 LIO_MAPPING) */
int vmp_lio_state(vmp_lio l, vmp_state* x, double* P, int* status);
/* the device handle behind builder->map */
vmp_handle vmp_lio_map(vmp_lio l);
/* motion compensation on the device (vmp_scan_raw; default) or on the host like the reference */
int vmp_lio_set_device_undistort(vmp_lio l, int on);
/* vmp_set_raw_writeback of the builder's map handle: 0 = process() does not copy the compensated cloud back into `cloud_xyzc` */
int vmp_lio_set_cloud_writeback(vmp_lio l, int on);
/* IMU propagation on the device too (vmp_scan_raw_predict; off by default: on the host it overlaps with the device's map update) */
int vmp_lio_set_device_predict(vmp_lio l, int on);
/* the prior (x, P after IMU propagation) that the last process() handed to the device update */
int vmp_lio_prior(vmp_lio l, vmp_state* x, double* P);

#ifdef __cplusplus
}
#endif
#endif /* VMP_B200_H */
