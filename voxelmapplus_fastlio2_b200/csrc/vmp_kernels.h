// vmp_kernels.h — host-callable launchers of the device kernels (vmp_iekf.cu, vmp_map.cu).
#pragma once
#include <cuda_runtime.h>

#include "../../include/vmp_b200.h"
#include "vmp_device.cuh"

namespace vmp {

constexpr int PARTIAL_STRIDE = 96;      // doubles per block partial (>= 78 + 12 + 1)
constexpr int PT_BLOCK = 1024;          // points per block in the order-preserving passes

// IEKF
// motion compensation of a time-sorted raw scan in place (+ a copy into mapped host memory)
void launch_undistort(cudaStream_t st, int grid, const ScanIn* in, const DevPose* poses, float4* cloud, float4* host_copy);
void launch_cloud_out(cudaStream_t st, int grid, const ScanIn* in, const float4* cloud, float4* host_copy);
// IESKF::predict for every IMU step of a scan + the IMU pose list, from the filter state resident on the device (one CTA)
void launch_predict(cudaStream_t st, DevFilter* f, ScanIn* in, const DevPredictIn* pin, DevPose* poses);
// stages one scan from the mailbox `in` (device-accessible): points -> raw / body points / body covariances, prior, counters
void launch_set_scan(cudaStream_t st, int grid, const DevScan& s, const ScanIn* in, DevFilter* f, DevCtl* ctl);
// solve != 0: one extra CTA runs the 23-dof solve of the iteration (one launch per IEKF iteration).  first != null: the first
// iteration of a scan's graph, which stages the scan named by that header itself (no launch_set_scan in front)
void launch_measure(cudaStream_t st, bool ext, int grid, const DevMap& m, const DevScan& s, DevFilter* f, DevCtl* ctl, double* partials, int solve,
                    const ScanIn* first = nullptr, bool reuse_slots = false);
cudaError_t launch_iekf_loop(cudaStream_t st, bool ext, int grid, const DevMap& m, const DevScan& s, DevFilter* f, DevCtl* ctl, double* partials, const ScanIn* in, bool reuse_slots);
bool iekf_loop_fits(bool ext, int grid, int sm_count);      // reuse_slots: later iterations inside a scan's graph (same map, same points)
// posterior -> host mailbox (x, P, iteration counters, then seq behind a system-scope fence)
void launch_state_out(cudaStream_t st, const DevFilter* f, const DevCtl* ctl, StateOut* out);

// optional per-launch hook (profiling mode records a CUDA event after each kernel)
struct Marker { void (*fn)(void* ctx, int id); void* ctx; };
inline void mark(const Marker* mk, int id) { if (mk && mk->fn) mk->fn(mk->ctx, id); }

// second stream + events for the two side branches of a map update (eviction next to the segment build, LRU-log append
// next to the merge simulation); null = everything on one stream
struct SideStream { cudaStream_t st; cudaEvent_t ev[6]; };

// pcl::VoxelGrid leaf-centroid filter of `cloud` (n_sort points, the same count on the device at *n_ptr) -> d.out / d.m
// (+ mapped host copies); patch != null: the scan header is redirected to the filtered cloud.  Returns the kernel count.
size_t downsample_temp_bytes(int nmax);
int launch_downsample(cudaStream_t st, const DevDown& d, const float4* cloud, const int* n_ptr, int n_sort, float leaf, int grid,
                      float4* host_out, int* host_m, ScanIn* patch, const Marker* mk);

// map read-back in LRU order: live slots -> (stamp, slot); then sort by stamp (descending) + one vmp_plane record per voxel
size_t dump_temp_bytes(int pool);
void launch_dump_collect(cudaStream_t st, int grid, const DevMap& m, const DevDump& d);
void launch_dump_sort_gather(cudaStream_t st, int grid, const DevMap& m, const DevDump& d, int n_live);

// map: returns the number of kernels launched
// `out`: mailbox written by the update's last kernel (counters, error bits, maintenance requests); may be null
// world_mode: 0 = world points / covariances of the scan from the posterior in `f` (lio_builder.cpp:155-163, 231-245), 1 = the same
// for the first scan (calcBodyCov on a local copy), 2 = s.pw / s.pcov were given by the caller
int launch_map_update(cudaStream_t st, const DevMap& m, const DevScan& s, const DevFilter* f, DevCtl* ctl, int sm_count, bool build, int world_mode, MapOut* out,
                      const Marker* mk, const SideStream* side);
int launch_map_maintenance(cudaStream_t st, const DevMap& m, DevCtl* ctl, int sm_count, int what);
void launch_map_init(cudaStream_t st, const DevMap& m, DevCtl* ctl);
cudaError_t map_configure_kernels(const DevMap& m);      // per-handle kernel attributes (dynamic shared memory of k_fill)

}  // namespace vmp
