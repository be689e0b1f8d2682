// vmp_readback.cu — map read-back on the device.
//
// The reference's consumers walk VoxelMap::cache front to back and read VoxelGrid / Plane fields directly
// (voxel2MarkerArray, utils.cpp:154-208: is_plane, update_enable, center, plane->cov trace, norm, merged; lio_node.cpp:185-190).
// Here the voxels live in SoA slot arrays with an LRU stamp per slot, so the walk is:
//   k_dump_collect   every live slot -> (stamp, slot) pair (order irrelevant)
//   radix sort       descending stamp = front of `cache` first (stamps are unique: one per inserting point); CUB
//   k_dump_gather    one warp per voxel assembles its vmp_plane record (61 words) with coalesced stores
// and ONE device-to-host copy of live x 488 bytes instead of the whole slot pool.
#include <cub/device/device_radix_sort.cuh>

#include "vmp_device.cuh"
#include "vmp_kernels.h"

namespace vmp {

static_assert(sizeof(vmp_plane) == 61 * 8, "vmp_plane is 61 eight-byte words");

__global__ void __launch_bounds__(256) k_dump_collect(DevMap m, unsigned long long* keys, int* vals, int* count) {
    const int lane = threadIdx.x & 31;
    for (int s0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); s0 < m.pool; s0 += gridDim.x * blockDim.x) {    // warp-uniform trip count
        const int s = s0 + lane;
        const unsigned long long st = s < m.pool ? m.stamp[s] : 0ull;
        const bool live = st != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, live);
        if (!bal) continue;
        int base = 0;
        if (lane == 0) base = atomicAdd(count, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (live) {
            const int k = base + __popc(bal & ((1u << lane) - 1));
            keys[k] = st;
            vals[k] = s;
        }
    }
}

// word w of the record of slot s at LRU rank r (layout of vmp_plane, include/vmp_b200.h)
__device__ __forceinline__ unsigned long long plane_word(const DevMap& m, int s, int r, int w) {
    if (w < 3) {
        long long x, y, z;
        unpack_key(m.skey[s], x, y, z);
        return (unsigned long long)(w == 0 ? x : w == 1 ? y : z);
    }
    if (w < 6) return (unsigned long long)__double_as_longlong(m.hot[(size_t)s * 8 + (w - 3)]);                // mean
    if (w < 15) {                                                                                                // ppt 3x3 from its 6 unique entries
        const int e = w - 6, i = e / 3, j = e % 3, a = i < j ? i : j, b = i < j ? j : i;
        const int idx = a == 0 ? (b == 0 ? 0 : b == 1 ? 1 : 3) : a == 1 ? (b == 1 ? 2 : 4) : 5;
        return (unsigned long long)__double_as_longlong(m.ppt[(size_t)s * 6 + idx]);
    }
    if (w < 18) return (unsigned long long)__double_as_longlong(m.hot[(size_t)s * 8 + 3 + (w - 15)]);           // norm
    if (w < 54) return (unsigned long long)__double_as_longlong(m.cov[(size_t)s * 36 + (w - 18)]);
    if (w < 57) return (unsigned long long)__double_as_longlong(m.center[(size_t)s * 3 + (w - 54)]);
    if (w == 57 || w == 58) {
        const long long fw = __double_as_longlong(m.hot[(size_t)s * 8 + 6]);
        const unsigned flags = (unsigned)(fw & 0xFFFFFFFFll);
        const unsigned n = (unsigned)(fw >> 32);
        if (w == 57) return (unsigned long long)n | ((unsigned long long)(unsigned)m.n_temp[s] << 32);          // n, n_temp
        return (unsigned long long)(unsigned)m.newly[s] | ((unsigned long long)flags << 32);                     // newly_add_point, flags
    }
    if (w == 59) return m.sgroup[s];
    return (unsigned long long)r;                                                                                // lru_rank
}

__global__ void __launch_bounds__(256) k_dump_gather(DevMap m, const int* __restrict__ sorted_slots, int n_live, unsigned long long* out) {
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_live; r += nwarps) {
        const int s = sorted_slots[r];
        unsigned long long* o = out + (size_t)r * 61;
        o[lane] = plane_word(m, s, r, lane);
        if (lane + 32 < 61) o[lane + 32] = plane_word(m, s, r, lane + 32);
    }
}

size_t dump_temp_bytes(int pool) {
    size_t a = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, a, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                              (const int*)nullptr, (int*)nullptr, pool);
    return a + 256;
}

void launch_dump_collect(cudaStream_t st, int grid, const DevMap& m, const DevDump& d) {
    cudaMemsetAsync(d.count, 0, sizeof(int), st);
    k_dump_collect<<<grid, 256, 0, st>>>(m, d.keys[0], d.vals[0], d.count);
}
void launch_dump_sort_gather(cudaStream_t st, int grid, const DevMap& m, const DevDump& d, int n_live) {
    if (n_live <= 0) return;
    size_t tb = d.temp_bytes;
    cub::DeviceRadixSort::SortPairsDescending(d.temp, tb, d.keys[0], d.keys[1], d.vals[0], d.vals[1], n_live, 0, 64, st);
    k_dump_gather<<<grid, 256, 0, st>>>(m, d.vals[1], n_live, d.records);
}

}  // namespace vmp
