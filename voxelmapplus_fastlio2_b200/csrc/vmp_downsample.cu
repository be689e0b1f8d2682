// vmp_downsample.cu — pcl::VoxelGrid<PointXYZINormal>::filter on the device (second half).
//
// Reference call site: scan_filter.setLeafSize(r, r, r) / scan_filter.filter(*lidar_cloud), lio_builder.cpp:13-14,
// 215-219; the algorithm is PCL's filters/impl/voxel_grid.hpp applyFilter (PCL is not vendored by the reference and not
// in this image: restated from its published algorithm, like oracle/oracle.cpp LIOBuilder::voxelGridFilter):
//
//   k_ds_minmax     bounding box of the finite points (getMinMax3D), float32, order-encoded atomics
//   k_ds_keys       leaf index ijk = floor(p * inverse_leaf) - min_b per axis, idx = i + j dx + k dx dy (uint32);
//                   "leaf size too small" (dx dy dz > INT32_MAX) -> idx = point index, i.e. output = input as PCL does
//   radix sort      (idx, point index) pairs by idx; STABLE, i.e. the original order inside a leaf — a valid outcome of
//                   PCL's std::sort and the order the oracle fixes.  CUB's DeviceRadixSort: a plain library sort
//   k_ds_heads      leaf boundaries of the sorted list; inclusive scan (CUB) -> output position of every leaf
//   k_ds_centroid   one thread per leaf: float32 sums of x, y, z, curvature in that order, divided by n
//                   (CentroidPoint / AccumulatorXYZ / AccumulatorCurvature), output in ascending idx order; the last
//                   block also patches the scan header so that the update that follows reads the filtered cloud
//
// Everything float32 with separate IEEE multiply / add / divide (no contraction): bit-exact against the oracle.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "vmp_device.cuh"
#include "vmp_kernels.h"

namespace vmp {

constexpr unsigned DS_INVALID = 0xFFFFFFFFu;

__device__ __forceinline__ unsigned f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned e) {
    return __uint_as_float((e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e);
}
__device__ __forceinline__ bool finite3(const float4& p) { return isfinite(p.x) && isfinite(p.y) && isfinite(p.z); }

// mm[0..2] = max over points of ~ord(min candidate) (so that one zero-initialised buffer serves both), mm[3..5] = ord(max),
// mm[6] = number of finite points
__global__ void __launch_bounds__(256) k_ds_minmax(const float4* __restrict__ cloud, const int* __restrict__ n_ptr, unsigned* mm) {
    const int n = *n_ptr;
    unsigned lo[3] = {0u, 0u, 0u}, hi[3] = {0u, 0u, 0u};
    unsigned cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = cloud[i];
        if (!finite3(p)) continue;
        const unsigned e[3] = {f2ord(p.x), f2ord(p.y), f2ord(p.z)};
#pragma unroll
        for (int a = 0; a < 3; a++) { lo[a] = max(lo[a], ~e[a]); hi[a] = max(hi[a], e[a]); }
        cnt++;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            lo[a] = max(lo[a], __shfl_down_sync(0xffffffffu, lo[a], o));
            hi[a] = max(hi[a], __shfl_down_sync(0xffffffffu, hi[a], o));
        }
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0 && cnt) {
#pragma unroll
        for (int a = 0; a < 3; a++) { atomicMax(&mm[a], lo[a]); atomicMax(&mm[3 + a], hi[a]); }
        atomicAdd(&mm[6], cnt);
    }
}

__global__ void __launch_bounds__(256) k_ds_keys(const float4* __restrict__ cloud, const int* __restrict__ n_ptr, const unsigned* __restrict__ mm,
                                                 float inv, unsigned* keys, int* vals, int n_sort) {
    const int n = *n_ptr;
    __shared__ int s_minb[3], s_mul[2], s_overflow;
    if (threadIdx.x == 0) {
        int ov = 0, minb[3] = {0, 0, 0}, div[3] = {1, 1, 1};
        if (mm[6] > 0) {
            long long d[3];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const float mn = ord2f(~mm[a]), mx = ord2f(mm[3 + a]);
                d[a] = (long long)__fmul_rn(__fsub_rn(mx, mn), inv) + 1;
                minb[a] = (int)floorf(__fmul_rn(mn, inv));
                div[a] = (int)floorf(__fmul_rn(mx, inv)) - minb[a] + 1;
            }
            ov = d[0] * d[1] * d[2] > (long long)INT_MAX;
        }
        s_minb[0] = minb[0]; s_minb[1] = minb[1]; s_minb[2] = minb[2];
        s_mul[0] = div[0]; s_mul[1] = div[0] * div[1];
        s_overflow = ov;
    }
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_sort; i += gridDim.x * blockDim.x) {
        unsigned key = DS_INVALID;
        if (i < n) {
            const float4 p = cloud[i];
            if (s_overflow) key = (unsigned)i;                 // PCL: warning + output = input
            else if (finite3(p)) {
                const int i0 = (int)__fsub_rn(floorf(__fmul_rn(p.x, inv)), (float)s_minb[0]);
                const int i1 = (int)__fsub_rn(floorf(__fmul_rn(p.y, inv)), (float)s_minb[1]);
                const int i2 = (int)__fsub_rn(floorf(__fmul_rn(p.z, inv)), (float)s_minb[2]);
                key = (unsigned)(i0 + i1 * s_mul[0] + i2 * s_mul[1]);
            }
        }
        keys[i] = key;
        vals[i] = i;
    }
}

__global__ void __launch_bounds__(256) k_ds_heads(const unsigned* __restrict__ skeys, int n_sort, int* head) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_sort; j += gridDim.x * blockDim.x) {
        const unsigned k = skeys[j];
        head[j] = (k != DS_INVALID && (j == 0 || skeys[j - 1] != k)) ? 1 : 0;
    }
}

// rank[j] = inclusive sum of head: a leaf that starts at sorted position j is output point rank[j] - 1
__global__ void __launch_bounds__(128) k_ds_centroid(const float4* __restrict__ cloud, const unsigned* __restrict__ skeys, const int* __restrict__ svals,
                                                    const int* __restrict__ head, const int* __restrict__ rank, int n_sort,
                                                    float4* out, float4* host_out, int* m_out, int* host_m, ScanIn* patch, unsigned* mm) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_sort; j += gridDim.x * blockDim.x) {
        if (!head[j]) continue;
        const unsigned k = skeys[j];
        float sx = 0.0f, sy = 0.0f, sz = 0.0f, sc = 0.0f;
        // The float sums are order-dependent (pcl::VoxelGrid adds the points of a leaf in their original order), so a leaf stays one
        // sequential chain of adds - but not of memory latencies: eight points are fetched per round trip (keys, indices, then the
        // scattered points), then added in order.  A dense near-field leaf holds thousands of points.
        int l = j;
        bool more = true;
        while (more) {
            unsigned kk[8];
            int vv[8];
#pragma unroll
            for (int u = 0; u < 8; u++) { kk[u] = (l + u < n_sort) ? skeys[l + u] : DS_INVALID; vv[u] = (l + u < n_sort) ? svals[l + u] : 0; }
            float4 pp[8];
#pragma unroll
            for (int u = 0; u < 8; u++) pp[u] = (l + u < n_sort && kk[u] == k) ? cloud[vv[u]] : make_float4(0.f, 0.f, 0.f, 0.f);
            int took = 0;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (more && l + u < n_sort && kk[u] == k) {
                    sx = __fadd_rn(sx, pp[u].x); sy = __fadd_rn(sy, pp[u].y); sz = __fadd_rn(sz, pp[u].z); sc = __fadd_rn(sc, pp[u].w);
                    took++;
                } else more = false;
            }
            l += took;
        }
        const float cnt = (float)(l - j);
        const float4 c = make_float4(__fdiv_rn(sx, cnt), __fdiv_rn(sy, cnt), __fdiv_rn(sz, cnt), __fdiv_rn(sc, cnt));
        const int r = rank[j] - 1;
        out[r] = c;
        if (host_out) host_out[r] = c;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int m = n_sort > 0 ? rank[n_sort - 1] : 0;
        *m_out = m;
        if (host_m) *host_m = m;
        if (patch) { patch->n = m; patch->pts = reinterpret_cast<const float*>(out); patch->stride = 4; }
        for (int q = 0; q < 8; q++) mm[q] = 0u;               // ready for the next scan (every reader of mm has finished: previous kernels)
    }
}

size_t downsample_temp_bytes(int nmax) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const unsigned*)nullptr, (unsigned*)nullptr, (const int*)nullptr, (int*)nullptr, nmax, 0, 32);
    cub::DeviceScan::InclusiveSum(nullptr, b, (const int*)nullptr, (int*)nullptr, nmax);
    return (a > b ? a : b) + 256;
}

// n_sort = the caller's point count (known on the host); n_ptr = the same on the device (header of the uploaded scan)
int launch_downsample(cudaStream_t st, const DevDown& d, const float4* cloud, const int* n_ptr, int n_sort, float leaf, int grid,
                      float4* host_out, int* host_m, ScanIn* patch, const Marker* mk) {
    const float inv = 1.0f / leaf;
    int k = 0;
    k_ds_minmax<<<grid, 256, 0, st>>>(cloud, n_ptr, d.mm); k++;
    if (n_sort > 0) {
        k_ds_keys<<<grid, 256, 0, st>>>(cloud, n_ptr, d.mm, inv, d.keys[0], d.vals[0], n_sort); k++;
        size_t tb = d.temp_bytes;
        cub::DeviceRadixSort::SortPairs(d.temp, tb, d.keys[0], d.keys[1], d.vals[0], d.vals[1], n_sort, 0, 32, st); k += 4;
        k_ds_heads<<<grid, 256, 0, st>>>(d.keys[1], n_sort, d.head); k++;
        tb = d.temp_bytes;
        cub::DeviceScan::InclusiveSum(d.temp, tb, d.head, d.rank, n_sort, st); k += 2;
    }
    k_ds_centroid<<<grid, 128, 0, st>>>(cloud, d.keys[1], d.vals[1], d.head, d.rank, n_sort, d.out, host_out, d.m, host_m, patch, d.mm); k++;
    mark(mk, VMP_K_DOWNSAMPLE);
    return k;
}

}  // namespace vmp
