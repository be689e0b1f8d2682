// vmp_std.cu — SURVEY.md 8(f) row 4: the voxelisation front end of the loop-closure package,
// STDManager::buildVoxels (std_matcher/src/std_manager/descriptor.cpp:70-122; key: VoxelKey::index :40-45, descriptor.h:28-51),
// on the device.  Everything after it (buildConnections, corner NMS, triangle descriptors, KD-tree, GTSAM) stays out of scope.
//
// buildVoxels is a stateless pass over one (sub-map) cloud: every point goes into its voxel (sum p, sum p p^T, point count, in
// point order), then every voxel with more than voxel_min_point points gets mean, covariance, an eigen-decomposition and the
// plane test lambda_min < voxel_plane_thresh.  Here:
//   k_std_insert   key + find-or-create in a scratch hash (same packed key / open addressing as the map), count, first touch
//   (CUB scans)    segment offsets of the voxels; rank of every voxel by its first touching point (= output order, deterministic:
//                  the reference iterates an unordered_map, whose order is unobservable)
//   k_std_seg      point indices grouped by voxel
//   k_std_voxel    one warp per voxel: its points in point order (warp sort), sum p / sum p p^T accumulated in that order on nine
//                  lanes (bit-exact against the reference's order), mean, covariance, 3x3 symmetric eigen-solve, plane test
// The reference calls Eigen::EigenSolver (general real Schur form) on the symmetric covariance; its eigenvalues are those of the
// symmetric solver up to rounding, its eigenvectors unit vectors of arbitrary sign.  The record carries eigenvalues ascending
// (lambda_min, lambda_mid, lambda_max as the reference sorts them) and the matching unit eigenvectors; signs are not defined.
#include <cub/device/device_scan.cuh>

#include <mutex>

#include "vmp_device.cuh"
#include "vmp_kernels.h"

namespace vmp {

static_assert(sizeof(vmp_std_voxel) == 3 * 8 + 4 + 4 + (3 + 9 + 3 + 3 + 9) * 8, "vmp_std_voxel layout");

struct StdScratch {
    unsigned long long* tkey; int* tval; unsigned hmask;
    unsigned long long* vkey; int* vcount; int* vft; int* vcursor; int* voff;     // per voxel
    int* pvox; int* seg; int* isfirst; int* prank;                                // per point
    int* nvox; int* err;
};

__global__ void __launch_bounds__(256) k_std_clear(StdScratch s, int n, size_t hs) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < hs; i += nth) { s.tkey[i] = KEY_EMPTY; s.tval[i] = -1; }
    for (size_t i = tid; i < (size_t)n + 1; i += nth) { s.vcount[i] = 0; s.vft[i] = T_INF; s.vcursor[i] = 0; s.isfirst[i] = 0; }
    if (tid == 0) { *s.nvox = 0; *s.err = 0; }
}

__global__ void __launch_bounds__(256) k_std_insert(StdScratch s, const float4* __restrict__ cloud, int n, double voxel_size) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = cloud[i];
        unsigned long long pk;
        s.pvox[i] = -1;
        // VoxelKey::index (descriptor.cpp:40-45): floor(point / resolution + bias), bias = 0
        if (!voxel_index((double)p.x, (double)p.y, (double)p.z, voxel_size, pk)) { atomicOr(s.err, 1); continue; }
        unsigned h = hash_key(pk) & s.hmask;
        int v = -1;
        for (unsigned probe = 0; probe <= s.hmask; probe++) {
            unsigned long long cur = __ldcg(&s.tkey[h]);
            if (cur == KEY_EMPTY) {
                cur = atomicCAS(&s.tkey[h], KEY_EMPTY, pk);
                if (cur == KEY_EMPTY) {
                    v = atomicAdd(s.nvox, 1);
                    s.vkey[v] = pk;
                    __threadfence();
                    *(volatile int*)&s.tval[h] = v;
                    break;
                }
            }
            if (cur == pk) { while ((v = *(volatile int*)&s.tval[h]) == -1) __nanosleep(20); break; }
            h = (h + 1) & s.hmask;
        }
        if (v < 0) { atomicOr(s.err, 2); continue; }
        atomicAdd(&s.vcount[v], 1);
        atomicMin(&s.vft[v], i);
        s.pvox[i] = v;
    }
}
__global__ void __launch_bounds__(256) k_std_first(StdScratch s, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int v = s.pvox[i];
        s.isfirst[i] = (v >= 0 && s.vft[v] == i) ? 1 : 0;
    }
}
__global__ void __launch_bounds__(256) k_std_seg(StdScratch s, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int v = s.pvox[i];
        if (v >= 0) s.seg[s.voff[v] + atomicAdd(&s.vcursor[v], 1)] = i;
    }
}

// ascending in-place sort of a[0..c) by one warp (distinct values, global memory)
__device__ void std_warp_sort(int* a, int c) {
    const int lane = threadIdx.x & 31;
    if (c <= 1) return;
    int npow = 2;
    while (npow < c) npow <<= 1;
    const int half = npow >> 1;
    for (int k = 2; k <= npow; k <<= 1) {
        const int hk = k >> 1;
        for (int t = lane; t < half; t += 32) {
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

__global__ void __launch_bounds__(128) k_std_voxel(StdScratch s, const float4* __restrict__ cloud, int min_point, double plane_thresh, vmp_std_voxel* out) {
    const int lane = threadIdx.x & 31;
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
    const int V = *s.nvox;
    // lane l < 3: sum[l]; lane 3 + 3 r + c: ppt(r, c)
    const int r = lane < 3 ? lane : (lane - 3) / 3, cc = lane < 3 ? 0 : (lane - 3) % 3;
    for (int v = wg; v < V; v += nW) {
        const int c = s.vcount[v], off = s.voff[v];
        int* order = s.seg + off;
        std_warp_sort(order, c);
        double acc = 0.0;
        for (int j = 0; j < c; j++) {                               // sum += p; ppt += p p^T, in point order (descriptor.cpp:85-88)
            const float4 p = cloud[order[j]];
            const double x[3] = {(double)p.x, (double)p.y, (double)p.z};
            if (lane < 3) acc += x[r];
            else if (lane < 12) acc += x[r] * x[cc];
        }
        vmp_std_voxel* o = out + s.prank[s.vft[v]];                 // output order: first touching point
        if (lane == 0) {
            long long kx, ky, kz;
            unpack_key(s.vkey[v], kx, ky, kz);
            o->key[0] = kx; o->key[1] = ky; o->key[2] = kz;
            o->count = c;
        }
        if (lane < 3) o->sum[lane] = acc;
        else if (lane < 12) o->ppt[lane - 3] = acc;
        double sum3[3], ppt9[9];
#pragma unroll
        for (int k = 0; k < 3; k++) sum3[k] = __shfl_sync(0xffffffffu, acc, k);
#pragma unroll
        for (int k = 0; k < 9; k++) ppt9[k] = __shfl_sync(0xffffffffu, acc, 3 + k);
        if (lane == 0) {
            unsigned flags = 0;
            double mean[3] = {0, 0, 0}, lam[3] = {0, 0, 0}, nrm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (c > min_point) {                                    // descriptor.cpp:93-97 (`size() <= voxel_min_point` is skipped)
                flags |= VMP_STD_F_VALID;
                const double nd = (double)c;
                for (int k = 0; k < 3; k++) mean[k] = sum3[k] / nd;
                // cov = ppt / n - mean mean^T (descriptor.cpp:99); lower triangle into the symmetric solver
                const double c00 = ppt9[0] / nd - mean[0] * mean[0];
                const double c10 = ppt9[3] / nd - mean[1] * mean[0];
                const double c11 = ppt9[4] / nd - mean[1] * mean[1];
                const double c20 = ppt9[6] / nd - mean[2] * mean[0];
                const double c21 = ppt9[7] / nd - mean[2] * mean[1];
                const double c22 = ppt9[8] / nd - mean[2] * mean[2];
                double ev[3];
                M3 evec;
                eig3_sym(c00, c10, c11, c20, c21, c22, ev, evec);    // ascending: min, mid, max (descriptor.cpp:102-106)
                if (ev[0] < plane_thresh) {                         // descriptor.cpp:108-118
                    flags |= VMP_STD_F_PLANE;
                    for (int k = 0; k < 3; k++) { lam[k] = ev[k]; for (int q = 0; q < 3; q++) nrm[3 * k + q] = evec(q, k); }
                }
            }
            o->flags = flags;
            for (int k = 0; k < 3; k++) { o->mean[k] = mean[k]; o->lamdas[k] = lam[k]; }
            for (int k = 0; k < 9; k++) o->norms[k] = nrm[k];
        }
        __syncwarp();
    }
}

}  // namespace vmp

using namespace vmp;

// Scratch comes from a private stream-ordered pool that keeps its memory between calls (release threshold = max): a sub-map pass
// is called once per key frame, and fifteen cudaMalloc / cudaFree pairs cost more than the kernels.
static cudaMemPool_t std_pool() {
    static cudaMemPool_t pool = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { cudaGetLastError(); pool = nullptr; return; }
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    });
    return pool;
}
static thread_local double t_std_device_ms = 0.0;

extern "C" double vmp_std_last_device_ms(void) { return t_std_device_ms; }

extern "C" int vmp_std_build_voxels(const float* cloud_xyzi, int n, double voxel_size, int voxel_min_point, double voxel_plane_thresh,
                                    vmp_std_voxel* out, int cap, int* count) {
    if (n < 0 || (n > 0 && !cloud_xyzi) || !(voxel_size > 0.0) || voxel_min_point < 0 || cap < 0 || (cap > 0 && !out)) { set_error("vmp_std_build_voxels: invalid argument"); return VMP_ERR_INVALID_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); set_error("vmp_std_build_voxels: no CUDA device (there is no CPU fallback)"); return VMP_ERR_NO_DEVICE; }
    if (count) *count = 0;
    t_std_device_ms = 0.0;
    if (n == 0) return VMP_OK;
    cudaMemPool_t pool = std_pool();
    if (!pool) { set_error("vmp_std_build_voxels: cudaMemPoolCreate failed"); return VMP_ERR_CUDA; }
    static cudaStream_t st = nullptr;
    static cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);                 // one pass at a time: the stream and the events are shared
    if (!st) {
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&ev0) != cudaSuccess || cudaEventCreate(&ev1) != cudaSuccess) {
            set_error("vmp_std_build_voxels: stream / event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
            st = nullptr;
            return VMP_ERR_CUDA;
        }
    }
    size_t hs = 1024;
    while (hs < (size_t)n * 2) hs <<= 1;
    const size_t np1 = (size_t)n + 1;
    size_t temp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, (int*)nullptr, (int*)nullptr, n + 1);
    // one arena: [tkey hs*8][vkey np1*8][cloud n*16][ints: tval hs, 8 x np1, 2][cub temp]
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_tkey = 0, o_vkey = o_tkey + up(hs * 8), o_cloud = o_vkey + up(np1 * 8), o_tval = o_cloud + up((size_t)n * 16);
    const size_t o_int = o_tval + up(hs * 4), o_ctr = o_int + 8 * up(np1 * 4), o_temp = o_ctr + 256, total = o_temp + up(temp_bytes + 16);
    char* base = nullptr;
    vmp_std_voxel* d_out = nullptr;
    auto cleanup = [&]() { if (base) cudaFreeAsync(base, st); if (d_out) cudaFreeAsync(d_out, st); cudaStreamSynchronize(st); };
#define STD_CHECK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error("%s failed: %s", #expr, cudaGetErrorString(_e)); cleanup(); return VMP_ERR_CUDA; } } while (0)
    STD_CHECK(cudaMallocFromPoolAsync((void**)&base, total, pool, st));
    StdScratch s{};
    s.hmask = (unsigned)(hs - 1);
    s.tkey = (unsigned long long*)(base + o_tkey); s.vkey = (unsigned long long*)(base + o_vkey); s.tval = (int*)(base + o_tval);
    int** ints[8] = {&s.vcount, &s.vft, &s.vcursor, &s.voff, &s.pvox, &s.seg, &s.isfirst, &s.prank};
    for (int k = 0; k < 8; k++) *ints[k] = (int*)(base + o_int + k * up(np1 * 4));
    s.nvox = (int*)(base + o_ctr); s.err = s.nvox + 1;
    float4* d_cloud = (float4*)(base + o_cloud);
    void* d_temp = base + o_temp;
    STD_CHECK(cudaMemcpyAsync(d_cloud, cloud_xyzi, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, st));
    const int grid = 148 * 4;
    STD_CHECK(cudaEventRecord(ev0, st));
    k_std_clear<<<grid, 256, 0, st>>>(s, n, hs);
    k_std_insert<<<grid, 256, 0, st>>>(s, d_cloud, n, voxel_size);
    k_std_first<<<grid, 256, 0, st>>>(s, n);
    STD_CHECK(cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, s.vcount, s.voff, n + 1, st));
    STD_CHECK(cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, s.isfirst, s.prank, n + 1, st));
    k_std_seg<<<grid, 256, 0, st>>>(s, n);
    int hv[2] = {0, 0};
    STD_CHECK(cudaMemcpyAsync(hv, s.nvox, 8, cudaMemcpyDeviceToHost, st));      // the record array is sized by the voxel count
    STD_CHECK(cudaStreamSynchronize(st));
    if (hv[1] & 1) { set_error("vmp_std_build_voxels: a point is NaN or outside +-2^20 voxels"); cleanup(); return VMP_ERR_INVALID_ARG; }
    if (hv[1] & 2) { set_error("vmp_std_build_voxels: scratch hash full"); cleanup(); return VMP_ERR_CAPACITY; }
    STD_CHECK(cudaMallocFromPoolAsync((void**)&d_out, sizeof(vmp_std_voxel) * (size_t)hv[0], pool, st));
    k_std_voxel<<<grid, 128, 0, st>>>(s, d_cloud, voxel_min_point, voxel_plane_thresh, d_out);
    STD_CHECK(cudaEventRecord(ev1, st));
    if (count) *count = hv[0];
    const int m = hv[0] < cap ? hv[0] : cap;
    if (m > 0) STD_CHECK(cudaMemcpyAsync(out, d_out, sizeof(vmp_std_voxel) * (size_t)m, cudaMemcpyDeviceToHost, st));
    STD_CHECK(cudaStreamSynchronize(st));
    STD_CHECK(cudaGetLastError());
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess) t_std_device_ms = (double)ms;
    cleanup();
    return VMP_OK;
#undef STD_CHECK
}
