// vmp_capi.cu — the C ABI of include/vmp_b200.h over the sm_100a kernels: handle
// life-cycle, HBM allocation, per-scan CUDA graph, host<->device staging.
//
// One handle owns: the voxel map (DevMap), the persistent residual buffer (DevScan), the
// filter (DevFilter), a control block (DevCtl), two streams and two instantiated CUDA graphs
// (with / without the motion compensation in front) that contain every kernel of a scan:
// set_scan, max_iter x measure (each with its solver CTA), state_out and world_points and the
// map-update kernels, three of them on side branches.  Later IEKF iterations become
// no-ops through the device-side `done` flag (ieskf.cpp:148 early break), so the graph
// topology never changes and a scan is: ONE H2D copy (header + prior + points from pinned
// staging), ONE graph launch; the results are written to mapped host mailboxes by the kernels.
//
// There is no CPU implementation behind this file: without a B200 every call fails.
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <vector>

#include "../../include/vmp_b200.h"
#include "vmp_device.cuh"
#include "vmp_kernels.h"
#include "vmp_state.cuh"
#include "vmp_stage.hpp"

namespace vmp {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// H (12x12) / b (12) / effect of one measurement pass from the block partials (vmp_measure)
__global__ void k_reduce_partials(const double* partials, int nblocks, int ext, double* out /*144+12+1*/) {
    const int D = ext ? 12 : 6;
    const int NH = D * (D + 1) / 2, NV = NH + D + 1;
    __shared__ double v[96];
    const int tid = threadIdx.x;
    if (tid < NV) {
        double t = partials[tid];
        for (int b = 1; b < nblocks; b++) t += partials[(size_t)b * PARTIAL_STRIDE + tid];
        v[tid] = t;
    }
    __syncthreads();
    if (tid < 144) {
        const int i = tid / 12, j = tid % 12;
        double h = 0.0;
        if (i < D && j < D) { const int a = i < j ? i : j, c = i < j ? j : i; h = v[a * D - a * (a - 1) / 2 + (c - a)]; }
        out[tid] = h;
    } else if (tid < 156) {
        const int a = tid - 144;
        out[tid] = a < D ? v[NH + a] : 0.0;
    } else if (tid == 156) out[156] = v[NH + D];
}
__global__ void k_reset_iter(DevCtl* ctl) { ctl->iter = 0; ctl->done = 0; ctl->converged = 0; }

}  // namespace vmp

using namespace vmp;

constexpr size_t IN_HDR = 4608;          // ScanIn rounded up
constexpr size_t POSE_BYTES = sizeof(DevPose) * MAX_POSES;
constexpr size_t PTS_OFF = IN_HDR + POSE_BYTES;      // staging layout: [ScanIn | IMU poses | points (xyz or xyz+time)]
static_assert(sizeof(ScanIn) <= IN_HDR, "ScanIn outgrew its slot");

struct vmp_handle_t {
    vmp_config cfg;
    int device = 0, sm_count = 148;
    cudaStream_t stream = nullptr;
    SideStream side{};                   // second stream of the map update's side branches
    cudaEvent_t ev_so[3] = {nullptr, nullptr, nullptr};   // fork / join of the posterior write-out; [2]: fork of the compensated cloud's copy to the host
    cudaGraphExec_t graph = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    DevMap m{};
    DevScan s{};
    DevFilter* f = nullptr;
    DevCtl* ctl = nullptr;
    double* partials = nullptr;
    double* meas_out = nullptr;          // 157 doubles
    // mailboxes in pinned, mapped host memory (h_*) and their device aliases (a_*): the kernels read the scan and write
    // the results directly, one scan is one graph launch
    // scan upload: [ScanIn header | points] contiguous in pinned memory, ONE DMA copy into the same layout on the device
    unsigned char* h_stage = nullptr; unsigned char* d_stage = nullptr;
    ScanIn* h_in = nullptr;     ScanIn* d_in = nullptr;
    StateOut* h_sout = nullptr; StateOut* a_sout = nullptr;
    MapOut* h_mout = nullptr;   MapOut* a_mout = nullptr;      // TWO mailboxes, indexed by scan parity (seq & 1)
    unsigned long long map_seq = 0;      // sequence number the last launched map update echoes in its mailbox
    vmp::StagePool pool;                 // helper threads of the pageable -> pinned staging copy (vmp_stage.hpp)
    bool raw_writeback = true;           // vmp_scan_raw copies the compensated cloud back into the caller's buffer (reference: package.cloud is edited in place)
    float* h_raw = nullptr;              // = h_stage + PTS_OFF
    float4* h_cloud = nullptr; float4* a_cloud = nullptr;    // undistorted cloud written by k_undistort (pinned, mapped)
    DevDown ds{};                        // pcl::VoxelGrid downsample scratch (vmp_downsample.cu)
    float4* h_ds = nullptr; float4* a_ds = nullptr;          // filtered cloud = LIOBuilder::lidar_cloud (pinned, mapped)
    int* h_ds_m = nullptr; int* a_ds_m = nullptr;
    DevDump dump{};                      // map read-back scratch, allocated by the first vmp_dump_map
    bool dump_ready = false;
    bool ds_valid = false;               // the last scan went through the downsample
    bool last_raw = false;               // the last scan was a vmp_scan_raw (its compensated cloud is in h_cloud)
    cudaGraphExec_t graph_raw = nullptr; // the scan graph with the motion compensation in front
    int graph_raw_kernels = 0;
    cudaGraphExec_t graph_pred = nullptr;    // ... and with the IMU propagation (k_predict) in front of that
    int graph_pred_kernels = 0;
    DevPredictIn* h_pred = nullptr; DevPredictIn* d_pred = nullptr;      // IMU steps of the scan (pinned / device)
    unsigned long long seq = 0;
    // streamed upload of vmp_scan (pageable pointer): chunk 0 and the graph go to `stream`, the other chunks to `up_st`, each followed by a
    // stream memory operation (cuStreamWriteValue64, taken from the driver at run time) that tells the running first measurement pass it is there
    cudaStream_t up_st = nullptr;
    void* write_value64 = nullptr;
    bool upload_gate = false;
    int gate_chunks = 8;                 // chunks of a streamed upload (VMP_UPLOAD_CHUNKS)
    bool upload_trace = false;           // VMP_UPLOAD_TRACE=1: host-side timeline of every 16th streamed upload on stderr
    size_t chunk_min = 1u << 20;         // point bytes from which a pageable scan is staged / uploaded in 4 chunks (8 from four times that)
    bool pipelined = false;              // vmp_set_pipelined: vmp_scan returns when the posterior is out, the map update runs on
    bool map_pending = false;            // a map update whose MapOut has not been consumed yet
    cudaEvent_t pe0[2] = {nullptr, nullptr}, pe1[2] = {nullptr, nullptr};
    vmp_update_stats lag_map{};          // pipelined mode: counters of the previous scan's map update
    double* d_up = nullptr;              // staging for vmp_map_update uploads (reuses s.pw / s.pcov)
    std::vector<void*> allocs;
    int grid_pts = 148, grid_meas = 148;
    bool iekf_loop = false;     // the IEKF iterations of a scan as one resident launch (needs grid_meas + 1 co-resident CTAs)
    int n_last = 0;
    bool map_built = false;
    int64_t launches = 0;
    int graph_kernels = 0;
    // profiling mode: direct launches with an event after each kernel
    bool prof_on = false;
    std::vector<cudaEvent_t> pev;        // pev[0] = start, pev[k] after the k-th kernel
    std::vector<int> pev_id;
    int pev_n = 0;
    double prof_ms[VMP_K_COUNT] = {};
    int64_t prof_cnt[VMP_K_COUNT] = {};
};

namespace {

template <typename T>
int dalloc(vmp_handle_t* h, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T));
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e)); return VMP_ERR_CUDA; }
    h->allocs.push_back(q);
    *p = (T*)q;
    return VMP_OK;
}
#define DALLOC(ptr, count) do { int _r = dalloc(h, &(ptr), (size_t)(count)); if (_r) return _r; } while (0)

int check_device_err(vmp_handle_t* h, int err) {
    if (!err) return VMP_OK;
    set_error("device reported error bits 0x%x:%s%s%s%s%s%s%s%s%s%s%s", err,
              (err & E_KEY_RANGE) ? " (key range);" : "",
              (err & E_POOL) ? " voxel slot pool exhausted;" : "",
              (err & E_LRU_EXHAUSTED) ? " map_capacity smaller than the voxels one scan touches (LRU victim was touched in the same scan);" : "",
              (err & E_REFIT_OVERFLOW) ? " refit of a voxel holding more than max_point_thresh points;" : "",
              (err & E_QUEUE) ? " internal queue overflow;" : "",
              (err & E_HASH_FULL) ? " hash table full;" : "",
              (err & E_MERGE_DEPTH) ? " merge phase: serial redo needed but the undo log had overflowed;" : "",
              (err & E_MERGE_CAP) ? " (merge cap);" : "",
              (err & E_LOG_CAP) ? " LRU log full;" : "",
              (err & E_FILL_CAP) ? " refit job / contribution staging exhausted;" : "",
              (err & E_UPLOAD) ? " a chunk of the streamed scan upload never arrived;" : "");
    (void)h;
    return VMP_ERR_CAPACITY;
}

void fill_update_stats(const DevStats& d, vmp_update_stats* st) {
    if (!st) return;
    st->n_points = d.n_points; st->n_ins = d.n_ins; st->n_touch = d.n_touch; st->n_created = d.n_created;
    st->n_refit = d.n_refit; st->refit_points = d.refit_points; st->n_full = d.n_full;
    st->n_mergeprobe = d.n_mergeprobe; st->n_merge = d.n_merge; st->n_evicted = d.n_evicted; st->map_size = d.map_size; st->n_mergevox = d.n_mergevox; st->n_skipped = d.n_skipped;
}

void prof_mark(void* ctx, int id) {
    vmp_handle_t* h = (vmp_handle_t*)ctx;
    if (h->pev_n + 1 >= (int)h->pev.size()) return;
    h->pev_n++;
    cudaEventRecord(h->pev[h->pev_n], h->stream);
    h->pev_id[h->pev_n] = id;
}

// enqueue every kernel of one scan on h->stream (used both for graph capture and, in
// profiling mode, directly with an event after each launch)
int enqueue_scan(vmp_handle_t* h, const Marker* mk, bool raw, bool predict = false) {
    cudaStream_t st = h->stream;
    const bool ext = h->cfg.estimate_ext != 0;
    int k = 0;
    if (predict) {  // ieskf.cpp:101-123 for every IMU step of the scan, from the state resident on the device
        launch_predict(st, h->f, h->d_in, h->d_pred, (DevPose*)(h->d_stage + IN_HDR)); k++; mark(mk, VMP_K_UPDATE_BEGIN);
    }
    if (raw) {      // lio_builder.cpp:127-152 in front of the timed region: the points are compensated where they were uploaded
        // the compensated cloud goes back to the host (reference: package.cloud is edited in place) from a side branch that runs beside
        // the IEKF: written from k_undistort itself, the stores to mapped host memory were half of that kernel (10 of 20 us at 20 000 points)
        const bool fork_cloud = mk == nullptr;
        launch_undistort(st, h->grid_pts, h->d_in, (const DevPose*)(h->d_stage + IN_HDR), (float4*)(h->d_stage + PTS_OFF), nullptr); k++; mark(mk, VMP_K_UNDISTORT);
        if (fork_cloud) {
            cudaEventRecord(h->ev_so[2], st); cudaStreamWaitEvent(h->side.st, h->ev_so[2], 0);
            launch_cloud_out(h->side.st, h->grid_pts, h->d_in, (const float4*)(h->d_stage + PTS_OFF), h->a_cloud); k++;
        } else {
            launch_cloud_out(st, h->grid_pts, h->d_in, (const float4*)(h->d_stage + PTS_OFF), h->a_cloud); k++; mark(mk, VMP_K_UNDISTORT);
        }
    }
    // the first iteration stages the scan itself (calcBodyCov per point, prior from the header): no staging kernel in front
    if (h->iekf_loop) {     // all iterations in one resident launch (k_iekf_loop)
        launch_iekf_loop(st, ext, h->grid_meas, h->m, h->s, h->f, h->ctl, h->partials, h->d_in, true); k++; mark(mk, VMP_K_MEASURE);
    } else {
        for (int it = 0; it < h->cfg.opti_max_iter; it++) {
            launch_measure(st, ext, h->grid_meas, h->m, h->s, h->f, h->ctl, h->partials, 1, it == 0 ? h->d_in : nullptr, it > 0); k++; mark(mk, VMP_K_MEASURE);
        }
    }
    // posterior -> host mailbox on the side stream (joined after the map update, which must not overwrite anything it
    // reads: f->x / f->P / the iteration counters are only written by the next scan)
    const bool fork = mk == nullptr;
    if (fork) {
        cudaEventRecord(h->ev_so[0], st); cudaStreamWaitEvent(h->side.st, h->ev_so[0], 0);
        launch_state_out(h->side.st, h->f, h->ctl, h->a_sout); k++;
        cudaEventRecord(h->ev_so[1], h->side.st);
    } else {
        launch_state_out(st, h->f, h->ctl, h->a_sout); k++; mark(mk, VMP_K_SCAN_OUT);
    }
    k += launch_map_update(st, h->m, h->s, h->f, h->ctl, h->sm_count, false, 0, h->a_mout, mk, &h->side);
    if (fork) cudaStreamWaitEvent(st, h->ev_so[1], 0);
    return k;
}

// run one scan: the instantiated graph, or (profiling) the same kernels one by one
int run_scan(vmp_handle_t* h, bool raw, int n_raw = 0, bool predict = false) {
    const bool ds = raw && h->cfg.scan_resolution > 0.0;
    h->ds_valid = ds;
    h->last_raw = raw;
    if (ds) {
        // lio_builder.cpp:127-152 + 215-219 in front of the update: the sort inside the filter is sized by the caller's point
        // count, so these launches are not part of the instantiated graph; the filter's last kernel redirects the scan header
        // (n, points, stride) to the leaf centroids and the plain scan graph follows
        Marker mk{prof_mark, h};
        if (h->prof_on) { h->pev_n = 0; VMP_CUDA_CHECK(cudaEventRecord(h->pev[0], h->stream)); }
        if (predict) { launch_predict(h->stream, h->f, h->d_in, h->d_pred, (DevPose*)(h->d_stage + IN_HDR)); h->launches++; if (h->prof_on) mark(&mk, VMP_K_UPDATE_BEGIN); }
        launch_undistort(h->stream, h->grid_pts, h->d_in, (const DevPose*)(h->d_stage + IN_HDR), (float4*)(h->d_stage + PTS_OFF), h->a_cloud);
        if (h->prof_on) mark(&mk, VMP_K_UNDISTORT);
        h->launches += 1 + launch_downsample(h->stream, h->ds, (const float4*)(h->d_stage + PTS_OFF), &h->d_in->n, n_raw, (float)h->cfg.scan_resolution,
                                             h->grid_pts, h->a_ds, h->a_ds_m, h->d_in, h->prof_on ? &mk : nullptr);
        raw = false; predict = false;
        if (h->prof_on) {
            h->launches += enqueue_scan(h, &mk, false);
            VMP_CUDA_CHECK(cudaGetLastError());
            return VMP_OK;
        }
    }
    if (!h->prof_on) {
        VMP_CUDA_CHECK(cudaGraphLaunch(predict ? h->graph_pred : raw ? h->graph_raw : h->graph, h->stream));
        h->launches += predict ? h->graph_pred_kernels : raw ? h->graph_raw_kernels : h->graph_kernels;
        return VMP_OK;
    }
    Marker mk{prof_mark, h};
    h->pev_n = 0;
    VMP_CUDA_CHECK(cudaEventRecord(h->pev[0], h->stream));
    h->launches += enqueue_scan(h, &mk, raw, predict);
    VMP_CUDA_CHECK(cudaGetLastError());
    return VMP_OK;
}

void prof_collect(vmp_handle_t* h) {      // after the stream has been synchronised
    if (!h->prof_on) return;
    for (int k = 1; k <= h->pev_n; k++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->pev[k - 1], h->pev[k]) == cudaSuccess) {
            h->prof_ms[h->pev_id[k]] += ms;
            h->prof_cnt[h->pev_id[k]] += 1;
        }
    }
    h->pev_n = 0;
}

int build_graph(vmp_handle_t* h) {
    cudaGraph_t g = nullptr;
    VMP_CUDA_CHECK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    h->graph_kernels = enqueue_scan(h, nullptr, false);
    VMP_CUDA_CHECK(cudaStreamEndCapture(h->stream, &g));
    VMP_CUDA_CHECK(cudaGraphInstantiate(&h->graph, g, 0));
    VMP_CUDA_CHECK(cudaGraphDestroy(g));
    VMP_CUDA_CHECK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    h->graph_raw_kernels = enqueue_scan(h, nullptr, true);
    VMP_CUDA_CHECK(cudaStreamEndCapture(h->stream, &g));
    VMP_CUDA_CHECK(cudaGraphInstantiate(&h->graph_raw, g, 0));
    VMP_CUDA_CHECK(cudaGraphDestroy(g));
    VMP_CUDA_CHECK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    h->graph_pred_kernels = enqueue_scan(h, nullptr, true, true);
    VMP_CUDA_CHECK(cudaStreamEndCapture(h->stream, &g));
    VMP_CUDA_CHECK(cudaGraphInstantiate(&h->graph_pred, g, 0));
    VMP_CUDA_CHECK(cudaGraphDestroy(g));
    return VMP_OK;
}

// consume the MapOut mailbox of the last map update (the stream has passed it): maintenance, error bits
const MapOut& last_mout(const vmp_handle_t* h) { return h->h_mout[h->map_seq & 1ull]; }
int finish_map(vmp_handle_t* h) {
    const MapOut& o = last_mout(h);
    h->map_pending = false;
    if (o.seq != h->map_seq) { set_error("map update mailbox out of sequence (%llu, expected %llu)", o.seq, h->map_seq); return VMP_ERR_CUDA; }
    if (o.need_maint) h->launches += launch_map_maintenance(h->stream, h->m, h->ctl, h->sm_count, o.need_maint);   // rare; runs before the next update
    return check_device_err(h, o.err);
}
void read_state(vmp_handle_t* h, vmp_state* x, double* P, vmp_scan_stats* stats) {
    const StateOut& o = *h->h_sout;
    if (x) std::memcpy(x, o.x, sizeof(double) * 36);
    if (P) std::memcpy(P, o.P, sizeof(double) * 529);
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->iters = o.iter; stats->converged = o.converged;
        for (int i = 0; i < 8; i++) stats->effect_num[i] = o.effect[i];
    }
}
// everything enqueued so far has finished when this returns
int finish_sync(vmp_handle_t* h) {
    VMP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    prof_collect(h);
    return finish_map(h);
}
// pipelined mode: wait for the posterior of scan `seq` only (the mailbox is written by the solver CTA)
int wait_state(vmp_handle_t* h, unsigned long long seq) {
    const unsigned long long* p = &h->h_sout->seq;
    for (unsigned long spins = 1;; spins++) {
        if (__atomic_load_n(p, __ATOMIC_ACQUIRE) == seq) return VMP_OK;
        if ((spins & 0x3FFF) == 0) {
            const cudaError_t e = cudaStreamQuery(h->stream);
            if (e == cudaSuccess) {
                if (__atomic_load_n(p, __ATOMIC_ACQUIRE) == seq) return VMP_OK;
                set_error("scan %llu finished without a posterior", seq);
                return VMP_ERR_CUDA;
            }
            if (e != cudaErrorNotReady) { set_error("CUDA error while waiting for the posterior: %s", cudaGetErrorString(e)); return VMP_ERR_CUDA; }
        }
        __builtin_ia32_pause();
    }
}
// before any call that looks at device state: let a pipelined map update finish and consume its mailbox
int drain(vmp_handle_t* h) {
    if (!h->map_pending) return VMP_OK;
    return finish_sync(h);
}

template <typename T>
int d2h(vmp_handle h, std::vector<T>& dst, const T* src, size_t count) {
    dst.resize(count);
    if (count == 0) return VMP_OK;
    VMP_CUDA_CHECK(cudaMemcpyAsync(dst.data(), src, sizeof(T) * count, cudaMemcpyDeviceToHost, h->stream));
    return VMP_OK;
}

}  // namespace

extern "C" {

const char* vmp_last_error(void) { return g_err; }
const char* vmp_version(void) { return "vmp_b200 0.1 (sm_100a, fp64, -fmad=false)"; }

void vmp_config_default(vmp_config* c) {
    if (!c) return;
    std::memset(c, 0, sizeof(*c));
    c->opti_max_iter = 5;
    c->na = 0.01; c->ng = 0.01; c->nba = 0.0001; c->nbg = 0.0001;
    c->imu_init_num = 20;
    c->r_il[0] = c->r_il[4] = c->r_il[8] = 1.0;
    c->gravity_align = 1; c->estimate_ext = 0;
    c->scan_resolution = 0.1; c->voxel_size = 0.5;
    c->update_size_thresh = 10; c->max_point_thresh = 100; c->plane_thresh = 0.01;
    c->ranging_cov = 0.04; c->angle_cov = 0.1;
    c->merge_thresh_for_angle = 0.1; c->merge_thresh_for_distance = 0.04;
    c->map_capacity = 100000;
    c->max_points_per_scan = 32768;
    c->device = 0;
}

int vmp_create(const vmp_config* cfg, vmp_handle* out) {
    if (!cfg || !out) { set_error("vmp_create: null argument"); return VMP_ERR_INVALID_ARG; }
    if (cfg->max_points_per_scan < 1 || cfg->map_capacity < 1 || cfg->max_point_thresh < 1 || cfg->update_size_thresh < 1 ||
        cfg->update_size_thresh > cfg->max_point_thresh || cfg->max_point_thresh > 256 || !(cfg->voxel_size > 0.0) || cfg->opti_max_iter < 1 || cfg->opti_max_iter > 8) {
        set_error("vmp_create: invalid configuration (need max_points_per_scan>=1, map_capacity>=1, 1<=update_size_thresh<=max_point_thresh<=256, voxel_size>0, 1<=opti_max_iter<=8)");
        return VMP_ERR_INVALID_ARG;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device >= ndev) {
        cudaGetLastError();
        set_error("vmp_create: no CUDA device %d available (there is no CPU fallback)", cfg->device);
        return VMP_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    VMP_CUDA_CHECK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) {
        set_error("vmp_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", cfg->device, prop.major, prop.minor);
        return VMP_ERR_NO_DEVICE;
    }
    VMP_CUDA_CHECK(cudaSetDevice(cfg->device));
    vmp_handle_t* h = new vmp_handle_t();
    h->cfg = *cfg;
    h->device = cfg->device;
    h->sm_count = prop.multiProcessorCount;
    *out = h;       // so that a failed create can still be destroyed by the caller
    VMP_CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    VMP_CUDA_CHECK(cudaStreamCreateWithFlags(&h->side.st, cudaStreamNonBlocking));
    VMP_CUDA_CHECK(cudaStreamCreateWithFlags(&h->up_st, cudaStreamNonBlocking));
    {
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        void* fn = nullptr;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue64", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) h->write_value64 = fn;
        else cudaGetLastError();
        h->upload_gate = false;
        if (const char* e = getenv("VMP_UPLOAD_CHUNK_MIN")) { const long v = atol(e); if (v >= 4096) h->chunk_min = (size_t)v; }      // test knob: chunked uploads of small scans
        if (const char* e = getenv("VMP_UPLOAD_CHUNKS")) { const int v = atoi(e); if (v >= 2 && v <= 16) h->gate_chunks = v; }
        if (const char* e = getenv("VMP_UPLOAD_TRACE")) h->upload_trace = atoi(e) != 0;
        if (const char* e = getenv("VMP_UPLOAD_GATE")) h->upload_gate = h->write_value64 != nullptr && atoi(e) != 0;       // A/B and test knob: 1 = streamed upload
    }
    for (auto& e : h->side.ev) VMP_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : h->ev_so) VMP_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    VMP_CUDA_CHECK(cudaEventCreate(&h->ev0));
    VMP_CUDA_CHECK(cudaEventCreate(&h->ev1));

    const int nmax = cfg->max_points_per_scan;
    DevMap& m = h->m;
    m.nmax = nmax;
    m.pool = cfg->map_capacity + nmax + 1024;
    m.maxpt = cfg->max_point_thresh; m.upt = cfg->update_size_thresh; m.capacity = cfg->map_capacity;
    m.plane_thresh = cfg->plane_thresh; m.voxel_size = cfg->voxel_size;
    {   // power-of-two voxel size: the key's divisions are exact scalings (vmp_device.cuh voxel_index)
        int ex = 0;
        const double mant = std::frexp(cfg->voxel_size, &ex);
        m.voxel_inv = (mant == 0.5 && ex > -60 && ex < 60) ? 1.0 / cfg->voxel_size : 0.0;
    }
    m.merge_cap = 2048;
    if (const char* e = getenv("VMP_MERGE_CAP")) m.merge_cap = std::max(1, std::min(2048, atoi(e)));      // test knob
    m.merge_r = 4;                  // MERGE_R of vmp_merge.cuh
    if (const char* e = getenv("VMP_MERGE_R")) m.merge_r = std::max(3, std::min(8, atoi(e)));      // test knob: 3 = exact radius (late activations do occur)
    m.merge_max_depth = 7;          // (cascade depths saturate at 7: the depth test of the merge rounds is off unless the test knob sets it)
    if (const char* e = getenv("VMP_MERGE_MAX_DEPTH")) m.merge_max_depth = atoi(e);     // test knob: -1 forces the exact serial redo after the first merge
    m.undo_cap = 65536;             // seven entries per executed event
    DALLOC(m.undo_slot, m.undo_cap); DALLOC(m.undo_rec, (size_t)m.undo_cap * 44);
    m.heavy_points = 160;
    if (const char* e = getenv("VMP_FILL_HEAVY")) m.heavy_points = atoi(e);      // tuning knob (0 = warp path only)
    m.th_angle = cfg->merge_thresh_for_angle; m.th_dist = cfg->merge_thresh_for_distance;
    size_t hs = 1024;
    while (hs < (size_t)m.pool * 4) hs <<= 1;
    m.hmask = (unsigned)(hs - 1);
    DALLOC(m.tkey, hs); DALLOC(m.tval, hs);
    DALLOC(m.hot, (size_t)m.pool * 8); DALLOC(m.ppt, (size_t)m.pool * 6); DALLOC(m.cov, (size_t)m.pool * 36);
    DALLOC(m.center, (size_t)m.pool * 3); DALLOC(m.tp, (size_t)m.pool * 12 * m.maxpt);
    DALLOC(m.skey, m.pool); DALLOC(m.sgroup, m.pool); DALLOC(m.stamp, m.pool);
    DALLOC(m.n_temp, m.pool); DALLOC(m.newly, m.pool); DALLOC(m.born_scan, m.pool); DALLOC(m.full_scan, m.pool); DALLOC(m.full_idx, m.pool);
    DALLOC(m.cnt, m.pool); DALLOC(m.cursor, m.pool); DALLOC(m.ft, m.pool); DALLOC(m.lt, m.pool); DALLOC(m.seg_off, m.pool);
    DALLOC(m.evict_t, m.pool); DALLOC(m.ghost, m.pool); DALLOC(m.evn, m.pool); DALLOC(m.free_slots, m.pool);
    DALLOC(m.pslot, nmax); DALLOC(m.seg, nmax);
    DALLOC(m.touched, nmax); DALLOC(m.newlist, nmax); DALLOC(m.hotlist, nmax); DALLOC(m.vox_cls, nmax); DALLOC(m.ev_slot, nmax); DALLOC(m.ev_time, nmax); DALLOC(m.ev_key, nmax);
    DALLOC(m.ct, nmax); DALLOC(m.act_slot, nmax); DALLOC(m.act_t, nmax);
    VMP_CUDA_CHECK(map_configure_kernels(m));
    const int nblk = (nmax + PT_BLOCK - 1) / PT_BLOCK;
    DALLOC(m.blk_last, nblk); DALLOC(m.blk_new, nblk);
    m.log_cap = std::max<long long>(8ll * m.pool, 8ll * nmax + 4096);
    for (int b = 0; b < 2; b++) { DALLOC(m.log_slot[b], m.log_cap); DALLOC(m.log_stamp[b], m.log_cap); }
    DALLOC(m.log_blk, m.log_cap / 1024 + 2);

    DevScan& s = h->s;
    s.nmax = nmax;
    DALLOC(s.pl, (size_t)3 * nmax); DALLOC(s.cl, (size_t)9 * nmax);
    DALLOC(s.rnorm, (size_t)3 * nmax); DALLOC(s.rmean, (size_t)3 * nmax); DALLOC(s.rres, nmax);
    DALLOC(s.rvalid, nmax); DALLOC(s.rstatus, nmax); DALLOC(s.rkey, nmax); DALLOC(s.rslot, nmax);
    DALLOC(s.pw, (size_t)3 * nmax); DALLOC(s.pcov, (size_t)9 * nmax);
    DALLOC(h->d_stage, PTS_OFF + sizeof(float) * 4 * (size_t)nmax + 64);
    h->d_in = (ScanIn*)h->d_stage;
    DALLOC(s.raw, (size_t)3 * nmax);
    s.range_var = cfg->ranging_cov * cfg->ranging_cov;
    const double sn = sin(cfg->angle_cov * 0.017453293);      // PCL's DEG2RAD (commons.cpp:27-28)
    s.sn2 = sn * sn;
    // ResidualData() value-initialised by vector::resize (lio_builder.cpp:25)
    VMP_CUDA_CHECK(cudaMemsetAsync(s.rnorm, 0, sizeof(double) * 3 * nmax, h->stream));
    VMP_CUDA_CHECK(cudaMemsetAsync(s.rmean, 0, sizeof(double) * 3 * nmax, h->stream));
    VMP_CUDA_CHECK(cudaMemsetAsync(s.rres, 0, sizeof(double) * nmax, h->stream));
    VMP_CUDA_CHECK(cudaMemsetAsync(s.rvalid, 0, nmax, h->stream));
    VMP_CUDA_CHECK(cudaMemsetAsync(s.rstatus, 0, nmax, h->stream));
    VMP_CUDA_CHECK(cudaMemsetAsync(s.pl, 0, sizeof(double) * 3 * nmax, h->stream));
    VMP_CUDA_CHECK(cudaMemsetAsync(s.cl, 0, sizeof(double) * 9 * nmax, h->stream));

    {
        DevDown& d = h->ds;
        DALLOC(d.mm, 8); DALLOC(d.keys[0], nmax); DALLOC(d.keys[1], nmax); DALLOC(d.vals[0], nmax); DALLOC(d.vals[1], nmax);
        DALLOC(d.head, nmax); DALLOC(d.rank, nmax); DALLOC(d.out, nmax); DALLOC(d.m, 1);
        d.temp_bytes = downsample_temp_bytes(nmax);
        unsigned char* tmp = nullptr;
        DALLOC(tmp, d.temp_bytes);
        d.temp = tmp;
        VMP_CUDA_CHECK(cudaMemsetAsync(d.mm, 0, sizeof(unsigned) * 8, h->stream));
        VMP_CUDA_CHECK(cudaHostAlloc((void**)&h->h_ds, sizeof(float4) * (size_t)nmax + 64, cudaHostAllocMapped));
        VMP_CUDA_CHECK(cudaHostGetDevicePointer((void**)&h->a_ds, h->h_ds, 0));
        VMP_CUDA_CHECK(cudaHostAlloc((void**)&h->h_ds_m, 64, cudaHostAllocMapped));
        VMP_CUDA_CHECK(cudaHostGetDevicePointer((void**)&h->a_ds_m, h->h_ds_m, 0));
        *h->h_ds_m = 0;
    }
    DALLOC(h->f, 1); DALLOC(h->ctl, 1); DALLOC(h->d_pred, 1);
    VMP_CUDA_CHECK(cudaMallocHost((void**)&h->h_pred, sizeof(DevPredictIn)));
    std::memset(h->h_pred, 0, sizeof(DevPredictIn));
    VMP_CUDA_CHECK(cudaMemsetAsync(h->f, 0, sizeof(DevFilter), h->stream));
    VMP_CUDA_CHECK(cudaMemsetAsync(h->ctl, 0, sizeof(DevCtl), h->stream));
    h->grid_pts = std::max(1, std::min(h->sm_count * 2, (nmax + 255) / 256));
    {
        const int tpb = cfg->estimate_ext ? 128 : 256;
        const int occ = cfg->estimate_ext ? 4 : 2;       // resident measurement CTAs per SM (k_measure's launch bounds)
        // one slot is left for the solver CTA: with sm_count * occ measurement CTAs the last of them waited for a free slot
        // and ran as a second wave of its own (the measurement of 200 000 points took twice as long as it had to)
        h->grid_meas = std::max(1, std::min(h->sm_count * occ - 1, (nmax + tpb - 1) / tpb));
    }
    h->iekf_loop = cfg->opti_max_iter <= 8 && iekf_loop_fits(cfg->estimate_ext != 0, h->grid_meas, h->sm_count);
    if (const char* e = getenv("VMP_IEKF_LOOP")) h->iekf_loop = h->iekf_loop && atoi(e) != 0;      // A/B knob: 0 = one launch per iteration
    DALLOC(h->partials, (size_t)h->grid_meas * PARTIAL_STRIDE);
    DALLOC(h->meas_out, 160);
    VMP_CUDA_CHECK(cudaMallocHost((void**)&h->h_stage, PTS_OFF + sizeof(float) * 4 * (size_t)nmax + 64));
    h->h_in = (ScanIn*)h->h_stage;
    h->h_raw = (float*)(h->h_stage + PTS_OFF);
    VMP_CUDA_CHECK(cudaHostAlloc((void**)&h->h_cloud, sizeof(float4) * (size_t)nmax + 64, cudaHostAllocMapped));
    VMP_CUDA_CHECK(cudaHostGetDevicePointer((void**)&h->a_cloud, h->h_cloud, 0));
    VMP_CUDA_CHECK(cudaHostAlloc((void**)&h->h_sout, sizeof(StateOut), cudaHostAllocMapped));
    VMP_CUDA_CHECK(cudaHostAlloc((void**)&h->h_mout, 2 * sizeof(MapOut), cudaHostAllocMapped));
    VMP_CUDA_CHECK(cudaHostGetDevicePointer((void**)&h->a_sout, h->h_sout, 0));
    VMP_CUDA_CHECK(cudaHostGetDevicePointer((void**)&h->a_mout, h->h_mout, 0));
    std::memset(h->h_in, 0, sizeof(ScanIn));
    std::memset(h->h_sout, 0, sizeof(StateOut));
    std::memset(h->h_mout, 0, 2 * sizeof(MapOut));
    for (int b = 0; b < 2; b++) { VMP_CUDA_CHECK(cudaEventCreate(&h->pe0[b])); VMP_CUDA_CHECK(cudaEventCreate(&h->pe1[b])); }

    launch_map_init(h->stream, m, h->ctl);
    h->launches += 1;
    const int mi = cfg->opti_max_iter;
    VMP_CUDA_CHECK(cudaMemcpyAsync(&h->ctl->max_iter, &mi, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (const char* e = getenv("VMP_DEBUG_ITER")) {          // diagnostics: which iteration's solver phase cycles to record
        const int di = atoi(e);
        VMP_CUDA_CHECK(cudaMemcpyAsync(&h->ctl->dbg_it, &di, sizeof(int), cudaMemcpyHostToDevice, h->stream));
        VMP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    }
    VMP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    if (const char* e = getenv("VMP_L2_PERSIST_MB")) {
        // experiment: an L2 persisting access-policy window on the hot plane records (mean / normal / flags word, 64 B per
        // voxel: what k_measure gathers per point); set on the stream before the capture, the graph's kernel nodes inherit it
        const size_t want = (size_t)std::max(0, atoi(e)) << 20;
        int max_persist = 0, max_window = 0, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        const size_t persist = std::min(want, (size_t)max_persist);
        if (persist > 0) {
            VMP_CUDA_CHECK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist));
            cudaStreamAttrValue v{};
            const size_t bytes = std::min((size_t)h->m.pool * 64, (size_t)max_window);
            v.accessPolicyWindow.base_ptr = h->m.hot;
            v.accessPolicyWindow.num_bytes = bytes;
            v.accessPolicyWindow.hitRatio = bytes <= persist ? 1.0f : (float)((double)persist / (double)bytes);
            v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            VMP_CUDA_CHECK(cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &v));
        }
    }
    int r = build_graph(h);
    if (r != VMP_OK && h->iekf_loop) {      // the cooperative loop kernel could not be captured on this driver: one launch per iteration instead
        cudaGetLastError();
        h->iekf_loop = false;
        r = build_graph(h);
    }
    if (r) return r;
    VMP_CUDA_CHECK(cudaGetLastError());
    return VMP_OK;
}

int vmp_destroy(vmp_handle h) {
    if (!h) return VMP_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->graph) cudaGraphExecDestroy(h->graph);
    for (void* p : h->allocs) cudaFree(p);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    if (h->h_cloud) cudaFreeHost(h->h_cloud);
    if (h->h_ds) cudaFreeHost(h->h_ds);
    if (h->h_ds_m) cudaFreeHost(h->h_ds_m);
    if (h->graph_raw) cudaGraphExecDestroy(h->graph_raw);
    if (h->graph_pred) cudaGraphExecDestroy(h->graph_pred);
    if (h->h_pred) cudaFreeHost(h->h_pred);
    if (h->h_sout) cudaFreeHost(h->h_sout);
    if (h->h_mout) cudaFreeHost(h->h_mout);
    for (int b = 0; b < 2; b++) { if (h->pe0[b]) cudaEventDestroy(h->pe0[b]); if (h->pe1[b]) cudaEventDestroy(h->pe1[b]); }
    for (auto& e : h->pev) if (e) cudaEventDestroy(e);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (auto& e : h->side.ev) if (e) cudaEventDestroy(e);
    for (auto& e : h->ev_so) if (e) cudaEventDestroy(e);
    if (h->side.st) cudaStreamDestroy(h->side.st);
    if (h->up_st) cudaStreamDestroy(h->up_st);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return VMP_OK;
}

static int check_args(vmp_handle h, int n, const char* who) {
    if (!h) { set_error("%s: null handle", who); return VMP_ERR_INVALID_ARG; }
    if (n < 0 || n > h->cfg.max_points_per_scan) {
        set_error("%s: n=%d outside [0, max_points_per_scan=%d]", who, n, h->cfg.max_points_per_scan);
        return VMP_ERR_INVALID_ARG;
    }
    cudaSetDevice(h->device);
    return VMP_OK;
}
// every entry point but the pipelined scan itself first lets a pending map update finish
static int check_n(vmp_handle h, int n, const char* who) {
    const int r = check_args(h, n, who);
    return r ? r : drain(h);
}

static int upload_n(vmp_handle h, int n) {
    VMP_CUDA_CHECK(cudaMemcpyAsync(&h->ctl->n, &n, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    h->n_last = n;
    return VMP_OK;
}

static int map_update_common(vmp_handle h, const double* pts, const double* cov, int n, vmp_update_stats* st, bool build) {
    int r = check_n(h, n, build ? "vmp_map_build" : "vmp_map_update");
    if (r) return r;
    if (n > 0 && (!pts || !cov)) { set_error("vmp_map_update: null input"); return VMP_ERR_INVALID_ARG; }
    if (build && h->map_built) { set_error("vmp_map_build: the map was already built"); return VMP_ERR_STATE; }
    r = upload_n(h, n);
    if (r) return r;
    VMP_CUDA_CHECK(cudaEventRecord(h->ev0, h->stream));
    if (n > 0) {
        VMP_CUDA_CHECK(cudaMemcpyAsync(h->s.pw, pts, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, h->stream));
        VMP_CUDA_CHECK(cudaMemcpyAsync(h->s.pcov, cov, sizeof(double) * 9 * n, cudaMemcpyHostToDevice, h->stream));
    }
    if (h->prof_on) {                       // per-kernel CUDA events (vmp_profile_*)
        Marker mk{prof_mark, h};
        h->pev_n = 0;
        VMP_CUDA_CHECK(cudaEventRecord(h->pev[0], h->stream));
        h->launches += launch_map_update(h->stream, h->m, h->s, h->f, h->ctl, h->sm_count, build, 2, h->a_mout, &mk, nullptr);
    } else {
        h->launches += launch_map_update(h->stream, h->m, h->s, h->f, h->ctl, h->sm_count, build, 2, h->a_mout, nullptr, &h->side);
    }
    h->map_built = true;
    h->map_seq = h->seq;
    r = finish_sync(h);
    fill_update_stats(last_mout(h).st, st);
    return r;
}

int vmp_map_build(vmp_handle h, const double* pts, const double* cov, int n, vmp_update_stats* st) {
    return map_update_common(h, pts, cov, n, st, true);
}
int vmp_map_update(vmp_handle h, const double* pts, const double* cov, int n, vmp_update_stats* st) {
    return map_update_common(h, pts, cov, n, st, false);
}

int vmp_set_state(vmp_handle h, const vmp_state* x, const double* P) {
    int r = check_n(h, 0, "vmp_set_state");
    if (r) return r;
    if (x) VMP_CUDA_CHECK(cudaMemcpyAsync(h->f->x, x, sizeof(double) * 36, cudaMemcpyHostToDevice, h->stream));
    if (P) VMP_CUDA_CHECK(cudaMemcpyAsync(h->f->P, P, sizeof(double) * 529, cudaMemcpyHostToDevice, h->stream));
    VMP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return VMP_OK;
}
int vmp_get_state(vmp_handle h, vmp_state* x, double* P) {
    int r = check_n(h, 0, "vmp_get_state");
    if (r) return r;
    if (x) VMP_CUDA_CHECK(cudaMemcpyAsync(x, h->f->x, sizeof(double) * 36, cudaMemcpyDeviceToHost, h->stream));
    if (P) VMP_CUDA_CHECK(cudaMemcpyAsync(P, h->f->P, sizeof(double) * 529, cudaMemcpyDeviceToHost, h->stream));
    VMP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return VMP_OK;
}

int vmp_get_prior(vmp_handle h, vmp_state* x, double* P) {
    int r = check_n(h, 0, "vmp_get_prior");
    if (r) return r;
    // the device copy of the last scan's header: uploaded by vmp_scan / vmp_scan_raw, written by k_predict for vmp_scan_raw_predict
    if (x) VMP_CUDA_CHECK(cudaMemcpyAsync(x, h->d_in->x, sizeof(double) * 36, cudaMemcpyDeviceToHost, h->stream));
    if (P) VMP_CUDA_CHECK(cudaMemcpyAsync(P, h->d_in->P, sizeof(double) * 529, cudaMemcpyDeviceToHost, h->stream));
    VMP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return VMP_OK;
}

int vmp_set_scan(vmp_handle h, const float* pts, int n) {
    int r = check_n(h, n, "vmp_set_scan");
    if (r) return r;
    if (n > 0 && !pts) { set_error("vmp_set_scan: null input"); return VMP_ERR_INVALID_ARG; }
    std::memcpy(h->h_raw, pts, sizeof(float) * 3 * (size_t)n);
    h->h_in->pts = (const float*)(h->d_stage + PTS_OFF); h->h_in->prior = nullptr; h->h_in->gate_pts = 0; h->h_in->seq = ++h->seq; h->h_in->n = n; h->h_in->mode = 0;
    h->h_in->n_poses = 0; h->h_in->stride = 3;
    VMP_CUDA_CHECK(cudaMemcpyAsync(h->d_stage, h->h_stage, PTS_OFF + sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    h->n_last = n;
    launch_set_scan(h->stream, h->grid_pts, h->s, h->d_in, h->f, h->ctl);
    h->launches += 1;
    VMP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return VMP_OK;
}

int vmp_measure(vmp_handle h, const vmp_state* x, const double* P, double* H, double* b, int* effect_num) {
    int r = check_n(h, 0, "vmp_measure");
    if (r) return r;
    if (!x || !P) { set_error("vmp_measure: null state"); return VMP_ERR_INVALID_ARG; }
    VMP_CUDA_CHECK(cudaMemcpyAsync(h->f->x, x, sizeof(double) * 36, cudaMemcpyHostToDevice, h->stream));
    VMP_CUDA_CHECK(cudaMemcpyAsync(h->f->P, P, sizeof(double) * 529, cudaMemcpyHostToDevice, h->stream));
    k_reset_iter<<<1, 1, 0, h->stream>>>(h->ctl);
    const bool ext = h->cfg.estimate_ext != 0;
    launch_measure(h->stream, ext, h->grid_meas, h->m, h->s, h->f, h->ctl, h->partials, 0);
    k_reduce_partials<<<1, 160, 0, h->stream>>>(h->partials, h->grid_meas, ext ? 1 : 0, h->meas_out);
    h->launches += 3;
    double out[157];
    VMP_CUDA_CHECK(cudaMemcpyAsync(out, h->meas_out, sizeof(out), cudaMemcpyDeviceToHost, h->stream));
    VMP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    if (H) std::memcpy(H, out, sizeof(double) * 144);
    if (b) std::memcpy(b, out + 144, sizeof(double) * 12);
    if (effect_num) *effect_num = (int)out[156];
    return VMP_OK;
}

// launch the scan that h->h_in describes and collect its results: everything (default), or - pipelined - the posterior
// only, with the map update still running and its counters / errors delivered by the next call
// src != null: the points are still in the caller's (pageable) buffer.  They are staged and uploaded in a few chunks, the DMA copy
// of a chunk running while the helpers stage the next one; check_stride > 0 also checks the time order of a raw scan on the way
// (*sorted).  src == null: everything is in the pinned staging already, one DMA copy.
static int scan_common(vmp_handle h, vmp_state* x, double* P, int n, size_t upload_bytes, vmp_scan_stats* stats, bool raw = false,
                       const void* src = nullptr, int check_stride = 0, bool predict = false) {
    const bool pipe = h->pipelined && !h->prof_on;
    const auto t_call = std::chrono::steady_clock::now();
    const unsigned long long seq = ++h->seq;
    h->h_in->seq = seq; h->h_in->n = n; h->h_in->gate_pts = 0;
    const int eb = (int)(seq & 1);
    VMP_CUDA_CHECK(cudaEventRecord(h->pe0[eb], h->stream));
    if (predict) VMP_CUDA_CHECK(cudaMemcpyAsync(h->d_pred, h->h_pred, sizeof(DevPredictIn), cudaMemcpyHostToDevice, h->stream));
    bool launched = false;
    if (!src) {
        VMP_CUDA_CHECK(cudaMemcpyAsync(h->d_stage, h->h_stage, upload_bytes, cudaMemcpyHostToDevice, h->stream));
    } else {
        const size_t pts_bytes = upload_bytes - PTS_OFF;
        const size_t rec = check_stride > 0 ? sizeof(float) * (size_t)check_stride : sizeof(float) * 3;
        int nchunk = pts_bytes >= 4 * h->chunk_min ? 8 : pts_bytes >= h->chunk_min ? 4 : 1;
        // Plain scans: the helpers stage chunk after chunk WITHOUT waiting for this thread (begin_chunks), which only ships what is ready.
        // STREAMED upload (VMP_UPLOAD_GATE=1, off by default - measured, it does not pay): the graph is launched behind chunk 0; the
        // other chunks travel on their own stream while the first measurement pass is already running, and a stream memory operation behind each
        // of them (DevCtl::up_pub) releases the warps that wait for its points.
        const bool chunked = nchunk > 1 && check_stride == 0 && (h->pool.helpers() >= 3 || h->upload_gate);     // (with one or two helpers the caller's own share of a copy matters: pool.copy)
        const bool gate = chunked && h->upload_gate && !raw && !predict && !h->prof_on;
        if (gate) nchunk = h->gate_chunks;
        size_t nrec_per = (pts_bytes / rec + nchunk - 1) / nchunk;
        if (gate) nrec_per = (nrec_per + 31) / 32 * 32;                 // a warp's 32 points never straddle two chunks
        const size_t per = nrec_per * rec;
        if (gate) h->h_in->gate_pts = (int)nrec_per;
        bool sorted = true;
        VMP_CUDA_CHECK(cudaMemcpyAsync(h->d_stage, h->h_stage, PTS_OFF, cudaMemcpyHostToDevice, h->stream));      // header (+ prior, IMU poses)
        if (chunked) {
            typedef int (*WriteValue64)(cudaStream_t, unsigned long long, unsigned long long, unsigned);
            const WriteValue64 write_value = reinterpret_cast<WriteValue64>(h->write_value64);
            const bool trace = h->upload_trace && (seq & 15ull) == 0;
            cudaStream_t later = gate ? h->up_st : h->stream;
            const auto tr0 = std::chrono::steady_clock::now();
            float tr_ready[16] = {}, tr_enq[16] = {}, tr_graph = 0.f;
            auto tr_us = [&] { return std::chrono::duration<float, std::micro>(std::chrono::steady_clock::now() - tr0).count(); };
            h->pool.begin_chunks(h->h_raw, src, pts_bytes, per);
            int c = 0, rr = VMP_OK;
            cudaError_t ce = cudaSuccess;
            for (size_t off = 0; off < pts_bytes && rr == VMP_OK && ce == cudaSuccess; off += per, c++) {
                const size_t len = std::min(per, pts_bytes - off);
                h->pool.wait_chunk(c);
                if (trace) tr_ready[c] = tr_us();
                ce = cudaMemcpyAsync(h->d_stage + PTS_OFF + off, h->h_stage + PTS_OFF + off, len, cudaMemcpyHostToDevice, c ? later : h->stream);
                if (ce != cudaSuccess) break;
                if (gate && c == 0) {
                    ce = cudaStreamWaitEvent(h->up_st, h->pe0[eb], 0);     // (pipelined mode: the previous scan's map update still reads the old points)
                    if (ce == cudaSuccess) rr = run_scan(h, raw, n, predict);
                    launched = true;
                    if (trace) tr_graph = tr_us();
                } else if (gate && write_value(h->up_st, (unsigned long long)(uintptr_t)&h->ctl->up_pub, seq * 64ull + (unsigned long long)c, 0u) != 0) {
                    set_error("vmp_scan: cuStreamWriteValue64 failed during the streamed upload");
                    h->upload_gate = false;
                    rr = VMP_ERR_CUDA;
                }
                if (trace) tr_enq[c] = tr_us();
            }
            h->pool.end_chunks();
            if (trace) {
                std::fprintf(stderr, "[vmp upload trace] seq %llu, %d chunks of %zu B; graph launched %.1f us; chunk ready / shipped (us):", seq, c, per, tr_graph);
                for (int q = 0; q < c; q++) std::fprintf(stderr, " %.1f/%.1f", tr_ready[q], tr_enq[q]);
                std::fprintf(stderr, "\n");
            }
            if (ce != cudaSuccess || rr != VMP_OK) {        // a launched scan gives up by itself after 2 s (E_UPLOAD)
                cudaStreamSynchronize(h->stream);
                if (ce != cudaSuccess) { set_error("vmp_scan: streamed upload failed: %s", cudaGetErrorString(ce)); return VMP_ERR_CUDA; }
                return rr;
            }
        } else
        for (size_t off = 0; off < pts_bytes; off += per) {
            const size_t len = std::min(per, pts_bytes - off);
            sorted &= h->pool.copy((char*)h->h_raw + off, (const char*)src + off, len, check_stride);
            if (check_stride > 0 && off > 0) {      // order across the chunk boundary
                const float* f = reinterpret_cast<const float*>((const char*)src + off);
                sorted &= !(f[check_stride - 1] < f[-1]);
            }
            VMP_CUDA_CHECK(cudaMemcpyAsync(h->d_stage + PTS_OFF + off, h->h_stage + PTS_OFF + off, len, cudaMemcpyHostToDevice, h->stream));
        }
        if (!sorted) {      // lio_builder.cpp:75; rare for real sensors (points arrive in time order); any order of equal keys is a valid std::sort result
            struct P4 { float x, y, z, t; };
            P4* q = reinterpret_cast<P4*>(h->h_raw);
            std::stable_sort(q, q + n, [](const P4& a, const P4& b) { return a.t < b.t; });
            VMP_CUDA_CHECK(cudaMemcpyAsync(h->d_stage + PTS_OFF, h->h_stage + PTS_OFF, pts_bytes, cudaMemcpyHostToDevice, h->stream));
        }
    }
    if (!launched) { const int rr = run_scan(h, raw, n, predict); if (rr) return rr; }
    VMP_CUDA_CHECK(cudaEventRecord(h->pe1[eb], h->stream));
    h->n_last = n;
    if (!pipe) {
        h->map_seq = seq;
        int r = finish_sync(h);
        if (h->upload_trace && (seq & 15ull) == 0) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, h->pe0[eb], h->pe1[eb]);
            std::fprintf(stderr, "[vmp upload trace] seq %llu: scan finished %.1f us after the call started; first event -> end of graph on the device %.1f us\n", seq,
                         std::chrono::duration<float, std::micro>(std::chrono::steady_clock::now() - t_call).count(), ms * 1000.f);
        }
        read_state(h, x, P, stats);
        if (stats) {
            fill_update_stats(last_mout(h).st, &stats->map);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, h->pe0[eb], h->pe1[eb]);
            stats->gpu_ms = ms;
        }
        return r;
    }
    int r = wait_state(h, seq);
    if (r) return r;
    read_state(h, x, P, stats);
    // the stream has passed the previous scan's map update: its mailbox (the other parity) is complete; this scan's update
    // writes its own mailbox, so nothing it does can tear what is read here
    if (h->map_pending) {
        fill_update_stats(last_mout(h).st, &h->lag_map);
        r = finish_map(h);
        if (stats) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, h->pe0[eb ^ 1], h->pe1[eb ^ 1]) == cudaSuccess) stats->gpu_ms = ms;
        }
    }
    if (stats) stats->map = h->lag_map;
    h->map_pending = true;
    h->map_seq = seq;
    return r;
}

// the scan is already in the pinned staging buffer (vmp_scan_buffer): header + prior + points go up in ONE DMA copy (SM
// reads of host memory reach a fraction of the copy engine's PCIe rate, small ones cost a round trip each)
static int scan_staged(vmp_handle h, vmp_state* x, double* P, int n, vmp_scan_stats* stats, const char* who, const float* src = nullptr) {
    if (!x || !P) { set_error("%s: null argument", who); return VMP_ERR_INVALID_ARG; }
    if (!h->map_built) { set_error("%s: no map yet (call vmp_first_scan or vmp_map_build first)", who); return VMP_ERR_STATE; }
    h->h_in->pts = (const float*)(h->d_stage + PTS_OFF); h->h_in->prior = nullptr; h->h_in->gate_pts = 0; h->h_in->mode = SCAN_STATE_HDR | SCAN_BEGIN_UPDATE;
    h->h_in->n_poses = 0; h->h_in->stride = 3;
    std::memcpy(h->h_in->x, x, sizeof(double) * 36);
    std::memcpy(h->h_in->P, P, sizeof(double) * 529);
    return scan_common(h, x, P, n, PTS_OFF + sizeof(float) * 3 * (size_t)n, stats, false, src);
}

int vmp_scan(vmp_handle h, vmp_state* x, double* P, const float* pts, int n, vmp_scan_stats* stats) {
    const auto t_enter = std::chrono::steady_clock::now();
    int r = h && h->pipelined && !h->prof_on ? check_args(h, n, "vmp_scan") : check_n(h, n, "vmp_scan");
    if (r) return r;
    if (n > 0 && !pts) { set_error("vmp_scan: null argument"); return VMP_ERR_INVALID_ARG; }
    r = scan_staged(h, x, P, n, stats, "vmp_scan", n > 0 ? pts : nullptr);          // staged + uploaded chunk by chunk
    if (stats) stats->host_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_enter).count();
    return r;
}

float* vmp_scan_buffer(vmp_handle h) { return h ? h->h_raw : nullptr; }

int vmp_scan_buffer_fill(vmp_handle h, const float* src, int stride_floats, int n) {
    const int r = check_n(h, n, "vmp_scan_buffer_fill");
    if (r) return r;
    if ((n > 0 && !src) || stride_floats < 3) { set_error("vmp_scan_buffer_fill: invalid argument"); return VMP_ERR_INVALID_ARG; }
    h->pool.gather_xyz(h->h_raw, src, stride_floats, (size_t)n);
    return VMP_OK;
}

int vmp_scan_staged(vmp_handle h, vmp_state* x, double* P, int n, vmp_scan_stats* stats) {
    const auto t_enter = std::chrono::steady_clock::now();
    int r = h && h->pipelined && !h->prof_on ? check_args(h, n, "vmp_scan_staged") : check_n(h, n, "vmp_scan_staged");
    if (r) return r;
    r = scan_staged(h, x, P, n, stats, "vmp_scan_staged");
    if (stats) stats->host_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_enter).count();
    return r;
}

int vmp_scan_dev(vmp_handle h, const float* pts_dev, const double* prior_dev, int n, vmp_scan_stats* stats) {
    const auto t_enter = std::chrono::steady_clock::now();
    int r = h && h->pipelined && !h->prof_on ? check_args(h, n, "vmp_scan_dev") : check_n(h, n, "vmp_scan_dev");
    if (r) return r;
    if (n > 0 && !pts_dev) { set_error("vmp_scan_dev: null points"); return VMP_ERR_INVALID_ARG; }
    if (!h->map_built) { set_error("vmp_scan_dev: no map yet"); return VMP_ERR_STATE; }
    h->h_in->pts = pts_dev; h->h_in->prior = prior_dev;
    h->h_in->mode = (prior_dev ? SCAN_STATE_DEV : 0) | SCAN_BEGIN_UPDATE;
    h->h_in->n_poses = 0; h->h_in->stride = 3;
    r = scan_common(h, nullptr, nullptr, n, offsetof(ScanIn, x), stats);
    if (stats) stats->host_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_enter).count();
    return r;
}

int vmp_scan_raw(vmp_handle h, vmp_state* x, double* P, float* cloud_xyzt, int n, const vmp_pose* poses, int n_poses, vmp_scan_stats* stats) {
    const auto t_enter = std::chrono::steady_clock::now();
    int r = h && h->pipelined && !h->prof_on ? check_args(h, n, "vmp_scan_raw") : check_n(h, n, "vmp_scan_raw");
    if (r) return r;
    if (!x || !P || (n > 0 && !cloud_xyzt) || !poses) { set_error("vmp_scan_raw: null argument"); return VMP_ERR_INVALID_ARG; }
    if (n_poses < 2 || n_poses > MAX_POSES) { set_error("vmp_scan_raw: n_poses=%d outside [2, %d]", n_poses, MAX_POSES); return VMP_ERR_INVALID_ARG; }
    if (!h->map_built) { set_error("vmp_scan_raw: no map yet (call vmp_first_scan or vmp_map_build first)"); return VMP_ERR_STATE; }
    static_assert(sizeof(vmp_pose) == sizeof(DevPose), "vmp_pose layout");
    // the caller's cloud is staged, checked for time order (lio_builder.cpp:75) and uploaded chunk by chunk in scan_common
    std::memcpy(h->h_stage + IN_HDR, poses, sizeof(vmp_pose) * (size_t)n_poses);
    h->h_in->pts = (const float*)(h->d_stage + PTS_OFF); h->h_in->prior = nullptr; h->h_in->gate_pts = 0; h->h_in->mode = SCAN_STATE_HDR | SCAN_BEGIN_UPDATE;
    h->h_in->n_poses = n_poses; h->h_in->stride = 4;
    std::memcpy(h->h_in->x, x, sizeof(double) * 36);
    std::memcpy(h->h_in->P, P, sizeof(double) * 529);
    r = scan_common(h, x, P, n, PTS_OFF + sizeof(float) * 4 * (size_t)n, stats, true, n > 0 ? cloud_xyzt : nullptr, 4);
    // the compensated cloud was written to mapped host memory by the first kernel of the graph, before the posterior; it
    // stays available through vmp_get_lidar_cloud, the copy into the caller's buffer can be switched off (vmp_set_raw_writeback)
    if (h->raw_writeback) h->pool.copy(cloud_xyzt, h->h_cloud, sizeof(float) * 4 * (size_t)n);
    if (stats) stats->host_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_enter).count();
    return r;
}

int vmp_scan_raw_predict(vmp_handle h, vmp_state* x_out, double* P_out, float* cloud_xyzt, int n, const vmp_imu_step* steps, int n_steps,
                         const double* Q, const double* last_acc_gyro, vmp_scan_stats* stats) {
    const auto t_enter = std::chrono::steady_clock::now();
    int r = h && h->pipelined && !h->prof_on ? check_args(h, n, "vmp_scan_raw_predict") : check_n(h, n, "vmp_scan_raw_predict");
    if (r) return r;
    if (!x_out || !P_out || (n > 0 && !cloud_xyzt) || !steps || !Q) { set_error("vmp_scan_raw_predict: null argument"); return VMP_ERR_INVALID_ARG; }
    if (n_steps < 1 || n_steps > MAX_POSES) { set_error("vmp_scan_raw_predict: n_steps=%d outside [1, %d]", n_steps, MAX_POSES); return VMP_ERR_INVALID_ARG; }
    if (!h->map_built) { set_error("vmp_scan_raw_predict: no map yet (call vmp_first_scan or vmp_map_build first)"); return VMP_ERR_STATE; }
    static_assert(sizeof(vmp_imu_step) == sizeof(DevImuStep), "vmp_imu_step layout");
    std::memcpy(h->h_pred->steps, steps, sizeof(vmp_imu_step) * (size_t)n_steps);
    std::memcpy(h->h_pred->Q, Q, sizeof(double) * 144);
    h->h_pred->n_steps = n_steps;
    h->h_pred->use_last = last_acc_gyro ? 1 : 0;
    if (last_acc_gyro) std::memcpy(h->h_pred->last, last_acc_gyro, sizeof(double) * 6);
    // the prior (x, P) and the IMU poses are written into the device copy of the header / staging by k_predict
    h->h_in->pts = (const float*)(h->d_stage + PTS_OFF); h->h_in->prior = nullptr; h->h_in->gate_pts = 0; h->h_in->mode = SCAN_STATE_HDR | SCAN_BEGIN_UPDATE | SCAN_PREDICT;
    h->h_in->n_poses = 0; h->h_in->stride = 4;
    r = scan_common(h, x_out, P_out, n, PTS_OFF + sizeof(float) * 4 * (size_t)n, stats, true, n > 0 ? cloud_xyzt : nullptr, 4, true);
    if (h->raw_writeback) h->pool.copy(cloud_xyzt, h->h_cloud, sizeof(float) * 4 * (size_t)n);
    if (stats) stats->host_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_enter).count();
    return r;
}

int vmp_downsample(vmp_handle h, const float* cloud_xyzc, int n, double leaf, float* out_xyzc, int cap, int* m) {
    int r = check_n(h, n, "vmp_downsample");
    if (r) return r;
    if ((n > 0 && !cloud_xyzc) || !(leaf > 0.0) || cap < 0 || (cap > 0 && !out_xyzc)) { set_error("vmp_downsample: invalid argument"); return VMP_ERR_INVALID_ARG; }
    std::memcpy(h->h_raw, cloud_xyzc, sizeof(float) * 4 * (size_t)n);
    h->h_in->n = n; h->h_in->stride = 4; h->h_in->n_poses = 0; h->h_in->mode = 0;
    h->h_in->pts = (const float*)(h->d_stage + PTS_OFF); h->h_in->prior = nullptr; h->h_in->gate_pts = 0;
    VMP_CUDA_CHECK(cudaMemcpyAsync(h->d_stage, h->h_stage, PTS_OFF + sizeof(float) * 4 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    h->launches += launch_downsample(h->stream, h->ds, (const float4*)(h->d_stage + PTS_OFF), &h->d_in->n, n, (float)leaf, h->grid_pts,
                                     h->a_ds, h->a_ds_m, nullptr, nullptr);
    VMP_CUDA_CHECK(cudaGetLastError());
    VMP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    const int mm = *h->h_ds_m;
    if (m) *m = mm;
    if (cap > 0) std::memcpy(out_xyzc, h->h_ds, sizeof(float) * 4 * (size_t)std::min(mm, cap));
    return VMP_OK;
}
int vmp_get_lidar_cloud(vmp_handle h, float* out_xyzc, int cap, int* m) {
    int r = check_n(h, 0, "vmp_get_lidar_cloud");
    if (r) return r;
    if (cap < 0 || (cap > 0 && !out_xyzc)) { set_error("vmp_get_lidar_cloud: invalid argument"); return VMP_ERR_INVALID_ARG; }
    if (!h->last_raw) { set_error("vmp_get_lidar_cloud: the last scan was not a vmp_scan_raw (the caller holds the filter input)"); return VMP_ERR_STATE; }
    // with the downsample on, the filter input of the last vmp_scan_raw are the leaf centroids; otherwise the compensated cloud
    const int mm = h->ds_valid ? *h->h_ds_m : h->n_last;
    const float* src = h->ds_valid ? reinterpret_cast<const float*>(h->h_ds) : reinterpret_cast<const float*>(h->h_cloud);
    if (m) *m = mm;
    if (cap > 0) std::memcpy(out_xyzc, src, sizeof(float) * 4 * (size_t)std::min(mm, cap));
    return VMP_OK;
}

int vmp_set_pipelined(vmp_handle h, int on) {
    int r = check_n(h, 0, "vmp_set_pipelined");
    if (r) return r;
    h->pipelined = on != 0;
    return VMP_OK;
}
int vmp_sync(vmp_handle h) { return check_n(h, 0, "vmp_sync"); }
int vmp_set_raw_writeback(vmp_handle h, int on) {
    int r = check_n(h, 0, "vmp_set_raw_writeback");
    if (r) return r;
    h->raw_writeback = on != 0;
    return VMP_OK;
}

int vmp_first_scan(vmp_handle h, const vmp_state* x, const double* P, const float* pts, int n, vmp_update_stats* st) {
    int r = check_n(h, n, "vmp_first_scan");
    if (r) return r;
    if (!x || !P || (n > 0 && !pts)) { set_error("vmp_first_scan: null argument"); return VMP_ERR_INVALID_ARG; }
    if (h->map_built) { set_error("vmp_first_scan: the map was already built"); return VMP_ERR_STATE; }
    r = upload_n(h, n);
    if (r) return r;
    VMP_CUDA_CHECK(cudaMemcpyAsync(h->f->x, x, sizeof(double) * 36, cudaMemcpyHostToDevice, h->stream));
    VMP_CUDA_CHECK(cudaMemcpyAsync(h->f->P, P, sizeof(double) * 529, cudaMemcpyHostToDevice, h->stream));
    if (n > 0) VMP_CUDA_CHECK(cudaMemcpyAsync(h->s.raw, pts, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, h->stream));
    VMP_CUDA_CHECK(cudaEventRecord(h->ev0, h->stream));
    h->launches += launch_map_update(h->stream, h->m, h->s, h->f, h->ctl, h->sm_count, true, 1, h->a_mout, nullptr, &h->side);
    h->map_built = true;
    h->map_seq = h->seq;
    r = finish_sync(h);
    fill_update_stats(last_mout(h).st, st);
    return r;
}

// ---------------------------------------------------------------- observation hooks
int vmp_dump_correspondences(vmp_handle h, int64_t* keys, uint8_t* status, double* residual, double* plane_norm, int n) {
    int r = check_n(h, n, "vmp_dump_correspondences");
    if (r) return r;
    const size_t NM = (size_t)h->s.nmax;
    std::vector<unsigned long long> k; std::vector<uint8_t> s; std::vector<double> res, nx, ny, nz;
    if ((r = d2h(h, k, h->s.rkey, n)) || (r = d2h(h, s, h->s.rstatus, n)) || (r = d2h(h, res, h->s.rres, n)) ||
        (r = d2h(h, nx, h->s.rnorm, n)) || (r = d2h(h, ny, h->s.rnorm + NM, n)) || (r = d2h(h, nz, h->s.rnorm + 2 * NM, n))) return r;
    VMP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < n; i++) {
        if (keys) {
            long long x = 0, y = 0, z = 0;
            if (k[i] != KEY_EMPTY) unpack_key(k[i], x, y, z);
            keys[3 * i] = x; keys[3 * i + 1] = y; keys[3 * i + 2] = z;
        }
        if (status) status[i] = s[i];
        if (residual) residual[i] = res[i];
        if (plane_norm) { plane_norm[3 * i] = nx[i]; plane_norm[3 * i + 1] = ny[i]; plane_norm[3 * i + 2] = nz[i]; }
    }
    return VMP_OK;
}

int vmp_dump_world_points(vmp_handle h, double* pts, double* cov, int n) {
    int r = check_n(h, n, "vmp_dump_world_points");
    if (r) return r;
    if (pts && n) VMP_CUDA_CHECK(cudaMemcpyAsync(pts, h->s.pw, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, h->stream));
    if (cov && n) VMP_CUDA_CHECK(cudaMemcpyAsync(cov, h->s.pcov, sizeof(double) * 9 * n, cudaMemcpyDeviceToHost, h->stream));
    VMP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return VMP_OK;
}

int vmp_map_size(vmp_handle h, int* count) {
    int r = check_n(h, 0, "vmp_map_size");
    if (r) return r;
    VMP_CUDA_CHECK(cudaMemcpyAsync(count, &h->ctl->n_live, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    VMP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return VMP_OK;
}

int vmp_dump_map(vmp_handle h, vmp_plane* out, int cap, int* count) {
    int r = check_n(h, 0, "vmp_dump_map");
    if (r) return r;
    if (cap < 0 || (cap > 0 && !out)) { set_error("vmp_dump_map: invalid argument"); return VMP_ERR_INVALID_ARG; }
    const DevMap& m = h->m;
    if (!h->dump_ready) {
        DevDump& d = h->dump;
        const size_t P = (size_t)m.pool;
        DALLOC(d.keys[0], P); DALLOC(d.keys[1], P); DALLOC(d.vals[0], P); DALLOC(d.vals[1], P); DALLOC(d.count, 1);
        DALLOC(d.records, P * 61);
        d.temp_bytes = dump_temp_bytes(m.pool);
        unsigned char* tmp = nullptr;
        DALLOC(tmp, d.temp_bytes);
        d.temp = tmp;
        h->dump_ready = true;
    }
    // the walk over VoxelMap::cache (utils.cpp:161-195) on the device: collect live slots, sort by LRU stamp (front of the
    // list = most recent insertion first), assemble the records, one copy of live x sizeof(vmp_plane) bytes
    const int grid = h->sm_count * 4;
    launch_dump_collect(h->stream, grid, m, h->dump);
    int n_live = 0;
    VMP_CUDA_CHECK(cudaMemcpyAsync(&n_live, h->dump.count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    VMP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    if (count) *count = n_live;
    const int n_out = std::min(n_live, cap);
    launch_dump_sort_gather(h->stream, grid, m, h->dump, n_live);
    h->launches += 3;
    if (n_out > 0) VMP_CUDA_CHECK(cudaMemcpyAsync(out, h->dump.records, sizeof(vmp_plane) * (size_t)n_out, cudaMemcpyDeviceToHost, h->stream));
    VMP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    VMP_CUDA_CHECK(cudaGetLastError());
    return VMP_OK;
}

int vmp_dump_evicted(vmp_handle h, int64_t* keys, int cap, int* count) {
    int r = check_n(h, 0, "vmp_dump_evicted");
    if (r) return r;
    int ne = 0;
    VMP_CUDA_CHECK(cudaMemcpyAsync(&ne, &h->ctl->n_evict, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    VMP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    std::vector<unsigned long long> k;
    if ((r = d2h(h, k, h->m.ev_key, ne))) return r;
    VMP_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < ne && i < cap; i++) {
        long long x, y, z;
        unpack_key(k[i], x, y, z);
        keys[3 * i] = x; keys[3 * i + 1] = y; keys[3 * i + 2] = z;
    }
    if (count) *count = ne;
    return VMP_OK;
}

int64_t vmp_launch_count(vmp_handle h) { return h ? h->launches : 0; }
int vmp_debug_counters(vmp_handle h, int* out8) {
    if (!h || !out8) return VMP_ERR_INVALID_ARG;
    if (drain(h)) return VMP_ERR_CUDA;
    for (int q = 0; q < 8; q++) out8[q] = last_mout(h).dbg[q];
    return VMP_OK;
}

int vmp_profile_enable(vmp_handle h, int on) {
    int r = check_n(h, 0, "vmp_profile_enable");
    if (r) return r;
    if (on && h->pev.empty()) {
        h->pev.resize(64); h->pev_id.assign(64, 0);
        for (auto& e : h->pev) VMP_CUDA_CHECK(cudaEventCreate(&e));
    }
    h->prof_on = on != 0;
    return VMP_OK;
}
int vmp_profile_reset(vmp_handle h) {
    if (!h) return VMP_ERR_INVALID_ARG;
    for (int k = 0; k < VMP_K_COUNT; k++) { h->prof_ms[k] = 0.0; h->prof_cnt[k] = 0; }
    return VMP_OK;
}
int vmp_profile_read(vmp_handle h, double* ms, int64_t* launches) {
    if (!h) return VMP_ERR_INVALID_ARG;
    for (int k = 0; k < VMP_K_COUNT; k++) { if (ms) ms[k] = h->prof_ms[k]; if (launches) launches[k] = h->prof_cnt[k]; }
    return VMP_OK;
}
const char* vmp_kernel_name(int id) {
    static const char* names[VMP_K_COUNT] = {
        "k_scan_in", "k_set_scan", "k_predict", "k_iekf_loop", "k_ieskf_solve", "k_world_insert_count",
        "k_map_begin", "k_map_insert", "k_map_count", "k_seg_scan", "k_seg_fill", "k_lru_evict",
        "k_fill", "k_merge_prefilter", "k_merge_rounds", "k_log_append", "k_map_finalize",
        "k_map_end", "k_rehash", "k_log_compact", "k_scan_out", "k_fill_classify", "k_fill_heavy", "k_undistort", "k_downsample"};
    return (id >= 0 && id < VMP_K_COUNT) ? names[id] : "?";
}

}  // extern "C"
