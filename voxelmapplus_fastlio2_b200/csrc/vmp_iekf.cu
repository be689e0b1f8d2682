// vmp_iekf.cu — IEKF measurement update on the device (sm_100a, fp64, no FMA contraction).
//
//   k_set_scan      lio_builder.cpp:224-229 + calcBodyCov (commons.cpp:18-45)
//   k_measure       LIOBuilder::sharedUpdateFunc (lio_builder.cpp:250-311) +
//                   VoxelMap::index / featmap.find / buildResidual (voxel_map.cpp:194-198,258-276):
//                   one thread per point, gather through the voxel hash; H / b accumulated by the warp
//                   (one entry per lane over the points staged in shared memory) -> per-block
//                   partials (no fp atomics).  Block 0 of the launch is the solver CTA (vmp_solve.cuh):
//                   IESKF::update body (ieskf.cpp:134-155) through the matrix-inversion lemma, fixed-order
//                   reduction of the partials, boxplus / boxminus, convergence flag, posterior covariance
//   k_undistort     point loop of undistortCloud (lio_builder.cpp:127-152)
//   k_state_out     posterior -> mapped host mailbox
//   (lidarToWorld + the pv_list loop, lio_builder.cpp:155-163, 231-245, open the map update: k_world_insert_count in vmp_map.cu)
//
// Why no tensor cores: per point this is a 64-byte gather and ~1 kflop of 3x3 fp64
// algebra followed by a 27-value reduction; there is no dense contraction to tile.
#include "vmp_device.cuh"
#include "vmp_kernels.h"
#include "vmp_state.cuh"

namespace vmp {

// ---------------------------------------------------------------------------- K0
// One kernel stages the whole scan: the header and the points are read straight from the mailbox (pinned host memory or
// the caller's device buffer), block 0 stages the prior and the counters of IESKF::update (ieskf.cpp:127-130).
__global__ void __launch_bounds__(256) k_set_scan(DevScan s, const ScanIn* __restrict__ in, DevFilter* f, DevCtl* ctl) {
    __shared__ float sh[1024];
    __shared__ int s_n, s_mode, s_stride;
    __shared__ const float* s_pts;
    __shared__ const double* s_prior;
    const int tid = threadIdx.x;
    if (tid == 0) { s_n = in->n; s_mode = in->mode; s_pts = in->pts; s_prior = in->prior; s_stride = in->stride == 4 ? 4 : 3; }
    __syncthreads();
    const int n = s_n, mode = s_mode, stride = s_stride;
    const float* pts = s_pts;                      // == s.raw when the host copied the points there itself
    const bool copy_raw = pts != s.raw;
    if (blockIdx.x == 0) {
        if (tid == 0) {
            ctl->n = n; ctl->seq = in->seq;
            if (mode & SCAN_BEGIN_UPDATE) { ctl->iter = 0; ctl->done = 0; ctl->converged = 0; ctl->ticket = 0; }
        }
        if ((mode & SCAN_BEGIN_UPDATE) && tid < 8) ctl->effect[tid] = 0;
        if (mode & (SCAN_STATE_HDR | SCAN_STATE_DEV)) {
            const double* xs = (mode & SCAN_STATE_HDR) ? in->x : s_prior;
            const double* Ps = (mode & SCAN_STATE_HDR) ? in->P : s_prior + 36;
            for (int q = tid; q < 529; q += 256) f->P[q] = Ps[q];
            if (tid < 36) { const double v = xs[tid]; f->x[tid] = v; f->xpred[tid] = v; }
        } else if ((mode & SCAN_BEGIN_UPDATE) && tid < 36) {
            f->xpred[tid] = f->x[tid];
        }
    }
    const bool al16 = (((size_t)pts) & 15) == 0;
    for (int b = blockIdx.x; b * 256 < n; b += gridDim.x) {
        const int base = b * 256;
        const int cnt = n - base < 256 ? n - base : 256, nfl = cnt * stride;
        const float* src = pts + (size_t)base * stride;              // 3072 b or 4096 b bytes past pts: keeps the 16-byte alignment
        if (al16) {
            const int n4 = nfl >> 2;
            if (tid < n4) reinterpret_cast<float4*>(sh)[tid] = reinterpret_cast<const float4*>(src)[tid];
            if (tid < (nfl & 3)) sh[n4 * 4 + tid] = src[n4 * 4 + tid];
        } else {
            for (int q = tid; q < nfl; q += 256) sh[q] = src[q];
        }
        __syncthreads();
        if (copy_raw) for (int q = tid; q < cnt * 3; q += 256) s.raw[(size_t)base * 3 + q] = sh[(q / 3) * stride + q % 3];
        if (tid < cnt) {
            const int i = base + tid;
            V3 p = v3((double)sh[stride * tid], (double)sh[stride * tid + 1], (double)sh[stride * tid + 2]);
            M3 c;
            calc_body_cov(p, s.range_var, s.sn2, c);
#pragma unroll
            for (int k = 0; k < 3; k++) s.pl[(size_t)k * s.nmax + i] = p[k];
#pragma unroll
            for (int k = 0; k < 9; k++) s.cl[(size_t)k * s.nmax + i] = c.a[k];
        }
        __syncthreads();
    }
}

#include "vmp_solve.cuh"

// ---------------------------------------------------------------------------- K1
struct MeasState {
    M3 r_wl, R, Rext, Prr, Ppp;
    V3 p_wl, pext;
};

// Shared memory of a measurement CTA: the pose products every point needs, the per-warp staging of (J, w J, residual) for
// the lane-parallel accumulation, and the per-warp sums.  SP is odd, so 64-bit accesses of a half-warp never share a bank.
template <bool EXT>
struct MeasShared {
    static constexpr int D = EXT ? 12 : 6;
    static constexpr int NV = D * (D + 1) / 2 + D + 1;
    static constexpr int SP = 2 * D + 1;
    static constexpr int NW = (EXT ? 128 : 256) / 32;
    MeasState ms;
    double red[NW][NV];
    double stage[NW][32 * SP];
};

// Streamed upload (ScanIn::gate_pts > 0): the points of the scan arrive chunk by chunk while the first measurement pass is already running.
// A warp about to read points of chunk `need` (>= 1) waits until DevCtl::up_pub - written in stream order behind that chunk's DMA copy - says
// it has landed.  Returns the number of chunks (after the first) known to be there.  The wait only delays loads: sums and their order are untouched.
__device__ __forceinline__ unsigned long long wait_upload(DevCtl* ctl, unsigned long long base, unsigned long long need) {
    unsigned long long v = 0;
    if ((threadIdx.x & 31) == 0) {
        unsigned long long t0 = 0;
        for (;;) {
            v = *(volatile unsigned long long*)&ctl->up_pub;        // relaxed poll (an acquire at system scope per poll is a MEMBAR.SYS per poll, on 2 000 warps)
            if (v >= base + need && v < base + 64ull) { __threadfence(); break; }     // the loads of the points stay behind the flag
            __nanosleep(400);
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (!t0) t0 = t;
            else if (t - t0 > 2000000000ull) { atomicOr(&ctl->err, E_UPLOAD); v = base + 63ull; break; }    // never hang the device on a host that gave up
        }
    }
    v = __shfl_sync(0xffffffffu, v, 0);
    return v - base;
}

// One measurement pass of one CTA (pb of npb) over the scan: LIOBuilder::sharedUpdateFunc (lio_builder.cpp:250-311) for its points,
// partial sums of H^T R^-1 H / H^T R^-1 z / effect count into partials[pb].
template <bool EXT, bool FIRST>
__device__ __forceinline__ void measure_pass(MeasShared<EXT>& sh, const DevMap& m, const DevScan& s, DevFilter* f, DevCtl* ctl, double* partials, int pb, int npb,
                                             const ScanIn* __restrict__ in, int reuse_slots) {
    constexpr int D = EXT ? 12 : 6;
    constexpr int NH = D * (D + 1) / 2;
    constexpr int NV = NH + D + 1;                 // upper triangle of H, b, effect count
    constexpr int NA = NV - 1;                     // accumulated products
    constexpr int SP = MeasShared<EXT>::SP;
    constexpr int NW = MeasShared<EXT>::NW;
    constexpr int VPL = (NA + 31) / 32;            // accumulated values per lane
    MeasState& ms = sh.ms;
    // the pose products and the two covariance blocks every point needs, one entry per thread (same evaluation order
    // as mul(): s = a0 b0; s += a1 b1; s += a2 b2)
    const double* xsrc = f->x;
    const double* Psrc = f->P;
    if (FIRST) {
        const int mode = in->mode;
        if (mode & SCAN_STATE_HDR) { xsrc = in->x; Psrc = in->P; }
        else if (mode & SCAN_STATE_DEV) { xsrc = in->prior; Psrc = in->prior + 36; }
    }
    {
        const int t = threadIdx.x;
        // pos3 rot9 rot_ext9 pos_ext3 ...  (L2 loads: inside k_iekf_loop the state was written by another SM during this launch)
        auto x = [&](int q) { return __ldcg(xsrc + q); };
        if (t < 9) {
            const int i = t / 3, j = t % 3;
            double sacc = x(3 + 3 * i) * x(12 + j);
            sacc += x(3 + 3 * i + 1) * x(15 + j);
            sacc += x(3 + 3 * i + 2) * x(18 + j);
            ms.r_wl.a[t] = sacc;
        } else if (t < 12) {
            const int i = t - 9;
            double sacc = x(3 + 3 * i) * x(21);
            sacc += x(3 + 3 * i + 1) * x(22);
            sacc += x(3 + 3 * i + 2) * x(23);
            ms.p_wl[i] = sacc + x(i);
        } else if (t < 21) ms.R.a[t - 12] = x(3 + (t - 12));
        else if (t < 30) ms.Rext.a[t - 21] = x(12 + (t - 21));
        else if (t < 33) ms.pext[t - 30] = x(21 + (t - 30));
        else if (t < 42) { const int e = t - 33; ms.Prr.a[e] = __ldcg(Psrc + (3 + e / 3) * 23 + 3 + e % 3); }
        else if (t < 51) { const int e = t - 42; ms.Ppp.a[e] = __ldcg(Psrc + (e / 3) * 23 + e % 3); }
    }
    const int n = FIRST ? in->n : ctl->n;
    const float* pts = FIRST ? in->pts : nullptr;
    const int stride = FIRST ? (in->stride == 4 ? 4 : 3) : 3;
    const bool copy_raw = FIRST && pts != s.raw;
    int gate_next = 0x7fffffff;                    // streamed upload: first point this warp does not know to have arrived yet
    if (FIRST) { const int g = in->gate_pts; if (g > 0) gate_next = g; }
    __syncthreads();
    const size_t NM = (size_t)s.nmax;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double* const stg = sh.stage[wid];

    // H^T R^-1 H / H^T R^-1 z are accumulated by the warp, not by the thread: every lane stages (J, w J, r) of its valid
    // point in shared memory, then lane v sums ITS entry (a, c) of the upper triangle (or entry a of b) over the staged
    // points in index order.  One accumulator per lane instead of NV per thread (the kernel was register-bound at one CTA
    // per SM) and no shuffle tree at the end; the order of the sum is fixed: points in index order inside a warp's batch,
    // batches in order, warps in order, CTAs in the solver's fixed order.
    int offA[VPL], offB[VPL];
#pragma unroll
    for (int q = 0; q < VPL; q++) {
        const int v = lane + 32 * q;
        int a = 0, c = 0, isb = 0;
        if (v < NH) {
            int rem = v;
            while (rem >= D - a) { rem -= D - a; a++; }
            c = a + rem;
        } else if (v < NA) { a = v - NH; isb = 1; }
        offA[q] = D + a;                            // w J[a]
        offB[q] = isb ? 2 * D : c;                  // J[c] or the residual
    }
    double acc[VPL];
#pragma unroll
    for (int q = 0; q < VPL; q++) acc[q] = 0.0;
    int cnt = 0;

    for (int base = (pb * NW + wid) * 32; base < n; base += npb * NW * 32) {
        const int i = base + lane;
        bool valid = false;
        if (FIRST) {
            const int last = min(base + 31, n - 1);
            if (last >= gate_next) {
                const int g = in->gate_pts;
                gate_next = ((int)wait_upload(ctl, in->seq * 64ull, (unsigned long long)(last / g)) + 1) * g;
            }
        }
        if (i < n) {
            V3 pl;
            M3 cl;
            // later iterations of a scan: the voxel found for this point in the previous iteration (the map does not change between the
            // iterations, and the state moves by far less than a voxel: the key is almost always the same and the hash probe is skipped)
            unsigned long long pk_prev = KEY_EMPTY;
            int slot_prev = -1;
            if (!FIRST && reuse_slots) { pk_prev = s.rkey[i]; slot_prev = s.rslot[i]; }
            if (FIRST) {
                const float* pp = pts + (size_t)i * stride;
                const float fx = __ldcg(pp), fy = __ldcg(pp + 1), fz = __ldcg(pp + 2);      // L2: with a streamed upload the DMA engine wrote them during this launch
                if (copy_raw) { s.raw[3 * (size_t)i] = fx; s.raw[3 * (size_t)i + 1] = fy; s.raw[3 * (size_t)i + 2] = fz; }
                pl = v3((double)fx, (double)fy, (double)fz);
                calc_body_cov(pl, s.range_var, s.sn2, cl);            // (edits pl.z == 0 -> 0.001, Q16)
#pragma unroll
                for (int k = 0; k < 3; k++) s.pl[(size_t)k * NM + i] = pl[k];
#pragma unroll
                for (int k = 0; k < 9; k++) s.cl[(size_t)k * NM + i] = cl.a[k];
            } else {
                pl = v3(s.pl[i], s.pl[NM + i], s.pl[2 * NM + i]);
            }
            const V3 pw = add(mul(ms.r_wl, pl), ms.p_wl);
            unsigned long long pk;
            int slot = -1;
            if (voxel_index(pw[0], pw[1], pw[2], m.voxel_size, pk, m.voxel_inv)) slot = (pk == pk_prev) ? slot_prev : hash_find(m, pk);
            if (FIRST || slot != slot_prev) s.rslot[i] = slot;
            if (!FIRST) {                                  // fetched up front: one memory latency less on the chain
#pragma unroll
                for (int k = 0; k < 9; k++) cl.a[k] = s.cl[(size_t)k * NM + i];
            }
            V3 nrm;
            double res;
            uint8_t status = 0;
            if (slot >= 0) {
                status = 1;
                const double* h = m.hot + (size_t)slot * 8;
                uint32_t flags; int nn;
                hot_get_fn(m.hot, slot, flags, nn);
                if (flags & F_PLANE) {
                    status |= 2;
                    const V3 mean = v3(h[0], h[1], h[2]);
                    nrm = v3(h[3], h[4], h[5]);
                    const V3 p2m = sub(pw, mean);
                    res = dot(nrm, p2m);
                    const M3 cw = world_cov(ms.r_wl, cl, pl, ms.Prr, ms.Ppp);
                    // sigma_l = J_nq plane_cov J_nq^T (plane_cov is never assigned -> 0, Q1) + n^T C_w n
                    const double sigma = mul(mul(tr(nrm), cw), nrm)[0];
                    valid = fabs(res) < 3.0 * sqrt(sigma);
#pragma unroll
                    for (int k = 0; k < 3; k++) { s.rnorm[(size_t)k * NM + i] = nrm[k]; s.rmean[(size_t)k * NM + i] = mean[k]; }
                    s.rres[i] = res;
                }
                s.rvalid[i] = valid ? 1 : 0;
            } else {
                // Q2: voxel not in the map -> the record keeps whatever an earlier pass left in it
                valid = s.rvalid[i] != 0;
                if (valid) {
                    nrm = v3(s.rnorm[i], s.rnorm[NM + i], s.rnorm[2 * NM + i]);
                    res = s.rres[i];
                }
            }
            if (valid) status |= 4;
            s.rstatus[i] = status;
            s.rkey[i] = pk;
            if (valid) {
                // lio_builder.cpp:294-297
                const Mat<1, 3> nt = tr(nrm);
                const double r_cov = mul(mul(mul(mul(nt, ms.r_wl), cl), tr(ms.r_wl)), nrm)[0];
                const double r_info = r_cov < 0.0002 ? 5000 : 1.0 / r_cov;
                double J[D];
                J[0] = nrm[0]; J[1] = nrm[1]; J[2] = nrm[2];
                const Mat<1, 3> jr = mul(mul(neg(nt), ms.R), hat(add(mul(ms.Rext, pl), ms.pext)));
                J[3] = jr[0]; J[4] = jr[1]; J[5] = jr[2];
                if (EXT) {
                    const Mat<1, 3> je = mul(mul(neg(nt), ms.r_wl), hat(pl));
                    const Mat<1, 3> jp = mul(nt, ms.R);
                    J[6] = je[0]; J[7] = je[1]; J[8] = je[2];
                    J[D - 3] = jp[0]; J[D - 2] = jp[1]; J[D - 1] = jp[2];
                }
                double* sp = stg + lane * SP;
#pragma unroll
                for (int a = 0; a < D; a++) { sp[a] = J[a]; sp[D + a] = J[a] * r_info; }
                sp[2 * D] = res;
            }
        }
        unsigned mask = __ballot_sync(0xffffffffu, valid);
        cnt += __popc(mask);
        __syncwarp();
        for (; mask; mask &= mask - 1) {
            const double* sp = stg + (__ffs(mask) - 1) * SP;
#pragma unroll
            for (int q = 0; q < VPL; q++) acc[q] += sp[offA[q]] * sp[offB[q]];
        }
        __syncwarp();
    }

    // per-warp sums -> shared -> one partial vector per block; fixed order everywhere
#pragma unroll
    for (int q = 0; q < VPL; q++) { const int v = lane + 32 * q; if (v < NA) sh.red[wid][v] = acc[q]; }
    if (lane == 0) sh.red[wid][NA] = (double)cnt;
    __syncthreads();
    for (int v = threadIdx.x; v < NV; v += blockDim.x) {
        double t = sh.red[0][v];
        for (int w = 1; w < NW; w++) t += sh.red[w][v];
        partials[(size_t)pb * PARTIAL_STRIDE + v] = t;
    }
}

// Two resident CTAs per SM (four of the 128-thread estimate_ext variant): 128 registers, no spills.  Three (80 registers,
// 196 B of spills) measured the same at 200 k points and slower at 20 k.
// FIRST: the first iteration of a scan.  It takes the scan straight from the header `in` (round 1 ran a staging kernel,
// k_set_scan, in front): every thread evaluates calcBodyCov (commons.cpp:18-45, lio_builder.cpp:224-229) for its points, keeps
// point_lidar / cov_lidar for the later iterations and the map update, and reads the prior from where the header says it is.
// One launch = one iteration (the S1 seam vmp_measure with solve = 0, and the scan graph when VMP_IEKF_LOOP=0).
template <bool EXT, bool FIRST>
__global__ void __launch_bounds__(EXT ? 128 : 256, EXT ? 4 : 2)
k_measure(DevMap m, DevScan s, DevFilter* f, DevCtl* ctl, double* partials, int solve, const ScanIn* __restrict__ in, int reuse_slots) {
    // a CTA is either the solver or a measurement CTA: one overlay for both (static shared memory is limited to 48 KB)
    union Overlay { SolveShared<EXT> sol; MeasShared<EXT> meas; };
    __shared__ __align__(16) unsigned char sh_raw[sizeof(Overlay)];
    if (!FIRST && ctl->done) return;
    // launched with solve = 1 the grid has one more CTA: block 0 runs this iteration's 23-dof solve (IESKF::update body,
    // vmp_solve.cuh), starting with the part that needs no measurement while the other CTAs measure
    if (solve && blockIdx.x == 0) {
        ieskf_solve_cta<EXT, EXT ? 128 : 256>(*reinterpret_cast<SolveShared<EXT>*>(sh_raw), f, ctl, partials, (int)gridDim.x - 1, FIRST ? in : nullptr, 0ull);
        return;
    }
    measure_pass<EXT, FIRST>(*reinterpret_cast<MeasShared<EXT>*>(sh_raw), m, s, f, ctl, partials, (int)blockIdx.x - solve, (int)gridDim.x - solve, in, reuse_slots);
    if (!solve) return;
    // release this CTA's partial sums to the solver CTA
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(&ctl->ticket, 1u);
}

// The whole iteration loop of IESKF::update (ieskf.cpp:125-156) as ONE launch: the measurement CTAs and the solver CTA stay resident,
// the solver publishes every new state through ctl->iter_pub (release) and the measurement CTAs pick it up (acquire) - no kernel
// boundary, no launch ramp and no empty launches for the iterations a converged scan does not run.  All CTAs must be co-resident
// (grid = 2 per SM, checked at create with the occupancy API; vmp_capi.cu falls back to one launch per iteration otherwise).
template <bool EXT>
__global__ void __launch_bounds__(EXT ? 128 : 256, EXT ? 4 : 2)
k_iekf_loop(DevMap m, DevScan s, DevFilter* f, DevCtl* ctl, double* partials, const ScanIn* __restrict__ in, int reuse_slots) {
    union Overlay { SolveShared<EXT> sol; MeasShared<EXT> meas; };
    __shared__ __align__(16) unsigned char sh_raw[sizeof(Overlay)];
    __shared__ int s_stop;
    const unsigned long long base = in->seq * 8ull;             // iter_pub values of this scan: base + executed iterations
    if (blockIdx.x == 0) {
        SolveShared<EXT>& S = *reinterpret_cast<SolveShared<EXT>*>(sh_raw);
        for (int it = 0; it < 8; it++) {
            ieskf_solve_cta<EXT, EXT ? 128 : 256>(S, f, ctl, partials, (int)gridDim.x - 1, it == 0 ? in : nullptr, base + (unsigned long long)it + 1ull);
            if (S.s_last) break;
            __syncthreads();
        }
        return;
    }
    MeasShared<EXT>& sh = *reinterpret_cast<MeasShared<EXT>*>(sh_raw);
    const int pb = (int)blockIdx.x - 1, npb = (int)gridDim.x - 1;
    for (int it = 0; it < 8; it++) {
        if (it == 0) measure_pass<EXT, true>(sh, m, s, f, ctl, partials, pb, npb, in, 0);
        else {
            if (threadIdx.x == 0) {       // iter_pub = 2 * (base + executed iterations) + stop flag: one look-up tells both
                unsigned long long v;
                while (((v = ld_acquire_u64(&ctl->iter_pub)) >> 1) < base + (unsigned long long)it) __nanosleep(20);
                s_stop = (int)(v & 1ull);
            }
            __syncthreads();
            if (s_stop) break;
            measure_pass<EXT, false>(sh, m, s, f, ctl, partials, pb, npb, nullptr, reuse_slots);
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(&ctl->ticket, 1u);
    }
}

void launch_measure(cudaStream_t st, bool ext, int grid, const DevMap& m, const DevScan& s, DevFilter* f, DevCtl* ctl, double* partials, int solve, const ScanIn* first,
                    bool reuse_slots) {
    const int g = grid + (solve ? 1 : 0);
    if (first) {
        if (ext) k_measure<true, true><<<g, 128, 0, st>>>(m, s, f, ctl, partials, solve, first, 0);
        else k_measure<false, true><<<g, 256, 0, st>>>(m, s, f, ctl, partials, solve, first, 0);
    } else {
        if (ext) k_measure<true, false><<<g, 128, 0, st>>>(m, s, f, ctl, partials, solve, nullptr, reuse_slots ? 1 : 0);
        else k_measure<false, false><<<g, 256, 0, st>>>(m, s, f, ctl, partials, solve, nullptr, reuse_slots ? 1 : 0);
    }
}
// Cooperative launch: the CTAs of this kernel wait for each other, so the grid has to be resident as a whole.  On an otherwise idle GPU a
// plain launch of a grid that fits would do; the cooperative attribute makes the driver gang-schedule it, which is what keeps two
// handles (two trajectories on one GPU, driven from two host threads) from starving each other's half-resident grids.
cudaError_t launch_iekf_loop(cudaStream_t st, bool ext, int grid, const DevMap& m, const DevScan& s, DevFilter* f, DevCtl* ctl, double* partials, const ScanIn* in, bool reuse_slots) {
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(grid + 1); lc.blockDim = dim3(ext ? 128 : 256); lc.dynamicSmemBytes = 0; lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    lc.attrs = at; lc.numAttrs = 1;
    const int reuse = reuse_slots ? 1 : 0;
    return ext ? cudaLaunchKernelEx(&lc, k_iekf_loop<true>, m, s, f, ctl, partials, in, reuse)
               : cudaLaunchKernelEx(&lc, k_iekf_loop<false>, m, s, f, ctl, partials, in, reuse);
}
// can grid + 1 CTAs of the loop kernel be resident at once?  (they wait for each other: anything less would never finish)
bool iekf_loop_fits(bool ext, int grid, int sm_count) {
    int per_sm = 0;
    const cudaError_t e = ext ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_iekf_loop<true>, 128, 0)
                              : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_iekf_loop<false>, 256, 0);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return (long long)per_sm * sm_count >= (long long)grid + 1;
}
// Motion compensation of a raw scan (LIOBuilder::undistortCloud, lio_builder.cpp:127-152) on the device: one thread per
// point.  The scan is time-sorted; a point at offset t belongs to the last IMU pose `head` with head.offset < t (points at
// t <= offset of the first pose stay as they are), is carried with head's pose and the next pose's body rates to t and
// expressed in the lidar frame at the end of the scan (= the propagated prior).  The reference's loop does not stop at
// the first point: it compensates point 0 once more for every earlier pose interval; reproduced as written.
// fp64 in the same operation order as the host restatement, result rounded to float like the in-place PCL point.
__global__ void __launch_bounds__(256) k_undistort(const ScanIn* __restrict__ in, const DevPose* __restrict__ poses, float4* cloud, float4* host_copy) {
    __shared__ DevPose sp[MAX_POSES];
    __shared__ double sx[24];
    const int n = in->n, K = in->n_poses;
    for (int q = threadIdx.x; q < K * (int)(sizeof(DevPose) / 8); q += blockDim.x) reinterpret_cast<double*>(sp)[q] = reinterpret_cast<const double*>(poses)[q];
    if (threadIdx.x < 24) sx[threadIdx.x] = in->x[threadIdx.x];          // pos3 rot9 rot_ext9 pos_ext3 of the propagated prior
    __syncthreads();
    M3 cur_rot, cur_ext;
    for (int k = 0; k < 9; k++) { cur_rot.a[k] = sx[3 + k]; cur_ext.a[k] = sx[12 + k]; }
    const V3 cur_pos = v3(sx[0], sx[1], sx[2]), cur_pext = v3(sx[21], sx[22], sx[23]);
    // Point 0 is carried through EVERY pose interval before its own (the quirk above): up to n_poses dependent steps, each with an
    // Exp() - one thread doing that was the length of the whole kernel (20 us for a 20 000-point scan).  The pose part of a step
    // (point_rot, point_pos) depends on the point's time only, so the lanes of one warp evaluate the steps' pose parts side by side
    // and lane 0 then applies them in the reference's order: the same operations on the same operands.
    __shared__ double s_pr[MAX_POSES][9], s_pp[MAX_POSES][3];
    if (blockIdx.x == 0 && threadIdx.x < 32 && n > 0) {
        const int lane = threadIdx.x;
        float4 p = cloud[0];
        const double t = (double)p.w / double(1000);
        int h0 = -1;
        for (int k = 0; k + 1 < K; k++) if (sp[k].offset < t) h0 = k;
        for (int h = lane; h <= h0; h += 32) {
            const DevPose& head = sp[h];
            const DevPose& tail = sp[h + 1];
            const double dt = t - head.offset;
            M3 hr;
            for (int k = 0; k < 9; k++) hr.a[k] = head.rot[k];
            const M3 point_rot = mul(hr, so3_exp(scale(v3(tail.gyro[0], tail.gyro[1], tail.gyro[2]), dt)));
            const V3 point_pos = add(add(v3(head.pos[0], head.pos[1], head.pos[2]), scale(v3(head.vel[0], head.vel[1], head.vel[2]), dt)),
                                     scale(scale(scale(v3(tail.acc[0], tail.acc[1], tail.acc[2]), 0.5), dt), dt));
            for (int k = 0; k < 9; k++) s_pr[h][k] = point_rot.a[k];
            for (int k = 0; k < 3; k++) s_pp[h][k] = point_pos[k];
        }
        __syncwarp();
        if (lane == 0) {
            for (int h = h0; h >= 0; h--) {
                const V3 point = v3((double)p.x, (double)p.y, (double)p.z);
                M3 point_rot;
                for (int k = 0; k < 9; k++) point_rot.a[k] = s_pr[h][k];
                const V3 point_pos = v3(s_pp[h][0], s_pp[h][1], s_pp[h][2]);
                const V3 inner = sub(add(mul(point_rot, add(mul(cur_ext, point), cur_pext)), point_pos), cur_pos);
                const V3 pc = mul(tr(cur_ext), sub(mul(tr(cur_rot), inner), cur_pext));
                p.x = (float)pc[0]; p.y = (float)pc[1]; p.z = (float)pc[2];
            }
            cloud[0] = p;
            if (host_copy) host_copy[0] = p;
        }
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (i == 0) continue;                                         // (done above)
        float4 p = cloud[i];
        const double t = (double)p.w / double(1000);
        int h = -1;
        for (int k = 0; k + 1 < K; k++) if (sp[k].offset < t) h = k;
        for (; h >= 0; h--) {
            const DevPose& head = sp[h];
            const DevPose& tail = sp[h + 1];
            const double dt = t - head.offset;
            const V3 point = v3((double)p.x, (double)p.y, (double)p.z);
            M3 hr;
            for (int k = 0; k < 9; k++) hr.a[k] = head.rot[k];
            const M3 point_rot = mul(hr, so3_exp(scale(v3(tail.gyro[0], tail.gyro[1], tail.gyro[2]), dt)));
            const V3 point_pos = add(add(v3(head.pos[0], head.pos[1], head.pos[2]), scale(v3(head.vel[0], head.vel[1], head.vel[2]), dt)),
                                     scale(scale(scale(v3(tail.acc[0], tail.acc[1], tail.acc[2]), 0.5), dt), dt));
            const V3 inner = sub(add(mul(point_rot, add(mul(cur_ext, point), cur_pext)), point_pos), cur_pos);
            const V3 pc = mul(tr(cur_ext), sub(mul(tr(cur_rot), inner), cur_pext));
            p.x = (float)pc[0]; p.y = (float)pc[1]; p.z = (float)pc[2];
            break;                                                    // only the first point is revisited (above)
        }
        cloud[i] = p;
        if (host_copy) host_copy[i] = p;                              // mapped host memory: the caller's cloud is edited in place
    }
}
// ---------------------------------------------------------------------------- IMU propagation on the device
// IESKF::predict (ieskf.cpp:101-123) for every IMU step of the scan, and the IMU pose list of LIOBuilder::undistortCloud
// (lio_builder.cpp:78-116), by ONE CTA at the head of the raw scan's graph: the filter state and covariance stay resident on the
// device from scan to scan.  F is the identity plus seven small blocks and G has four, so P = F P F^T + G Q G^T is evaluated on
// the structural non-zeros only, in ascending column order - the same terms in the same order as the host / oracle evaluation
// (x + 0 * y == x).  Writes the propagated prior into the scan header (where the motion compensation and the first IEKF
// iteration read it) and the poses into the staging area.
struct PredictRows { signed char n[23]; signed char col[23][9]; };
__constant__ PredictRows c_frows = {
    {2, 2, 2, 6, 6, 6, 1, 1, 1, 1, 1, 1, 9, 9, 9, 1, 1, 1, 1, 1, 1, 2, 2},
    {{0, 12}, {1, 13}, {2, 14}, {3, 4, 5, 15, 16, 17}, {3, 4, 5, 15, 16, 17}, {3, 4, 5, 15, 16, 17}, {6}, {7}, {8}, {9}, {10}, {11},
     {3, 4, 5, 12, 18, 19, 20, 21, 22}, {3, 4, 5, 13, 18, 19, 20, 21, 22}, {3, 4, 5, 14, 18, 19, 20, 21, 22},
     {15}, {16}, {17}, {18}, {19}, {20}, {21, 22}, {21, 22}}};
__constant__ PredictRows c_grows = {
    {0, 0, 0, 3, 3, 3, 0, 0, 0, 0, 0, 0, 3, 3, 3, 1, 1, 1, 1, 1, 1, 0, 0},
    {{0}, {0}, {0}, {0, 1, 2}, {0, 1, 2}, {0, 1, 2}, {0}, {0}, {0}, {0}, {0}, {0}, {3, 4, 5}, {3, 4, 5}, {3, 4, 5}, {6}, {7}, {8}, {9}, {10}, {11}, {0}, {0}}};

__global__ void __launch_bounds__(256) k_predict(DevFilter* f, ScanIn* in, const DevPredictIn* __restrict__ pin, DevPose* poses) {
    __shared__ double sx[36], sxn[36], sP[529], sF[529], sT1[529], sT2[529], sG[276], sT3[276], sQ[144], slast[6];
    const int tid = threadIdx.x;
    for (int q = tid; q < 36; q += 256) sx[q] = f->x[q];
    for (int q = tid; q < 529; q += 256) sP[q] = f->P[q];
    for (int q = tid; q < 144; q += 256) sQ[q] = pin->Q[q];
    if (tid < 6) slast[tid] = pin->use_last ? pin->last[tid] : f->last[tid];
    const int ns = pin->n_steps;
    __syncthreads();
    int npose = 0;
    if (tid == 0) {         // Pose{0.0, last_acc, last_gyro, vel, pos, rot} (lio_builder.cpp:78-79)
        DevPose& p = poses[0];
        p.offset = 0.0;
        for (int k = 0; k < 3; k++) { p.acc[k] = slast[k]; p.gyro[k] = slast[3 + k]; p.vel[k] = sx[24 + k]; p.pos[k] = sx[k]; }
        for (int k = 0; k < 9; k++) p.rot[k] = sx[3 + k];
    }
    npose = 1;
    for (int s = 0; s < ns; s++) {
        const DevImuStep st = pin->steps[s];
        const double dt = st.dt;
        const St x = st_load(sx);
        const V3 w = sub(v3(st.gyro[0], st.gyro[1], st.gyro[2]), x.bg);
        const V3 a = sub(v3(st.acc[0], st.acc[1], st.acc[2]), x.ba);
        // F = I + blocks, G (ieskf.cpp:104-117); five independent pieces on five warps
        for (int q = tid; q < 529; q += 256) sF[q] = (q / 23 == q % 23) ? 1.0 : 0.0;
        for (int q = tid; q < 276; q += 256) sG[q] = 0.0;
        __syncthreads();
        if (tid == 0) {
            const M3 e = so3_exp(scale(neg(w), dt));
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) sF[(3 + i) * 23 + 3 + j] = e(i, j);
            for (int i = 0; i < 3; i++) { sF[i * 23 + 12 + i] = dt; sG[(15 + i) * 12 + 6 + i] = dt; sG[(18 + i) * 12 + 9 + i] = dt; }
        } else if (tid == 32) {
            const M3 j = scale(neg(right_jacobian(scale(w, dt))), dt);
            for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) { sF[(3 + i) * 23 + 15 + k] = j(i, k); sG[(3 + i) * 12 + k] = j(i, k); }
        } else if (tid == 64) {
            const M3 ra = scale(mul(neg(x.rot), hat(a)), dt);
            const M3 nr = scale(neg(x.rot), dt);
            for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) { sF[(12 + i) * 23 + 3 + k] = ra(i, k); sF[(12 + i) * 23 + 18 + k] = nr(i, k); sG[(12 + i) * 12 + 3 + k] = nr(i, k); }
        } else if (tid == 96) {
            const Mat<3, 2> mx = st_Mx(x.g);
            const Mat<3, 2> mxd = scale(mx, dt);
            const Mat<2, 2> nm = mul(st_Nx(x.g), mx);
            for (int i = 0; i < 3; i++) for (int k = 0; k < 2; k++) sF[(12 + i) * 23 + 21 + k] = mxd(i, k);
            for (int i = 0; i < 2; i++) for (int k = 0; k < 2; k++) sF[(21 + i) * 23 + 21 + k] = nm(i, k);
        } else if (tid == 128) {
            // x_ += delta (Vector24d variant, ieskf.cpp:23-33): only pos, rot, vel change (the other blocks add exact zeros)
            St xn = x;
            xn.pos = add(x.pos, scale(x.vel, dt));
            xn.rot = mul(x.rot, so3_exp(scale(w, dt)));
            xn.vel = add(x.vel, scale(add(mul(x.rot, a), x.g), dt));
            st_store(xn, sxn);
        }
        __syncthreads();
        for (int q = tid; q < 529; q += 256) {              // T1 = F P
            const int i = q / 23, j = q % 23;
            double acc = sF[i * 23 + c_frows.col[i][0]] * sP[c_frows.col[i][0] * 23 + j];
            for (int e = 1; e < c_frows.n[i]; e++) { const int c = c_frows.col[i][e]; acc += sF[i * 23 + c] * sP[c * 23 + j]; }
            sT1[q] = acc;
        }
        for (int q = tid; q < 276; q += 256) {              // T3 = G Q
            const int i = q / 12, c = q % 12;
            double acc = 0.0;
            if (c_grows.n[i] > 0) {
                acc = sG[i * 12 + c_grows.col[i][0]] * sQ[c_grows.col[i][0] * 12 + c];
                for (int e = 1; e < c_grows.n[i]; e++) { const int k = c_grows.col[i][e]; acc += sG[i * 12 + k] * sQ[k * 12 + c]; }
            }
            sT3[q] = acc;
        }
        __syncthreads();
        for (int q = tid; q < 529; q += 256) {              // T2 = T1 F^T ; P = T2 + T3 G^T
            const int i = q / 23, j = q % 23;
            double acc = sT1[i * 23 + c_frows.col[j][0]] * sF[j * 23 + c_frows.col[j][0]];
            for (int e = 1; e < c_frows.n[j]; e++) { const int c = c_frows.col[j][e]; acc += sT1[i * 23 + c] * sF[j * 23 + c]; }
            double g = 0.0;
            if (c_grows.n[i] > 0 && c_grows.n[j] > 0) {
                g = sT3[i * 12 + c_grows.col[j][0]] * sG[j * 12 + c_grows.col[j][0]];
                for (int e = 1; e < c_grows.n[j]; e++) { const int c = c_grows.col[j][e]; g += sT3[i * 12 + c] * sG[j * 12 + c]; }
            }
            sT2[q] = acc + g;
        }
        __syncthreads();
        for (int q = tid; q < 529; q += 256) sP[q] = sT2[q];
        if (tid < 36) sx[tid] = sxn[tid];
        __syncthreads();
        if (st.offset > -1.0e299) {     // (the closing step carries -1e300) last_gyro / last_acc with the propagated state, pose at the end of the step (lio_builder.cpp:107-111)
            if (tid == 0) {
                const St xu = st_load(sx);
                const V3 lg = sub(v3(st.gyro[0], st.gyro[1], st.gyro[2]), xu.bg);
                const V3 la = add(mul(xu.rot, sub(v3(st.acc[0], st.acc[1], st.acc[2]), xu.ba)), xu.g);
                for (int k = 0; k < 3; k++) { slast[k] = la[k]; slast[3 + k] = lg[k]; }
                if (npose < MAX_POSES) {
                    DevPose& p = poses[npose];
                    p.offset = st.offset;
                    for (int k = 0; k < 3; k++) { p.acc[k] = la[k]; p.gyro[k] = lg[k]; p.vel[k] = xu.vel[k]; p.pos[k] = xu.pos[k]; }
                    for (int k = 0; k < 9; k++) p.rot[k] = xu.rot.a[k];
                }
            }
            npose++;
        }
        __syncthreads();
    }
    // the propagated prior -> scan header (k_undistort, first IEKF iteration) and filter; last_acc / last_gyro for the next scan
    for (int q = tid; q < 36; q += 256) { in->x[q] = sx[q]; f->x[q] = sx[q]; }
    for (int q = tid; q < 529; q += 256) { in->P[q] = sP[q]; f->P[q] = sP[q]; }
    if (tid < 6) f->last[tid] = slast[tid];
    if (tid == 0) in->n_poses = npose < MAX_POSES ? npose : MAX_POSES;
}
void launch_predict(cudaStream_t st, DevFilter* f, ScanIn* in, const DevPredictIn* pin, DevPose* poses) { k_predict<<<1, 256, 0, st>>>(f, in, pin, poses); }

__global__ void __launch_bounds__(256) k_cloud_out(const ScanIn* __restrict__ in, const float4* __restrict__ cloud, float4* host_copy) {
    const int n = in->n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) host_copy[i] = cloud[i];
}
void launch_cloud_out(cudaStream_t st, int grid, const ScanIn* in, const float4* cloud, float4* host_copy) {
    k_cloud_out<<<grid, 256, 0, st>>>(in, cloud, host_copy);
}
void launch_undistort(cudaStream_t st, int grid, const ScanIn* in, const DevPose* poses, float4* cloud, float4* host_copy) {
    k_undistort<<<grid, 256, 0, st>>>(in, poses, cloud, host_copy);
}

// The posterior goes to the host mailbox from a side branch of the graph (posted PCIe writes + a system-scope fence cost
// ~2 us that the map update need not wait for); seq is written last.
__global__ void __launch_bounds__(256) k_state_out(const DevFilter* __restrict__ f, const DevCtl* __restrict__ ctl, StateOut* out) {
    const int tid = threadIdx.x;
    for (int q = tid; q < 529; q += 256) out->P[q] = f->P[q];
    if (tid < 36) out->x[tid] = f->x[tid];
    if (tid < 8) out->effect[tid] = ctl->effect[tid];
    __syncthreads();
    if (tid == 0) {
        out->iter = ctl->iter;
        out->converged = ctl->converged;
        __threadfence_system();
        *(volatile unsigned long long*)&out->seq = ctl->seq;
    }
}
void launch_state_out(cudaStream_t st, const DevFilter* f, const DevCtl* ctl, StateOut* out) { k_state_out<<<1, 256, 0, st>>>(f, ctl, out); }

void launch_set_scan(cudaStream_t st, int grid, const DevScan& s, const ScanIn* in, DevFilter* f, DevCtl* ctl) { k_set_scan<<<grid, 256, 0, st>>>(s, in, f, ctl); }

}  // namespace vmp
