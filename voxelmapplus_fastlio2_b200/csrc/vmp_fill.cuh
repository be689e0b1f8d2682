// vmp_fill.cuh — VoxelGrid::pushPoint / addToPlane / updatePlane (voxel_map.cpp:29-136) and the per-voxel
// part of VoxelMap::build (voxel_map.cpp:200-230).  Included by vmp_map.cu inside namespace vmp.
//
// ONE kernel, k_fill: a warp takes a touched voxel (dynamic hand-out: voxels differ by three orders of magnitude
// in work) and runs its whole fill phase of the scan with everything it needs staged in SHARED MEMORY:
//   1. select the (at most max_point_thresh) points the voxel consumes, in point order (histogram select over the
//      voxel's unordered segment, two passes; bisection on the index range for scans beyond 512 k points);
//   2. gather them (xyz + covariance) behind the voxel's stored points, all of them component-major in shared memory;
//   3. the pushPoint state machine: its CONTROL FLOW never depends on the result of a refit (is_plane only selects
//      between two identical branches while the voxel is filling), so which steps refit, which step closes the voxel
//      and the final counters are evaluated in closed form up front; the per-point loop holds the two recurrences that
//      ARE sequential (running mean, sum p p^T; nine lanes) and takes a snapshot (n, mean, sum pp^T, #stored points)
//      wherever the reference calls updatePlane() past its early return;
//   4. the refits of those snapshots: the 3x3 symmetric eigen-solves (restated SelfAdjointEigenSolver) of up to 32
//      snapshots run ONE PER LANE, concurrently (round 1 solved every job redundantly on all 32 lanes of every
//      32-point batch); then, snapshot by snapshot in refit order, the 6x6 contributions J Sigma J^T of the stored
//      points are computed one point per lane into a shared tile and added to plane->cov entry-wise in stored-point
//      order (Q7: accumulates across refits, never reset) - lane e owns entry e, so the order of the adds is exactly
//      the reference's;
//   5. write-back: running mean, sum pp^T, normal / centre / is_plane of the last refit, covariance, counters, and the
//      consumed points appended to the voxel's stored points (skipped when the voxel closes: temp_points is freed).
// Nothing is staged in HBM between those steps (round 1: refit jobs + 288 B per stored point per refit written by
// k_fill_refit and read back by k_fill_acc, ~190 MB per 200 k-point scan against ~19 MB of compulsory traffic).
// k_fill_build is the build() variant (no cap, one updatePlane() at the end, Q18); it runs once per map.
#pragma once

constexpr int FILL_WARPS = 4;           // warps per CTA of k_fill, warp path (a voxel per warp)
constexpr int HEAVY_WARPS = 4;          // warps per CTA of the CTA path (a voxel per CTA): chunks of a refit group evaluated side by side
constexpr int TILE_LD = 37;             // doubles per point in the contribution tile (odd: conflict-free both ways)
constexpr int SNAP_W = 12;              // doubles per snapshot: n, nt, mean[3], ppt[6], pad
constexpr int SEL_BINS = 256;
constexpr int CAND_CAP = 32 * TILE_LD * 2 - SEL_BINS;      // ints that fit in the tile behind the histogram (selection scratch overlays the tile)

// ascending in-place sort of a[0..c) by one warp (distinct values); a may be shared or global
__device__ __noinline__ void warp_sort(int* a, int c) {
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
    if (c <= 128) {
        // rank sort: every lane keeps up to four values and counts the smaller ones while the array is read once, broadcast
        // (c loads and 4 c compares, no exchange stages: the bitonic network below took ~16 k cycles for the hundred indices of
        // a voxel that fills in one scan - 28 stages of shared-memory exchanges with a barrier each)
        int v[4], r[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const int q = lane + 32 * u; v[u] = q < c ? a[q] : INT_MAX; r[u] = 0; }
        __syncwarp();
#pragma unroll 2
        for (int l = 0; l < c; l++) {
            const int o = a[l];
#pragma unroll
            for (int u = 0; u < 4; u++) r[u] += (o < v[u]) ? 1 : 0;
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 4; u++) if (lane + 32 * u < c) a[r[u]] = v[u];
        __syncwarp();
        return;
    }
    int npow = 256, lnp = 8;
    while (npow < c) { npow <<= 1; lnp++; }
    const int half = npow >> 1;
    for (int lk = 1; lk <= lnp; lk++) {                         // k = 2^lk (shifts: the divisions by k and j were most of a stage)
        const int k = 1 << lk, hk = k >> 1;
        for (int t = lane; t < half; t += 32) {                 // flip stage: all comparators ascending
            const int blk = t >> (lk - 1), o = t & (hk - 1);
            const int lo = blk * k + o, hi = blk * k + k - 1 - o;
            if (hi < c) { const int x = a[lo], y = a[hi]; if (x > y) { a[lo] = y; a[hi] = x; } }
        }
        __syncwarp();
        for (int lj = lk - 2; lj >= 0; lj--) {
            const int j = 1 << lj;
            for (int t = lane; t < half; t += 32) {
                const int lo = ((t >> lj) << (lj + 1)) + (t & (j - 1)), hi = lo + j;
                if (hi < c) { const int x = a[lo], y = a[hi]; if (x > y) { a[lo] = y; a[hi] = x; } }
            }
            __syncwarp();
        }
    }
}

// the K smallest of seg[0..c) (distinct point indices in [0, n)), ascending, into sel[0..K) (shared).
// hist: SEL_BINS ints, cand: CAND_CAP ints of shared scratch.
__device__ void warp_select_sorted(const int* seg, int c, int K, int n, int* sel, int* hist, int* cand) {
    const int lane = threadIdx.x & 31;
    if (K >= c || c <= 32) {                                     // small segment: sort all of it, the caller takes the first K
        for (int q = lane; q < c; q += 32) sel[q] = seg[q];
        __syncwarp();
        warp_sort(sel, c);
        return;
    }
    int S = 0;
    while (((n - 1) >> S) >= SEL_BINS) S++;
    if ((1 << S) <= CAND_CAP) {
        // histogram select: the segment is unordered (its positions were handed out by atomics), the K smallest indices are
        // those of the leading histogram bins plus the smallest of one boundary bin.  Two passes over the segment (the
        // bisection below takes log2(n) of them: a near voxel collects thousands of points of a 200 k-point scan)
        for (int q = lane; q < SEL_BINS; q += 32) hist[q] = 0;
        __syncwarp();
        // (both passes fetch SEL_U x 32 indices per round trip: one load per iteration made a near voxel's few thousand indices
        // a chain of ~100 global-memory latencies, twice)
        constexpr int SEL_U = 8;
        for (int base = 0; base < c; base += 32 * SEL_U) {
            int v[SEL_U];
#pragma unroll
            for (int u = 0; u < SEL_U; u++) { const int q = base + 32 * u + lane; v[u] = q < c ? seg[q] : -1; }
#pragma unroll
            for (int u = 0; u < SEL_U; u++) if (v[u] >= 0) atomicAdd(&hist[v[u] >> S], 1);
        }
        __syncwarp();
        constexpr int BPL = SEL_BINS / 32;
        int h[BPL], s = 0;
#pragma unroll
        for (int u = 0; u < BPL; u++) { h[u] = hist[BPL * lane + u]; s += h[u]; }
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        int run = incl - s, kb = -1, below = 0;
#pragma unroll
        for (int u = 0; u < BPL; u++) { if (kb < 0 && run + h[u] >= K) { kb = BPL * lane + u; below = run; } run += h[u]; }
        const unsigned has = __ballot_sync(0xffffffffu, kb >= 0);
        const int src = __ffs(has) - 1;                         // K < c: some lane holds the boundary bin
        kb = __shfl_sync(0xffffffffu, kb, src);
        below = __shfl_sync(0xffffffffu, below, src);
        int w = 0, wc = 0;
        for (int base0 = 0; base0 < c; base0 += 32 * SEL_U) {
            int vv[SEL_U];
#pragma unroll
            for (int u = 0; u < SEL_U; u++) { const int q = base0 + 32 * u + lane; vv[u] = q < c ? seg[q] : -1; }
#pragma unroll
            for (int u = 0; u < SEL_U; u++) {
                const int v = vv[u];
                const int b = v >= 0 ? (v >> S) : SEL_BINS;
                const unsigned ba = __ballot_sync(0xffffffffu, b < kb), bc = __ballot_sync(0xffffffffu, b == kb);
                const unsigned lt = (1u << lane) - 1u;
                if (b < kb) sel[w + __popc(ba & lt)] = v;
                if (b == kb) cand[wc + __popc(bc & lt)] = v;
                w += __popc(ba); wc += __popc(bc);
            }
        }
        __syncwarp();
        warp_sort(cand, wc);                                     // wc <= 2^S <= CAND_CAP
        for (int q = lane; q < K - below; q += 32) sel[below + q] = cand[q];
        __syncwarp();
        warp_sort(sel, K);
        return;
    }
    // K-th smallest value by bisection on the value range (counts are warp reductions over the segment)
    int lo = 0, hi = n - 1;                                      // invariant: count(<= hi) >= K, count(<= lo - 1) < K
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        int cntl = 0;
        for (int q = lane; q < c; q += 32) cntl += (seg[q] <= mid) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cntl += __shfl_xor_sync(0xffffffffu, cntl, o);
        if (cntl >= K) hi = mid; else lo = mid + 1;
    }
    int w = 0;
    for (int base = 0; base < c; base += 32) {
        const int q = base + lane;
        const int v = q < c ? seg[q] : INT_MAX;
        const bool take = v <= lo;
        const unsigned bal = __ballot_sync(0xffffffffu, take);
        if (take) sel[w + __popc(bal & ((1u << lane) - 1))] = v;
        w += __popc(bal);
    }
    __syncwarp();
    warp_sort(sel, K);
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

// covariance of a snapshot -> eigen-decomposition (voxel_map.cpp:101-106)
__device__ __forceinline__ void snap_eig(const double* sn, double* evals, M3& evecs) {
    const double nd = sn[0];
    const double c00 = sn[5] / nd - sn[2] * sn[2];
    const double c10 = sn[6] / nd - sn[3] * sn[2];
    const double c11 = sn[7] / nd - sn[3] * sn[3];
    const double c20 = sn[8] / nd - sn[4] * sn[2];
    const double c21 = sn[9] / nd - sn[4] * sn[3];
    const double c22 = sn[10] / nd - sn[4] * sn[4];
    eig3_sym(c00, c10, c11, c20, c21, c22, evals, evecs);
}

// ---- one voxel's fill phase, by one warp (CTA = false) or by all warps of the CTA (CTA = true: a voxel that goes from empty to
// full within one scan runs ~10 refits over 10 .. 100 stored points, ~22 chunks of 32 contributions; the chunks are independent,
// only their accumulation is ordered, so the warps of the CTA compute a batch of chunks at a time and the leader warp adds them
// up in order).  The leader (warp 0 of the CTA, or the warp itself) owns the sequential parts.
// Where the stored points come from: the warp path reads them straight from the voxel's block in HBM / L2 (twelve coalesced loads
// per chunk; its shared-memory footprint is what limits how many voxels an SM works on at a time), the CTA path stages them in
// shared memory once (cp.async.bulk) for its four warps.
// Snapshots refitted together (their eigen-solves run one per lane, and an eigen-solve - ~25 k cycles on one lane - is the longest
// single step of a voxel).  The warp path keeps six (its shared-memory footprint decides how many voxels an SM works on at a time);
// the CTA path takes twelve, the ten refits of a voxel that fills up within one scan in ONE pass instead of two: -10 us on the
// longest voxel of a C2 scan, which is what the CTA launch lasts.
constexpr int FILL_GROUP = 6;
constexpr int HEAVY_GROUP = 12;
template <bool CTA> struct FillGroup { static constexpr int N = CTA ? HEAVY_GROUP : FILL_GROUP; };
constexpr int CH_MAX = FILL_GROUP * 8;              // chunks of one group (max_point_thresh <= 256)
constexpr int CH_MAX_HEAVY = HEAVY_GROUP * 8;
// group tables of the CTA path (its own area behind the tiles): chunk_k / chunk_base [CH_MAX_HEAVY] shorts, ev [HEAVY_GROUP][12], snap [HEAVY_GROUP][SNAP_W]
constexpr size_t HEAVY_TABLE_BYTES = 2 * CH_MAX_HEAVY * 2 + HEAVY_GROUP * 12 * 8 + HEAVY_GROUP * SNAP_W * 8;
static_assert(HEAVY_TABLE_BYTES % 16 == 0, "16-byte granules");

struct alignas(16) FillWork {           // control block of the voxel in shared memory (written by the leader)
    unsigned long long bar;             // (unused slot, keeps the layout 16-byte granular)
    int consumed, nt0, closes, any_refit, overflow, avail, nsnap, nchunks;
    unsigned pmask;
    int bulk;                           // the stored points of the voxel are arriving through cp.async.bulk (wait on the mbarrier before reading them)
    short chunk_k[CH_MAX], chunk_base[CH_MAX];
    double ev[FILL_GROUP][12];          // eigenvalues (3) + eigenvectors (9) of the group's snapshots
};

struct FillCounters { long long ins, full, probe, pvox, refit, rpts; };

// shared memory of one warp of k_fill: [sel] [FillWork] [cx: xyz of the consumed points, 3 x ld] [tile (+ selection scratch)] [snap];
// the CTA path appends [pts: 12 x ld, the voxel's stored + consumed points] [tiles of warps 1 ..]
__host__ __device__ inline int fill_sel_len(int maxpt) { const int ld = (maxpt + 3) & ~3; return ld < 32 ? 32 : ld; }   // a small segment is sorted whole (<= 32); 16-byte granules
__host__ __device__ inline size_t fill_warp_bytes(int maxpt) {
    const size_t ld = (size_t)((maxpt + 1) & ~1);
    return (size_t)fill_sel_len(maxpt) * 4 + sizeof(FillWork) + 3 * ld * 8 + 32 * TILE_LD * 8 + FILL_GROUP * SNAP_W * 8;     // every term a multiple of 16
}
__host__ __device__ inline size_t fill_heavy_bytes(int maxpt) {
    const size_t ld = (size_t)((maxpt + 1) & ~1);
    return fill_warp_bytes(maxpt) + 12 * ld * 8 + (size_t)(HEAVY_WARPS - 1) * 32 * TILE_LD * 8 + HEAVY_TABLE_BYTES;
}

// ---- TMA-style bulk copies (cp.async.bulk, sm_90+ / sm_100a) of a voxel's stored points (CTA path): twelve contiguous rows of the
// slot's component-major block go global -> shared asynchronously, completion counted in bytes on an mbarrier, while the leader
// selects and gathers the points of the scan; no registers, no LSU instructions per element (SASS: UBLKCP / SYNCS)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned phase) {
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct FillRegion { int* sel; FillWork* W; double* cx; double* tile; double* snap; double* pts; double* tiles_rest; int* hist; int* cand;
                    short* chunk_k; short* chunk_base; double (*ev)[12]; };
constexpr int CAND_OFF = SEL_BINS;                  // selection scratch inside the tile: [hist: SEL_BINS ints] [cand]
template <bool CTA>
__device__ __forceinline__ FillRegion fill_region(unsigned char* base, int maxpt) {
    const int ld = (maxpt + 1) & ~1;
    FillRegion r;
    r.sel = reinterpret_cast<int*>(base);
    r.W = reinterpret_cast<FillWork*>(r.sel + fill_sel_len(maxpt));
    r.cx = reinterpret_cast<double*>(r.W + 1);
    r.tile = r.cx + 3 * ld;
    r.snap = r.tile + 32 * TILE_LD;
    r.pts = r.snap + FILL_GROUP * SNAP_W;           // (CTA path only)
    r.tiles_rest = r.pts + 12 * ld;                 // (CTA path only)
    r.hist = reinterpret_cast<int*>(r.tile);
    r.cand = r.hist + CAND_OFF;
    r.chunk_k = r.W->chunk_k; r.chunk_base = r.W->chunk_base; r.ev = r.W->ev;
    if (CTA) {                                      // the bigger group tables of the CTA path
        double* t = r.tiles_rest + (size_t)(HEAVY_WARPS - 1) * 32 * TILE_LD;
        r.ev = reinterpret_cast<double (*)[12]>(t);
        r.snap = t + HEAVY_GROUP * 12;
        r.chunk_k = reinterpret_cast<short*>(r.snap + HEAVY_GROUP * SNAP_W);
        r.chunk_base = r.chunk_k + CH_MAX_HEAVY;
    }
    return r;
}

template <bool CTA>
__device__ __forceinline__ void fill_sync() { if (CTA) __syncthreads(); else __syncwarp(); }
// CTA path: the contribution tile of warp w
__device__ __forceinline__ double* cta_tile(const FillRegion& R, int w) { return w == 0 ? R.tile : R.tiles_rest + (size_t)(w - 1) * 32 * TILE_LD; }

// updatePlane() for the snapshots snap[0..ns) of the voxel, in order.  Leader registers: acc0 / acc1 (plane->cov, lane e holds
// entry e, lanes 0..3 also entry 32 + e), the normal / centre of the last refit that found a plane.
struct RefitAcc {
    double acc0, acc1;
    double nrm[3], ctr[3];
    int loaded, any_plane, plane_final, n_refit;
    long long refit_points;
};

// P / pstride: the voxel's points, component-major (P[k * pstride + q]: xyz k = 0..2, covariance k = 3..11), stored-point order
template <bool CTA>
__device__ void refit_group(const DevMap& m, DevCtl* ctl, int slot, const FillRegion& R, const double* P, int pstride, int ns, RefitAcc& ra) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const bool leader = !CTA || wib == 0;
    const int nwarps = CTA ? HEAVY_WARPS : 1, wix = CTA ? wib : 0;
    FillWork* W = R.W;
    const double* snap = R.snap;
    if (leader) {
        if (!ra.loaded) {
            const double* cv = m.cov + (size_t)slot * 36;
            ra.acc0 = cv[lane];
            ra.acc1 = lane < 4 ? cv[32 + lane] : 0.0;
            ra.loaded = 1;
        }
        // the eigen-solves of the group, one per lane
        int plane = 0, nch = 0;
        if (lane < ns) {
            double evals[3];
            M3 evecs;
            snap_eig(snap + lane * SNAP_W, evals, evecs);
            plane = !(evals[0] > m.plane_thresh);                 // Q13: otherwise norm / cov stay
#pragma unroll
            for (int e = 0; e < 3; e++) R.ev[lane][e] = evals[e];
#pragma unroll
            for (int e = 0; e < 9; e++) R.ev[lane][3 + e] = evecs.a[e];
            if (plane) {
                int nt = (int)snap[lane * SNAP_W + 1];
                if (nt > W->avail) { atomicOr(&ctl->err, E_REFIT_OVERFLOW); nt = W->avail; }      // build overflow + thresh 1
                nch = (nt + 31) >> 5;
            }
        }
        const unsigned pmask = __ballot_sync(0xffffffffu, plane != 0);
        int incl = nch;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        const int first = incl - nch;
        for (int b = 0; b < nch; b++) { R.chunk_k[first + b] = (short)lane; R.chunk_base[first + b] = (short)(32 * b); }
        if (lane == 31) { W->nchunks = incl; W->pmask = pmask; }
    }
    fill_sync<CTA>();
    const int nchunks = W->nchunks;
    const unsigned pmask = W->pmask;
    for (int batch = 0; batch < nchunks; batch += nwarps) {
        const int ch = batch + wix;
        if (ch < nchunks) {
            const int k = R.chunk_k[ch], base = R.chunk_base[ch];
            const double* sn = snap + k * SNAP_W;
            const int n = (int)sn[0];
            int nt = (int)sn[1];
            if (nt > W->avail) nt = W->avail;
            const int q = base + lane;
            if (q < nt) {
                V3 p;
                M3 S;
                p[0] = P[q]; p[1] = P[pstride + q]; p[2] = P[2 * pstride + q];
#pragma unroll
                for (int e = 0; e < 9; e++) S.a[e] = P[(3 + e) * pstride + q];
                const V3 mean = v3(sn[2], sn[3], sn[4]);
                double ev[3];
                M3 evc;
#pragma unroll
                for (int e = 0; e < 3; e++) ev[e] = R.ev[k][e];
#pragma unroll
                for (int e = 0; e < 9; e++) evc.a[e] = R.ev[k][3 + e];
                const V3 nrm = v3(evc(0, 0), evc(1, 0), evc(2, 0));
                double* tile = CTA ? cta_tile(R, wib) : R.tile;
                plane_contrib(p, S, mean, n, ev, evc, nrm, tile + lane * TILE_LD);
            }
        }
        fill_sync<CTA>();
        if (leader) {
            const int nb = nchunks - batch < nwarps ? nchunks - batch : nwarps;
            for (int w = 0; w < nb; w++) {                            // chunks in order, points in stored order (Q7)
                const int k = R.chunk_k[batch + w], base = R.chunk_base[batch + w];
                int nt = (int)snap[k * SNAP_W + 1];
                if (nt > W->avail) nt = W->avail;
                const int np = nt - base < 32 ? nt - base : 32;
                const double* tile = CTA ? cta_tile(R, w) : R.tile;
                for (int qq = 0; qq < np; qq++) {
                    ra.acc0 += tile[qq * TILE_LD + lane];
                    if (lane < 4) ra.acc1 += tile[qq * TILE_LD + 32 + lane];
                }
            }
        }
        fill_sync<CTA>();
    }
    if (leader) {
        ra.n_refit += ns;
        ra.plane_final = (pmask >> (ns - 1)) & 1u;
        for (int k = 0; k < ns; k++) if ((pmask >> k) & 1u) ra.refit_points += (int)snap[k * SNAP_W + 1];
        if (pmask) {
            ra.any_plane = 1;
            const int k = 31 - __clz(pmask);                          // the last refit that found a plane
            const double* sn = snap + k * SNAP_W;
            const V3 mean = v3(sn[2], sn[3], sn[4]);
            const V3 nrm = v3(R.ev[k][3], R.ev[k][6], R.ev[k][9]);
            V3 ns_ = nrm;
            if (-dot(mean, nrm) < 0.0) ns_ = neg(nrm);
#pragma unroll
            for (int e = 0; e < 3; e++) { ra.nrm[e] = ns_[e]; ra.ctr[e] = mean[e]; }
        }
    }
}

// region: the shared-memory region the voxel is staged in (the warp's own, or the CTA's)
template <bool CTA>
__device__ void fill_voxel(const DevMap& m, const DevScan& s, DevCtl* ctl, int slot, unsigned char* region, int npts, unsigned scan_id,
                           FillCounters& fc, unsigned long long* bar, unsigned& phase) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const bool leader = !CTA || wib == 0;
    const int nth = CTA ? HEAVY_WARPS * 32 : 32, tix = CTA ? (int)threadIdx.x : lane;
    const int ld = (m.maxpt + 1) & ~1;
    const FillRegion R = fill_region<CTA>(region, m.maxpt);
    FillWork* W = R.W;
    double* tp = m.tp + (size_t)slot * 12 * m.maxpt;
    // ---- leader: counters of the voxel, closed-form control flow, selection of the consumed points
    uint32_t flags = 0;
    int n = 0, c = 0, nw0 = 0, n0 = 0, jA = 0, j_init = 0, jc = 0, next_refit = 0, K = 0;
    bool init0 = false;
    double mean_l = 0.0, ppt_l = 0.0;
    // addToPlane (voxel_map.cpp:29-34) is a dependent chain per point; its nine scalar updates run on nine lanes' worth of
    // registers instead of one lane's: lane l < 3 owns mean[l], lane l < 6 owns ppt[l] (xx, yx, yy, zx, zy, zz).
    const int cm = lane < 3 ? lane : 0;
    const int ia = lane == 0 ? 0 : lane < 3 ? 1 : lane < 6 ? 2 : 0;
    const int ib = (lane == 2 || lane == 4) ? 1 : lane == 5 ? 2 : 0;
    if (leader) {
        c = m.cnt[slot];
        const int off = m.seg_off[slot];
        hot_get_fn(m.hot, slot, flags, n);
        if (!(flags & F_UE)) {
            if (lane == 0) W->consumed = -1;                      // full before this scan: merge() or nothing per point
        } else {
            const int nt0 = m.n_temp[slot];
            nw0 = m.newly[slot];
            mean_l = m.hot[(size_t)slot * 8 + cm];
            ppt_l = m.ppt[(size_t)slot * 6 + (lane < 6 ? lane : 0)];
            // the voxel consumes at most K points before it closes (never more than its free room, at least one)
            const int room = m.maxpt - nt0;
            K = room < 1 ? 1 : room;
            if (K > c) K = c;
            // a voxel that build() left with more than max_point_thresh stored points (Q18) consumes one point and closes; a refit
            // of such a voxel cannot be served (its points beyond max_point_thresh are not kept) and is reported: E_REFIT_OVERFLOW
            const bool overflow = nt0 + K > m.maxpt;
            // pushPoint's control flow (voxel_map.cpp:42-95) depends on the counters only, never on the points: which steps
            // refit, which step closes the voxel and the final counters have a closed form in (n, n_temp, newly_add_point, K)
            // (checked against the step-by-step state machine on 6.2e6 parameter combinations).
            init0 = (flags & F_INIT) != 0;
            n0 = n;
            jA = init0 ? -1 : (m.upt - 1 - n0 > 0 ? m.upt - 1 - n0 : 0);       // first updatePlane() past the early return (voxel not yet initialised)
            j_init = init0 ? 0 : jA + 1;                                       // first step that sees is_init == true
            jc = m.maxpt - nt0 - 1;                                            // closing step: is_init before it, temp_points.size() >= max_point_thresh after it
            if (jc < j_init) jc = j_init;
            const bool closes = jc <= K - 1;
            const int consumed = closes ? jc + 1 : K;
            next_refit = init0 ? m.upt - nw0 - 1 : jA;
            const bool any_refit = next_refit < consumed;
            // CTA path: the stored points of earlier scans go to shared memory through twelve bulk copies, issued now so that they fly
            // while the points of this scan are selected and gathered.  Rows start 16-byte aligned when max_point_thresh is even.
            const bool bulk = CTA && any_refit && !overflow && nt0 >= 2 && (m.maxpt & 1) == 0;
            if (lane == 0) {
                W->consumed = consumed; W->nt0 = nt0; W->closes = closes; W->any_refit = any_refit;
                W->overflow = overflow; W->avail = overflow ? 0 : nt0 + consumed;
                W->bulk = bulk;
                if (bulk) {
                    const unsigned bytes = (unsigned)((nt0 & ~1) * 8);                // whole 16-byte granules; an odd last point is copied by hand below
                    fence_proxy_async();                                           // earlier generic-proxy writes to the staging before the async ones
                    mbar_expect_tx(bar, 12 * bytes);
#pragma unroll
                    for (int k = 0; k < 12; k++) bulk_g2s(R.pts + k * ld, tp + (size_t)k * m.maxpt, bytes, bar);
                }
            }
            warp_select_sorted(m.seg + off, c, K, npts, R.sel, R.hist, R.cand);      // K <= max_point_thresh
        }
    }
    fill_sync<CTA>();
    const int consumed = W->consumed;
    int events = 0;
    if (consumed < 0) {
        events = c;
    } else {
        const int nt0 = W->nt0;
        const bool closes = W->closes != 0, any_refit = W->any_refit != 0, overflow = W->overflow != 0;
        // ---- gather the consumed points once, in parallel (twelve loads in flight per point): xyz -> shared (the state machine walks them
        // in order), point + covariance -> behind the voxel's stored points (temp_points.push_back; also when the voxel closes in this
        // scan - its refits of this scan read them from there) and, CTA path, into the shared staging
        for (int q = tix; q < consumed; q += nth) {
            const size_t i = (size_t)R.sel[q];
            double v[12];
            v[0] = s.pw[3 * i]; v[1] = s.pw[3 * i + 1]; v[2] = s.pw[3 * i + 2];
#pragma unroll
            for (int k = 0; k < 9; k++) v[3 + k] = s.pcov[9 * i + k];
#pragma unroll
            for (int k = 0; k < 3; k++) R.cx[k * ld + q] = v[k];
            if (!overflow) {
#pragma unroll
                for (int k = 0; k < 12; k++) tp[(size_t)k * m.maxpt + nt0 + q] = v[k];
                if (CTA) {
#pragma unroll
                    for (int k = 0; k < 12; k++) R.pts[k * ld + nt0 + q] = v[k];
                }
            }
        }
        if (CTA && any_refit && !overflow) {
            if (W->bulk) {                                          // wait for the twelve rows (byte count on the mbarrier)
                if ((nt0 & 1) && tix < 12) R.pts[tix * ld + nt0 - 1] = tp[(size_t)tix * m.maxpt + nt0 - 1];
                while (!mbar_try_wait(bar, phase)) {}
                phase ^= 1u;
            } else {
                for (int q = tix; q < nt0; q += nth) {
                    double v[12];
#pragma unroll
                    for (int k = 0; k < 12; k++) v[k] = tp[(size_t)k * m.maxpt + q];
#pragma unroll
                    for (int k = 0; k < 12; k++) R.pts[k * ld + q] = v[k];
                }
            }
        }
        fill_sync<CTA>();
        const double* P = CTA ? R.pts : tp;
        const int pstride = CTA ? ld : m.maxpt;
        // ---- the state machine (leader); the refits of every FillGroup<CTA>::N snapshots by everybody
        RefitAcc ra;
        ra.acc0 = ra.acc1 = 0.0; ra.loaded = 0; ra.any_plane = 0; ra.plane_final = 0; ra.n_refit = 0; ra.refit_points = 0;
        for (int e = 0; e < 3; e++) { ra.nrm[e] = 0.0; ra.ctr[e] = 0.0; }
        int j = 0;
        while (true) {
            int nsnap = 0;
            if (leader) {
                for (; j < consumed && nsnap < FillGroup<CTA>::N; j++) {   // point order
                    const double pm = R.cx[cm * ld + j], pa = R.cx[ia * ld + j], pb = R.cx[ib * ld + j];
                    mean_l = mean_l + (pm - mean_l) / (double)(n0 + j + 1);
                    ppt_l += pa * pb;
                    if (j == next_refit) {
                        next_refit = (!init0 && j == jA) ? j_init + (m.upt - nw0) - 1 : next_refit + m.upt;
                        double* sn = R.snap + nsnap * SNAP_W;
                        if (lane == 0) { sn[0] = n0 + j + 1; sn[1] = nt0 + j + 1; }
                        if (lane < 3) sn[2 + lane] = mean_l;
                        if (lane < 6) sn[5 + lane] = ppt_l;
                        nsnap++;
                    }
                }
                if (lane == 0) W->nsnap = nsnap;
            }
            fill_sync<CTA>();
            nsnap = W->nsnap;
            if (nsnap == 0) break;                                         // (the leader's loop has reached `consumed`)
            refit_group<CTA>(m, ctl, slot, R, P, pstride, nsnap, ra);
            fill_sync<CTA>();
        }
        // ---- write-back
        if (leader) {
            n = n0 + consumed;
            const int nt = closes ? 0 : nt0 + consumed;                    // closing frees temp_points
            int nw;
            { const int sdone = consumed - j_init; nw = (nw0 + (sdone > 0 ? sdone : 0)) % m.upt; }
            if (!init0 && consumed - 1 >= jA) flags |= F_INIT;
            if (closes) flags &= ~F_UE;
            if (!closes && K < c && lane == 0) atomicOr(&ctl->err, E_QUEUE);          // cannot happen: K points always close the voxel
            if (ra.n_refit > 0) flags = ra.plane_final ? (flags | F_PLANE) : (flags & ~F_PLANE);
            if (lane < 3) m.hot[(size_t)slot * 8 + lane] = mean_l;
            if (lane < 6) m.ppt[(size_t)slot * 6 + lane] = ppt_l;
            if (ra.any_plane) {
                double* cv = m.cov + (size_t)slot * 36;
                cv[lane] = ra.acc0;
                if (lane < 4) cv[32 + lane] = ra.acc1;
                if (lane < 3) {      // (selects: a lane-indexed ra.nrm[lane] would move the accumulator struct to local memory)
                    m.hot[(size_t)slot * 8 + 3 + lane] = lane == 0 ? ra.nrm[0] : lane == 1 ? ra.nrm[1] : ra.nrm[2];
                    m.center[(size_t)slot * 3 + lane] = lane == 0 ? ra.ctr[0] : lane == 1 ? ra.ctr[1] : ra.ctr[2];
                }
            }
            if (lane == 0) {
                hot_set_fn(m.hot, slot, flags, n);
                m.n_temp[slot] = nt; m.newly[slot] = nw;
                if (closes) { m.full_scan[slot] = scan_id; m.full_idx[slot] = R.sel[jc]; }
            }
            fc.ins += consumed;
            fc.refit += ra.n_refit;
            fc.rpts += ra.refit_points;
            events = c - consumed;
        }
    }
    if (leader) {
        fc.full += events;
        // merge() runs for the points that land in a full plane voxel (Q11); everything else is inert (Q12)
        if (!(flags & F_UE) && (flags & F_PLANE)) { fc.probe += events; fc.pvox += events > 0 ? 1 : 0; } else events = 0;
        if (lane == 0) m.evn[slot] = events;
    }
    fill_sync<CTA>();
}

// which touched voxels go to the CTA path of k_fill: those whose refits of this scan loop over many stored points (estimated
// from the counters alone; a hint - either path handles every voxel)
__device__ __forceinline__ void fill_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ void fill_classify(const DevMap& m, DevCtl* ctl, int vi) {
    {
        const int slot = m.touched[vi];
        uint32_t flags; int n;
        hot_get_fn(m.hot, slot, flags, n);
        int heavy = 0;
        if (flags & F_UE) {
            const int c = m.cnt[slot], nt0 = m.n_temp[slot], nw0 = m.newly[slot];
            const int room = m.maxpt - nt0;
            int K = room < 1 ? 1 : room;
            if (K > c) K = c;
            const int nref = (K + nw0) / m.upt;
            heavy = m.heavy_points > 0 && nref * (nt0 + (K + 1) / 2) >= m.heavy_points;
            // the map data of a voxel is cold at the start of a scan (the L2 does not keep a 100 000-voxel map between scans):
            // pull what k_fill reads about a filling voxel into the L2 one kernel ahead (k_fill is a chain of dependent loads per voxel)
            fill_prefetch_l2(m.ppt + (size_t)slot * 6);
            if (nref > 0) {
                const char* cv = reinterpret_cast<const char*>(m.cov + (size_t)slot * 36);
                fill_prefetch_l2(cv); fill_prefetch_l2(cv + 128); fill_prefetch_l2(cv + 256);
                const double* tp = m.tp + (size_t)slot * 12 * m.maxpt;
                for (int k = 0; k < 12; k++)
                    for (int q = 0; q < nt0; q += 16) fill_prefetch_l2(tp + (size_t)k * m.maxpt + q);
            }
        }
        m.vox_cls[vi] = heavy;
        if (heavy) m.hotlist[atomicAdd(&ctl->n_heavy, 1)] = slot;
    }
}

// heavy != 0: the CTA path over the voxels classified heavy (shared memory: fill_heavy_bytes); heavy == 0: a warp per voxel over the
// rest (FILL_WARPS x fill_warp_bytes).  The two launches run side by side on two streams of the scan's graph.
template <bool HEAVY>
__global__ void __launch_bounds__(HEAVY ? HEAVY_WARPS * 32 : FILL_WARPS * 32, 3) k_fill(DevMap m, DevScan s, DevCtl* ctl) {
    constexpr int heavy = HEAVY ? 1 : 0;
    extern __shared__ __align__(16) unsigned char fill_smem[];
    __shared__ int s_vi;
    __shared__ __align__(8) unsigned long long s_bar;         // mbarrier of the CTA path's bulk copies
    const int lane = threadIdx.x & 31;
    const size_t wbytes = fill_warp_bytes(m.maxpt);
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    fence_proxy_async();
    __syncthreads();
    unsigned phase_c = 0;
    const int V = ctl->n_touched, npts = ctl->n, NH = ctl->n_heavy;
    const unsigned scan_id = ctl->scan_id;
    FillCounters fc = {0, 0, 0, 0, 0, 0};
    // heavy voxels: one per CTA at a time
    while (heavy) {
        if (threadIdx.x == 0) s_vi = atomicAdd(&ctl->heavy_next, 1);
        __syncthreads();
        const int hi = s_vi;
        __syncthreads();
        if (hi >= NH) break;
        fill_voxel<true>(m, s, ctl, m.hotlist[hi], fill_smem, npts, scan_id, fc, &s_bar, phase_c);
    }
    // everything else: a warp per voxel
    while (!heavy) {
        int vi = 0;
        if (lane == 0) vi = atomicAdd(&ctl->fill_next, 1);
        vi = __shfl_sync(0xffffffffu, vi, 0);
        if (vi >= V) break;
        if (m.vox_cls[vi]) continue;
        fill_voxel<false>(m, s, ctl, m.touched[vi], fill_smem + (size_t)(threadIdx.x >> 5) * wbytes, npts, scan_id, fc, &s_bar, phase_c);
    }
    if (lane == 0) {
        if (fc.ins) atomicAdd((unsigned long long*)&ctl->st.n_ins, (unsigned long long)fc.ins);
        if (fc.full) atomicAdd((unsigned long long*)&ctl->st.n_full, (unsigned long long)fc.full);
        if (fc.probe) atomicAdd((unsigned long long*)&ctl->st.n_mergeprobe, (unsigned long long)fc.probe);
        if (fc.pvox) atomicAdd((unsigned long long*)&ctl->st.n_mergevox, (unsigned long long)fc.pvox);
        if (fc.refit) atomicAdd((unsigned long long*)&ctl->st.n_refit, (unsigned long long)fc.refit);
        if (fc.rpts) atomicAdd((unsigned long long*)&ctl->st.refit_points, (unsigned long long)fc.rpts);
    }
}

// VoxelMap::build (voxel_map.cpp:200-230): every point of the voxel in order, no cap (Q18), one updatePlane() at the end.
// One warp per touched voxel, points straight from the scan's pv_list through the (sorted) segment; runs once per map.
__global__ void __launch_bounds__(128) k_fill_build(DevMap m, DevScan s, DevCtl* ctl) {
    __shared__ double tile_all[4][32 * TILE_LD];
    __shared__ double snap_all[4][SNAP_W];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* tile = tile_all[wib];
    double* snap = snap_all[wib];
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
    const int V = ctl->n_touched;
    long long c_ins = 0, c_refit = 0, c_rpts = 0;
    for (int vi = wg; vi < V; vi += nW) {
        const int slot = m.touched[vi];
        const int c = m.cnt[slot], off = m.seg_off[slot];
        uint32_t flags; int n;
        hot_get_fn(m.hot, slot, flags, n);
        const int nt0 = m.n_temp[slot];
        const int cm = lane < 3 ? lane : 0;
        const int ia = lane == 0 ? 0 : lane < 3 ? 1 : lane < 6 ? 2 : 0;
        const int ib = (lane == 2 || lane == 4) ? 1 : lane == 5 ? 2 : 0;
        double mean_l = m.hot[(size_t)slot * 8 + cm];
        double ppt_l = m.ppt[(size_t)slot * 6 + (lane < 6 ? lane : 0)];
        int* order = m.seg + off;
        warp_sort(order, c);
        for (int j = 0; j < c; j++) {                                   // addPoint for every point (voxel_map.cpp:218-224)
            const size_t i3 = 3 * (size_t)order[j];
            const double pm = s.pw[i3 + cm], pa = s.pw[i3 + ia], pb = s.pw[i3 + ib];
            mean_l = mean_l + (pm - mean_l) / (n + 1.0);
            ppt_l += pa * pb;
            n += 1;
        }
        double* tp = m.tp + (size_t)slot * 12 * m.maxpt;
        for (int q = lane; q < c && nt0 + q < m.maxpt; q += 32) {       // temp_points (the first max_point_thresh of them are kept)
            const int i = order[q];
#pragma unroll
            for (int k = 0; k < 3; k++) tp[(size_t)k * m.maxpt + nt0 + q] = s.pw[3 * (size_t)i + k];
#pragma unroll
            for (int k = 0; k < 9; k++) tp[(size_t)(3 + k) * m.maxpt + nt0 + q] = s.pcov[9 * (size_t)i + k];
        }
        const int nt = nt0 + c;
        if (n >= m.upt) {                                               // updatePlane() (voxel_map.cpp:226-229, 97-136)
            flags |= F_INIT;
            if (lane == 0) { snap[0] = n; snap[1] = nt; }
            if (lane < 3) snap[2 + lane] = mean_l;
            if (lane < 6) snap[5 + lane] = ppt_l;
            __syncwarp();
            double evals[3];
            M3 evecs;
            snap_eig(snap, evals, evecs);                               // (all lanes: same cost as one)
            c_refit += 1;
            const bool plane = !(evals[0] > m.plane_thresh);
            if (plane) {
                const V3 mean = v3(snap[2], snap[3], snap[4]);
                const V3 nrm = v3(evecs(0, 0), evecs(1, 0), evecs(2, 0));
                double* cv = m.cov + (size_t)slot * 36;
                double acc0 = cv[lane], acc1 = lane < 4 ? cv[32 + lane] : 0.0;
                c_rpts += nt;
                for (int base = 0; base < c; base += 32) {              // (nt0 == 0: build() starts from an empty map)
                    const int q = base + lane;
                    if (q < c) {
                        const size_t i = (size_t)order[q];
                        const V3 p = v3(s.pw[3 * i], s.pw[3 * i + 1], s.pw[3 * i + 2]);
                        M3 S;
#pragma unroll
                        for (int k = 0; k < 9; k++) S.a[k] = s.pcov[9 * i + k];
                        plane_contrib(p, S, mean, n, evals, evecs, nrm, tile + lane * TILE_LD);
                    }
                    __syncwarp();
                    const int np = c - base < 32 ? c - base : 32;
                    for (int qq = 0; qq < np; qq++) {
                        acc0 += tile[qq * TILE_LD + lane];
                        if (lane < 4) acc1 += tile[qq * TILE_LD + 32 + lane];
                    }
                    __syncwarp();
                }
                cv[lane] = acc0;
                if (lane < 4) cv[32 + lane] = acc1;
                V3 ns_ = nrm;
                if (-dot(mean, nrm) < 0.0) ns_ = neg(nrm);
                if (lane < 3) { m.hot[(size_t)slot * 8 + 3 + lane] = ns_[lane]; m.center[(size_t)slot * 3 + lane] = mean[lane]; }
                flags |= F_PLANE;
            } else {
                flags &= ~F_PLANE;
            }
            if (nt0 != 0 && lane == 0) atomicOr(&ctl->err, E_REFIT_OVERFLOW);      // build() on a non-empty voxel is not supported
        }
        c_ins += c;
        __syncwarp();
        if (lane < 3) m.hot[(size_t)slot * 8 + lane] = mean_l;
        if (lane < 6) m.ppt[(size_t)slot * 6 + lane] = ppt_l;
        if (lane == 0) {
            hot_set_fn(m.hot, slot, flags, n);
            m.n_temp[slot] = nt;
            m.evn[slot] = 0;
        }
        __syncwarp();
    }
    if (lane == 0) {
        if (c_ins) atomicAdd((unsigned long long*)&ctl->st.n_ins, (unsigned long long)c_ins);
        if (c_refit) atomicAdd((unsigned long long*)&ctl->st.n_refit, (unsigned long long)c_refit);
        if (c_rpts) atomicAdd((unsigned long long*)&ctl->st.refit_points, (unsigned long long)c_rpts);
    }
}
