// vmp_solve.cuh — the 23-dof algebra of IESKF::update (ieskf.cpp:125-156) on the device.
// Included by vmp_iekf.cu inside namespace vmp.
//
//   k_update_begin   predict_x = x_ (ieskf.cpp:127), iteration counters, and P_^-1: P_ does not change
//                    inside update(), so the inverse the reference evaluates twice per iteration (Q6) is
//                    computed once per scan here (on a forked branch of the scan graph)
//   k_ieskf_solve    one CTA per iteration: fixed-order reduction of the measurement partials, boxminus,
//                    J, H_ = J^T P^-1 J + H, b_, delta = -H_^-1 b_, boxplus, convergence flag; on the
//                    last executed iteration also P_ = L H_^-1 L^T
//
// The 23x23 inverse is LU with partial pivoting + substitution against the permuted identity (what
// Eigen's PartialPivLU-based inverse() does).  ncu on the first versions showed that a single warp
// doing this is bound by its own instruction count (23k warp-instructions, ~4.6 cycles each), so the
// work is spread over the CTA, two matrix entries per thread: per elimination step one pivot search
// by warp 0 and one rank-1 update by everybody (two barriers), per substitution step one barrier.
// Rows are never moved (a position array tracks the pivot order), and each entry sees exactly the
// serial sequence of operations of lu_inverse<> in vmp_math.cuh -> bit-identical results.
#pragma once

constexpr int NS = 23;
constexpr int NE = NS * NS;
constexpr int SOLVE_THREADS = 288;      // 9 warps; entries tid and tid + 288 (< 529) per thread
constexpr int W_MAN = 8;                // the warp that runs the manifold operations
constexpr int N_WORK = 32 * W_MAN;      // threads of the other warps

struct LuShared {
    double Lm[NE];                      // multipliers by physical row
    int pos[NS];                        // physical row -> index in pivot order
    int perm[NS];                       // pivot order -> physical row
    int piv;                            // physical pivot row of the current step
};

// A: NE doubles in shared, destroyed (ends as U by physical row); inv: NE out.  Whole CTA must call.
__device__ void block_lu_inverse(double* A, double* inv, LuShared& w) {
    const int tid = threadIdx.x;
    const int e0 = tid, e1 = tid + SOLVE_THREADS;
    const bool h1 = e1 < NE;
    const int r0 = e0 / NS, c0 = e0 % NS, r1 = h1 ? e1 / NS : 0, c1 = h1 ? e1 % NS : 0;
    // Every warp runs the (cheap) pivot search redundantly on lanes 0..22, so there is no barrier between the
    // search and the rank-1 update; each thread tracks the pivot-order position of "its" rows in registers.
    const int lane = tid & 31;
    const bool act = lane < NS;
    int lpos = act ? lane : 1000 + lane;                 // position of physical row `lane`
    int pos0 = r0, pos1 = r1;                            // positions of the rows of this thread's two entries
    for (int k = 0; k < NS; k++) {
        // pivot: first row (in pivot order) of maximal |a_ik|, i >= k.  |a| >= 0 orders like its bit pattern:
        // three warp reductions (high word, low word, smallest position) instead of a 5-deep shuffle chain.
        const bool cand = act && lpos >= k;
        const unsigned long long bits = cand ? (unsigned long long)__double_as_longlong(fabs(A[lane * NS + k])) : 0ull;
        const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
        const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
        const bool c1_ = cand && hi == mhi;
        const unsigned mlo = __reduce_max_sync(0xffffffffu, c1_ ? lo : 0u);
        const bool c2_ = c1_ && lo == mlo;
        const int p = (int)__reduce_min_sync(0xffffffffu, c2_ ? (unsigned)lpos : 0x7fffffffu);
        const int L = __ffs(__ballot_sync(0xffffffffu, act && lpos == p)) - 1;      // physical pivot row
        if (lpos == p) lpos = k; else if (lpos == k) lpos = p;                        // row swap k <-> p in the pivot order
        if (pos0 == p) pos0 = k; else if (pos0 == k) pos0 = p;
        if (pos1 == p) pos1 = k; else if (pos1 == k) pos1 = p;
        const double d = A[L * NS + k];
        // rank-1 update.  No hazard inside the phase: column k (read by everybody) is only ever written to
        // Lm, the pivot row L is not written at all, and A[r][c] is read and written by its own thread only.
        if (pos0 > k && c0 >= k) {
            const double l = A[r0 * NS + k] / d;
            if (c0 == k) w.Lm[e0] = l; else A[e0] = A[e0] - l * A[L * NS + c0];
        }
        if (h1 && pos1 > k && c1 >= k) {
            const double l = A[r1 * NS + k] / d;
            if (c1 == k) w.Lm[e1] = l; else A[e1] = A[e1] - l * A[L * NS + c1];
        }
        __syncthreads();
    }
    if (tid < NS) w.perm[lpos] = tid;
    __syncthreads();
    // thread entries (i, c): entry i (pivot order) of column c of the inverse
    const int p0 = w.perm[r0], p1 = h1 ? w.perm[r1] : 0;
    double s0 = (p0 == c0) ? 1.0 : 0.0, s1 = (h1 && p1 == c1) ? 1.0 : 0.0;
    for (int j = 0; j < NS; j++) {                       // L y = P e_c
        if (r0 == j) inv[j * NS + c0] = s0;
        if (h1 && r1 == j) inv[j * NS + c1] = s1;
        __syncthreads();
        if (r0 > j) s0 = s0 - w.Lm[p0 * NS + j] * inv[j * NS + c0];
        if (h1 && r1 > j) s1 = s1 - w.Lm[p1 * NS + j] * inv[j * NS + c1];
    }
    __syncthreads();
    for (int j = NS - 1; j >= 0; j--) {                  // U x = y, unknowns in the order they become available
        if (r0 == j) inv[j * NS + c0] = s0 / A[p0 * NS + j];
        if (h1 && r1 == j) inv[j * NS + c1] = s1 / A[p1 * NS + j];
        __syncthreads();
        if (r0 < j) s0 = s0 - A[p0 * NS + j] * inv[j * NS + c0];
        if (h1 && r1 < j) s1 = s1 - A[p1 * NS + j] * inv[j * NS + c1];
    }
    __syncthreads();
}

// C = op(A) * op(B), NS x NS in shared memory, entries strided over the block, left-to-right sums
__device__ __forceinline__ void block_mm(const double* A, const double* B, double* C, bool transA, bool transB) {
    const int sa = transA ? NS : 1, sb = transB ? 1 : NS;            // strides along k
    for (int q = threadIdx.x; q < NE; q += blockDim.x) {
        const int i = q / NS, j = q % NS;
        const double* a = A + (transA ? i : i * NS);
        const double* b = B + (transB ? j * NS : j);
        double s = a[0] * b[0];
#pragma unroll
        for (int k = 1; k < NS; k++) s += a[k * sa] * b[k * sb];
        C[q] = s;
    }
}

// the three non-identity blocks of J / L (ieskf.cpp:136-139, 151-154), one lane each
__device__ __noinline__ void jac_blocks(double* J, const double* delta, const double* g_cur, const double* g_pred, int which) {
    if (which == 0 || which == 1) {
        const int o = which == 0 ? 3 : 6;
        const M3 j = right_jacobian(v3(delta[o], delta[o + 1], delta[o + 2]));
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) J[(o + a) * NS + o + b] = j(a, b);
    } else {
        Mat<2, 1> dg; dg[0] = delta[21]; dg[1] = delta[22];
        const Mat<2, 2> jg = mul(st_Nx(v3(g_cur[0], g_cur[1], g_cur[2])), st_Mx_res(v3(g_pred[0], g_pred[1], g_pred[2]), dg));
        J[21 * NS + 21] = jg(0, 0); J[21 * NS + 22] = jg(0, 1); J[22 * NS + 21] = jg(1, 0); J[22 * NS + 22] = jg(1, 1);
    }
}

// P_^-1 only; runs on a forked branch of the scan graph, concurrently with k_set_scan and the first
// k_measure (predict_x / counters are initialised by k_scan_in)
__global__ void __launch_bounds__(SOLVE_THREADS) k_update_begin(DevFilter* f, DevCtl* ctl) {
    __shared__ double sA[NE], sInv[NE];
    __shared__ LuShared lu;
    const int tid = threadIdx.x;
    (void)ctl;
    for (int q = tid; q < NE; q += SOLVE_THREADS) sA[q] = f->P[q];
    __syncthreads();
    block_lu_inverse(sA, sInv, lu);
    for (int q = tid; q < NE; q += SOLVE_THREADS) f->Pinv[q] = sInv[q];
}
void launch_update_begin(cudaStream_t st, DevFilter* f, DevCtl* ctl) { k_update_begin<<<1, SOLVE_THREADS, 0, st>>>(f, ctl); }

constexpr int RED_CHUNKS = 2;

template <bool EXT>
__global__ void __launch_bounds__(SOLVE_THREADS) k_ieskf_solve(DevFilter* f, DevCtl* ctl, const double* __restrict__ partials, int nblocks) {
    constexpr int D = EXT ? 12 : 6;
    constexpr int NH = D * (D + 1) / 2;
    constexpr int NV = NH + D + 1;
    static_assert(NV * RED_CHUNKS <= N_WORK, "reduction does not fit the worker threads");
    __shared__ double sA[NE], sB[NE], sC[NE], sJ[NE], sHinv[NE];
    __shared__ double sHm[NV], sRed[RED_CHUNKS][NV], sdelta[NS], sb[NS], sdx[NS], sx[36], sxp[36];
    __shared__ LuShared lu;
    __shared__ int s_last;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    if (ctl->done) return;
    const int it = ctl->iter;
    long long tk0 = clock64(), tk1 = 0, tk2 = 0, tk3 = 0, tk4 = 0, tk5 = 0;

    // (1) concurrently: reduction of the per-block partials (fixed order: RED_CHUNKS contiguous block
    //     ranges, then the chunk sums in ascending order) | boxminus | P^-1 load | J := I
    if (wid == W_MAN) {
        for (int q = lane; q < 36; q += 32) { sx[q] = f->x[q]; sxp[q] = f->xpred[q]; }
        __syncwarp();
        if (lane == 0) { const St x = st_load(sx), xp = st_load(sxp); st_boxminus(x, xp, sdelta); }
    } else {
        if (tid < NV * RED_CHUNKS) {
            const int v = tid % NV, c = tid / NV;
            const int b0 = (int)((long long)nblocks * c / RED_CHUNKS), b1 = (int)((long long)nblocks * (c + 1) / RED_CHUNKS);
            double t = 0.0;
#pragma unroll 8
            for (int b = b0; b < b1; b++) t += partials[(size_t)b * PARTIAL_STRIDE + v];
            sRed[c][v] = t;
        }
        for (int q = tid; q < NE; q += N_WORK) { sB[q] = f->Pinv[q]; sJ[q] = (q / NS == q % NS) ? 1.0 : 0.0; }
    }
    __syncthreads();
    tk1 = clock64();
    if (tid < NV) {
        double t = sRed[0][tid];
#pragma unroll
        for (int c = 1; c < RED_CHUNKS; c++) t += sRed[c][tid];
        sHm[tid] = t;
    }
    // (2) the blocks of J (ieskf.cpp:136-139)
    if (wid == W_MAN && lane < 3) jac_blocks(sJ, sdelta, sx + 33, sxp + 33, lane);
    __syncthreads();
    tk2 = clock64();
    // (3) JtPinv = J^T P^-1 ; b_ = JtPinv delta ; H_ = JtPinv J (+ measurement H, b in the top-left corner)
    block_mm(sJ, sB, sC, true, false);
    __syncthreads();
    if (wid == W_MAN && lane < NS) {
        double t = sC[lane * NS] * sdelta[0];
        for (int k = 1; k < NS; k++) t += sC[lane * NS + k] * sdelta[k];
        t = 0.0 + t;
        if (lane < D) t += sHm[NH + lane];
        sb[lane] = t;
    }
    for (int q = tid; q < NE; q += SOLVE_THREADS) {
        const int i = q / NS, j = q % NS;
        double h = sC[i * NS] * sJ[j];
#pragma unroll
        for (int k = 1; k < NS; k++) h += sC[i * NS + k] * sJ[k * NS + j];
        h = 0.0 + h;
        if (i < D && j < D) { const int a = i < j ? i : j, c = i < j ? j : i; h += sHm[a * D - a * (a - 1) / 2 + (c - a)]; }
        sA[q] = h;
    }
    __syncthreads();
    // (4) H_^-1, delta = -H_^-1 b_
    tk3 = clock64();
    block_lu_inverse(sA, sHinv, lu);
    tk4 = clock64();
    if (wid == W_MAN) {
        if (lane < NS) {
            double t = (-sHinv[lane * NS]) * sb[0];
            for (int k = 1; k < NS; k++) t += (-sHinv[lane * NS + k]) * sb[k];
            sdx[lane] = t;
        }
        __syncwarp();
        // (5) x_ += delta, counters, convergence on the signed maximum (Q5)
        if (lane == 0) {
            St x = st_load(sx);
            st_boxplus(x, sdx);
            st_store(x, sx);
            st_store(x, f->x);
            ctl->effect[it & 7] = (int)sHm[NH + D];
            const int nit = it + 1;
            ctl->iter = nit;
            double mx = sdx[0];
            for (int k = 1; k < NS; k++) if (sdx[k] > mx) mx = sdx[k];
            int last = 0;
            if (mx < 0.001) { ctl->converged = 1; last = 1; }
            if (nit >= ctl->max_iter) last = 1;
            if (last) ctl->done = 1;
            s_last = last;
        }
    } else {
        for (int q = tid; q < NE; q += N_WORK) sJ[q] = (q / NS == q % NS) ? 1.0 : 0.0;     // L := I meanwhile
    }
    __syncthreads();
    tk5 = clock64();
    if (tid == 0 && it == 0) {      // phase cycles of the first iteration: reduce|boxminus, J blocks, products, LU inverse, boxplus
        ctl->dbg[3] = (int)(tk1 - tk0); ctl->dbg[4] = (int)(tk2 - tk1); ctl->dbg[5] = (int)(tk3 - tk2); ctl->dbg[6] = (int)(tk4 - tk3); ctl->dbg[7] = (int)(tk5 - tk4);
    }
    if (!s_last) return;
    // (6) P_ = L H_^-1 L^T with L from the final delta and the updated state (ieskf.cpp:151-155)
    if (wid == W_MAN && lane < 3) jac_blocks(sJ, sdx, sx + 33, sxp + 33, lane);
    __syncthreads();
    block_mm(sJ, sHinv, sC, false, false);
    __syncthreads();
    block_mm(sC, sJ, sB, false, true);
    __syncthreads();
    for (int q = tid; q < NE; q += SOLVE_THREADS) f->P[q] = sB[q];
}

void launch_solve(cudaStream_t st, bool ext, DevFilter* f, DevCtl* ctl, const double* partials, int nblocks) {
    if (ext) k_ieskf_solve<true><<<1, SOLVE_THREADS, 0, st>>>(f, ctl, partials, nblocks);
    else k_ieskf_solve<false><<<1, SOLVE_THREADS, 0, st>>>(f, ctl, partials, nblocks);
}
