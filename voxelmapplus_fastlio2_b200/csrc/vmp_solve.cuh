// vmp_solve.cuh — the 23-dof algebra of IESKF::update (ieskf.cpp:125-156) on the device.
// Included by vmp_iekf.cu inside namespace vmp.
//
//   k_update_begin   predict_x = x_ (ieskf.cpp:127), iteration counters, and P_^-1: P_ does not change
//                    inside update(), so the inverse the reference evaluates twice per iteration (Q6) is
//                    computed once per scan here
//   k_ieskf_solve    one CTA per iteration: fixed-order reduction of the measurement partials, boxminus,
//                    J, H_ = J^T P^-1 J + H, b_, delta = -H_^-1 b_, boxplus, convergence flag; on the
//                    last executed iteration also P_ = L H_^-1 L^T
//
// The 23x23 inverse is LU with partial pivoting + substitution against the permuted identity
// (what Eigen's PartialPivLU-based inverse() does), by ONE warp: lane r keeps matrix row r in
// registers, the pivot search and the pivot-row broadcast go through shuffles.  To index the
// current column with a compile-time register index inside a ROLLED loop (the code runs once
// per launch: its size, i.e. instruction fetch, matters as much as its arithmetic) the row is
// rotated left by one element per elimination step, so the pivot column is always a[0]; the
// substitutions use the same trick on the right-hand-side columns (lane c = column c).  The
// per-entry operation order equals the serial lu_inverse<> in vmp_math.cuh, bit for bit.
#pragma once

constexpr int NS = 23;

// A: NS*NS input (shared), inv: NS*NS output (shared), w1/w2: NS*NS scratch each (shared), perm: NS ints.
// A may alias w1.  Whole warp must call.
__device__ void warp_lu_inverse(const double* A, double* inv, double* w1, double* w2, int* perm) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const bool act = lane < NS;
    double a[NS];
#pragma unroll
    for (int j = 0; j < NS; j++) a[j] = act ? A[lane * NS + j] : 0.0;
    __syncwarp();
    double* Lm = w1;            // multipliers by PHYSICAL row: Lm[r][k]
    double* U = w2;             // U[k][j], j >= k, by pivoted row
    int pos = act ? lane : 1000 + lane;                              // index of this lane's row in the pivoted order
#pragma unroll 1
    for (int k = 0; k < NS; k++) {
        // pivot: first row (in pivoted order) of maximal |a_ik|, i >= k
        double v = (act && pos >= k) ? fabs(a[0]) : -1.0;
        int p = pos;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double vo = __shfl_xor_sync(FULL, v, o);
            const int po = __shfl_xor_sync(FULL, p, o);
            if (vo > v || (vo == v && po < p)) { v = vo; p = po; }
        }
        if (pos == p) pos = k; else if (pos == k) pos = p;          // row swap k <-> p
        const int L = __ffs(__ballot_sync(FULL, pos == k)) - 1;
        const double d = __shfl_sync(FULL, a[0], L);
        const bool below = act && pos > k;
        const double l = a[0] / d;
        if (below) Lm[lane * NS + k] = l;
        if (lane == L) U[k * NS + k] = a[0];
#pragma unroll
        for (int j = 1; j < NS; j++) {
            const double u = __shfl_sync(FULL, a[j], L);
            if (lane == L && k + j < NS) U[k * NS + k + j] = a[j];
            a[j - 1] = below ? a[j] - l * u : a[j];                  // eliminate and rotate left
        }
        a[NS - 1] = 0.0;
    }
    if (act) perm[pos] = lane;                                       // pivoted row i is physical row perm[i]
    __syncwarp();
    if (act) {
        // forward substitution L y = P e_c (column c = lane), rotating right-hand side
        double s[NS];
#pragma unroll
        for (int i = 0; i < NS; i++) s[i] = (perm[i] == lane) ? 1.0 : 0.0;
#pragma unroll 1
        for (int j = 0; j < NS; j++) {
            const double y = s[0];
            inv[j * NS + lane] = y;
#pragma unroll
            for (int i = 1; i < NS; i++) {
                const int r = j + i;
                const double lv = (r < NS) ? Lm[perm[r < NS ? r : 0] * NS + j] : 0.0;
                s[i - 1] = s[i] - lv * y;
            }
            s[NS - 1] = 0.0;
        }
        // backward substitution U x = y, from the last unknown, rotating the same way
#pragma unroll
        for (int i = 0; i < NS; i++) s[i] = inv[(NS - 1 - i) * NS + lane];
#pragma unroll 1
        for (int q = 0; q < NS; q++) {
            const int i0 = NS - 1 - q;
            const double x = s[0] / U[i0 * NS + i0];
            inv[i0 * NS + lane] = x;
#pragma unroll
            for (int i = 1; i < NS; i++) {
                const int r = i0 - i;
                const double uv = (r >= 0) ? U[(r >= 0 ? r : 0) * NS + i0] : 0.0;
                s[i - 1] = s[i] - uv * x;
            }
            s[NS - 1] = 0.0;
        }
    }
    __syncwarp();
}

// C = op(A) * op(B), NS x NS in shared memory, entries strided over the block, left-to-right sums
__device__ __forceinline__ void block_mm(const double* A, const double* B, double* C, bool transA, bool transB) {
    const int sa = transA ? NS : 1, sb = transB ? 1 : NS;            // strides along k
    for (int q = threadIdx.x; q < NS * NS; q += blockDim.x) {
        const int i = q / NS, j = q % NS;
        const double* a = A + (transA ? i : i * NS);
        const double* b = B + (transB ? j * NS : j);
        double s = a[0] * b[0];
#pragma unroll 2
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
__device__ __forceinline__ void block_identity(double* J) {
    for (int q = threadIdx.x; q < NS * NS; q += blockDim.x) J[q] = (q / NS == q % NS) ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(64) k_update_begin(DevFilter* f, DevCtl* ctl) {
    __shared__ double sA[NS * NS], sInv[NS * NS], sW[NS * NS];
    __shared__ int perm[NS];
    const int tid = threadIdx.x;
    if (tid >= 32) {
        const int t = tid - 32;
        f->xpred[t] = f->x[t];
        if (t < 4) f->xpred[32 + t] = f->x[32 + t];
        if (t == 0) { ctl->iter = 0; ctl->done = 0; ctl->converged = 0; }
        if (t < 8) ctl->effect[t] = 0;
        return;
    }
    for (int q = tid; q < NS * NS; q += 32) sA[q] = f->P[q];
    __syncwarp();
    warp_lu_inverse(sA, sInv, sA, sW, perm);
    for (int q = tid; q < NS * NS; q += 32) f->Pinv[q] = sInv[q];
}
void launch_update_begin(cudaStream_t st, DevFilter* f, DevCtl* ctl) { k_update_begin<<<1, 64, 0, st>>>(f, ctl); }

constexpr int RED_CHUNKS = 8;

template <bool EXT>
__global__ void __launch_bounds__(256) k_ieskf_solve(DevFilter* f, DevCtl* ctl, const double* __restrict__ partials, int nblocks) {
    constexpr int D = EXT ? 12 : 6;
    constexpr int NH = D * (D + 1) / 2;
    constexpr int NV = NH + D + 1;
    __shared__ double sA[NS * NS], sB[NS * NS], sC[NS * NS], sJ[NS * NS], sHinv[NS * NS];
    __shared__ double sHm[NV], sRed[RED_CHUNKS][NV], sdelta[NS], sb[NS], sdx[NS], sx[36], sxp[36];
    __shared__ int perm[NS], s_last;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    if (ctl->done) return;
    const int it = ctl->iter;

    // (1) concurrently: reduction of the per-block partials (fixed order: RED_CHUNKS contiguous block
    //     ranges, then the chunk sums in ascending order) | boxminus | P^-1 load | J := I
    if (wid == 7) {
        for (int q = lane; q < 36; q += 32) { sx[q] = f->x[q]; sxp[q] = f->xpred[q]; }
        __syncwarp();
        if (lane == 0) { const St x = st_load(sx), xp = st_load(sxp); st_boxminus(x, xp, sdelta); }
    } else {
        for (int q = tid; q < NV * RED_CHUNKS; q += 224) {
            const int v = q % NV, c = q / NV;
            const int b0 = (int)((long long)nblocks * c / RED_CHUNKS), b1 = (int)((long long)nblocks * (c + 1) / RED_CHUNKS);
            double t = 0.0;
#pragma unroll 4
            for (int b = b0; b < b1; b++) t += partials[(size_t)b * PARTIAL_STRIDE + v];
            sRed[c][v] = t;
        }
        for (int q = tid; q < NS * NS; q += 224) { sB[q] = f->Pinv[q]; sJ[q] = (q / NS == q % NS) ? 1.0 : 0.0; }
    }
    __syncthreads();
    if (tid < NV) {
        double t = sRed[0][tid];
#pragma unroll
        for (int c = 1; c < RED_CHUNKS; c++) t += sRed[c][tid];
        sHm[tid] = t;
    }
    // (2) the blocks of J (ieskf.cpp:136-139)
    if (wid == 7 && lane < 3) jac_blocks(sJ, sdelta, sx + 33, sxp + 33, lane);
    __syncthreads();
    // (3) JtPinv = J^T P^-1 ; b_ = JtPinv delta ; H_ = JtPinv J (+ measurement H, b in the top-left corner)
    block_mm(sJ, sB, sC, true, false);
    __syncthreads();
    if (tid < NS) {
        double t = sC[tid * NS] * sdelta[0];
        for (int k = 1; k < NS; k++) t += sC[tid * NS + k] * sdelta[k];
        t = 0.0 + t;
        if (tid < D) t += sHm[NH + tid];
        sb[tid] = t;
    }
    block_mm(sC, sJ, sA, false, false);
    __syncthreads();
    for (int q = tid; q < NS * NS; q += blockDim.x) {
        const int i = q / NS, j = q % NS;
        double h = 0.0 + sA[q];
        if (i < D && j < D) { const int a = i < j ? i : j, c = i < j ? j : i; h += sHm[a * D - a * (a - 1) / 2 + (c - a)]; }
        sA[q] = h;
    }
    __syncthreads();
    // (4) H_^-1 (one warp), delta = -H_^-1 b_
    if (wid == 0) {
        warp_lu_inverse(sA, sHinv, sB, sC, perm);
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
        // L := I while warp 0 inverts (sJ is free: H_ is formed)
        for (int q = tid - 32; q < NS * NS; q += 224) sJ[q] = (q / NS == q % NS) ? 1.0 : 0.0;
    }
    __syncthreads();
    if (!s_last) return;
    // (6) P_ = L H_^-1 L^T with L from the final delta and the updated state (ieskf.cpp:151-155)
    if (wid == 7 && lane < 3) jac_blocks(sJ, sdx, sx + 33, sxp + 33, lane);
    __syncthreads();
    block_mm(sJ, sHinv, sC, false, false);
    __syncthreads();
    block_mm(sC, sJ, sB, false, true);
    __syncthreads();
    for (int q = tid; q < NS * NS; q += blockDim.x) f->P[q] = sB[q];
}

void launch_solve(cudaStream_t st, bool ext, DevFilter* f, DevCtl* ctl, const double* partials, int nblocks) {
    if (ext) k_ieskf_solve<true><<<1, 256, 0, st>>>(f, ctl, partials, nblocks);
    else k_ieskf_solve<false><<<1, 256, 0, st>>>(f, ctl, partials, nblocks);
}
