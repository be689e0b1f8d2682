// vmp_solve.cuh — the 23-dof algebra of IESKF::update (ieskf.cpp:125-156) on the device.
// Included by vmp_iekf.cu inside namespace vmp.
//
// The reference forms H_ = J^T P^-1 J + [H 0; 0 0] (23x23), b_ = J^T P^-1 delta + [b; 0], inverts P (twice) and H_
// every iteration and, at the end, sets P = L H_^-1 L^T.  The measurement information H only touches the first
// D = 6 (12 with estimate_ext) error-state coordinates, so with A = J^-1 (block diagonal), S = A_D^T H A_D,
// mm = A_D^T b and the matrix-inversion lemma
//        (P^-1 + U S U^T)^-1 = P - P U S (I_D + P_DD S)^-1 U^T P ,      U = [I_D; 0]
// the same quantities are obtained from ONE D x D inverse and a few 23 x D products, with no P^-1 at all:
//        G = (I_D + P_DD S)^-1 ,  Q = P_{:,D} S G
//        delta_x = -A [ (delta - Q delta_D) + (P_{:,D} mm - Q P_DD mm) ]
//        P_post  = (L A) (P - Q P_{D,:}) (L A)^T
// This is algebraically identical to ieskf.cpp:134-155 and differs only in rounding (measured against the
// oracle by tests/test_gpu_parity.py: posterior position < 1e-9 m, P to 2e-5 relative, i.e. the conditioning of either form).  The first version of this
// kernel restated the reference's three 23x23 LU inverses literally; ncu showed them to be a chain of dependent
// fp64 / barrier latencies (27 us per inverse) that no amount of threads shortens.
#pragma once

constexpr int NS = 23;
constexpr int NE = NS * NS;

// the three non-identity blocks of J / L (ieskf.cpp:136-139, 151-154): jb[0..8] rows 3-5, jb[9..17] rows 6-8, jb[18..21] rows 21-22
__device__ __noinline__ void jac_blocks(double* jb, const double* delta, const double* g_cur, const double* g_pred, int which) {
    if (which == 0 || which == 1) {
        const int o = which == 0 ? 3 : 6;
        const M3 j = right_jacobian(v3(delta[o], delta[o + 1], delta[o + 2]));
        for (int a = 0; a < 9; a++) jb[which * 9 + a] = j.a[a];
    } else {
        Mat<2, 1> dg; dg[0] = delta[21]; dg[1] = delta[22];
        const Mat<2, 2> jg = mul(st_Nx(v3(g_cur[0], g_cur[1], g_cur[2])), st_Mx_res(v3(g_pred[0], g_pred[1], g_pred[2]), dg));
        jb[18] = jg(0, 0); jb[19] = jg(0, 1); jb[20] = jg(1, 0); jb[21] = jg(1, 1);
    }
}
// inverses of those blocks (3x3 by cofactors, 2x2 closed form)
__device__ __forceinline__ void inv_blocks(const double* jb, double* ab, int which) {
    if (which < 2) {
        const double* m = jb + which * 9;
        double* o = ab + which * 9;
        const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
        const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
        o[0] = c00 / det; o[1] = (m[2] * m[7] - m[1] * m[8]) / det; o[2] = (m[1] * m[5] - m[2] * m[4]) / det;
        o[3] = c01 / det; o[4] = (m[0] * m[8] - m[2] * m[6]) / det; o[5] = (m[2] * m[3] - m[0] * m[5]) / det;
        o[6] = c02 / det; o[7] = (m[1] * m[6] - m[0] * m[7]) / det; o[8] = (m[0] * m[4] - m[1] * m[3]) / det;
    } else {
        const double det = jb[18] * jb[21] - jb[19] * jb[20];
        ab[18] = jb[21] / det; ab[19] = -jb[19] / det; ab[20] = -jb[20] / det; ab[21] = jb[18] / det;
    }
}
// entry (i, j) of a block-diagonal 23x23 matrix given by its three blocks (identity elsewhere)
__device__ __forceinline__ double bd_at(const double* blk, int i, int j) {
    if (i >= 3 && i < 6 && j >= 3 && j < 6) return blk[(i - 3) * 3 + (j - 3)];
    if (i >= 6 && i < 9 && j >= 6 && j < 9) return blk[9 + (i - 6) * 3 + (j - 6)];
    if (i >= 21 && j >= 21) return blk[18 + (i - 21) * 2 + (j - 21)];
    return i == j ? 1.0 : 0.0;
}

// G = T^-1 for a small D x D matrix in shared memory (LU with partial pivoting + substitution), by one warp:
// lane i owns row i during the elimination and column i of G during the substitution.  T is destroyed.
template <int D>
__device__ void warp_small_inverse(double* T, double* G, int* perm) {
    const int lane = threadIdx.x & 31;
    if (lane < D) perm[lane] = lane;
    __syncwarp();
    for (int k = 0; k < D; k++) {
        int piv = k; double best = fabs(T[k * D + k]);           // every lane finds the same pivot
        for (int i = k + 1; i < D; i++) { const double v = fabs(T[i * D + k]); if (v > best) { best = v; piv = i; } }
        __syncwarp();
        if (piv != k) {
            if (lane < D) { const double t = T[k * D + lane]; T[k * D + lane] = T[piv * D + lane]; T[piv * D + lane] = t; }
            if (lane == 0) { const int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t; }
            __syncwarp();
        }
        if (lane > k && lane < D) {
            const double l = T[lane * D + k] / T[k * D + k];
            T[lane * D + k] = l;
            for (int j = k + 1; j < D; j++) T[lane * D + j] -= l * T[k * D + j];
        }
        __syncwarp();
    }
    if (lane < D) {
        double y[D];
#pragma unroll
        for (int i = 0; i < D; i++) {
            double sacc = (perm[i] == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int j = 0; j < i; j++) sacc -= T[i * D + j] * y[j];
            y[i] = sacc;
        }
#pragma unroll
        for (int i = D - 1; i >= 0; i--) {
            double sacc = y[i];
#pragma unroll
            for (int j = i + 1; j < D; j++) sacc -= T[i * D + j] * y[j];
            y[i] = sacc / T[i * D + i];
        }
#pragma unroll
        for (int i = 0; i < D; i++) G[i * D + lane] = y[i];
    }
    __syncwarp();
}

// kept as a no-op stage id for the profiler; P^-1 is not needed any more
__global__ void k_update_begin(DevFilter* f, DevCtl* ctl) { (void)f; (void)ctl; }
void launch_update_begin(cudaStream_t st, DevFilter* f, DevCtl* ctl) { (void)st; (void)f; (void)ctl; }

constexpr int RED_CHUNKS = 2;

// Phases: (A) all warps: partial sums of the measurement blocks + P load | manifold warp: boxminus, J, A = J^-1
//         (B) manifold warp alone, warp-synchronous: every D x D quantity, Q, delta_x, boxplus, convergence, L A
//         (C) only on the last executed iteration, all warps: P = (L A)(P - Q P_D)(L A)^T
// Runs inside k_measure's LAST-arriving CTA (vmp_iekf.cu): THREADS = that kernel's CTA size.
template <bool EXT, int THREADS>
__device__ void ieskf_solve_block(DevFilter* f, DevCtl* ctl, const double* partials, int nblocks) {
    constexpr int SOLVE_THREADS = THREADS;
    constexpr int W_MAN = THREADS / 32 - 1;        // the warp that runs the manifold operations
    constexpr int N_WORK = 32 * W_MAN;             // threads of the other warps
    constexpr int D = EXT ? 12 : 6;
    constexpr int NH = D * (D + 1) / 2;
    constexpr int NV = NH + D + 1;
    __shared__ double sP[NE], sPn[NE], sT1[NE];
    __shared__ double sHm[NV], sRed[RED_CHUNKS][NV];
    __shared__ double sM[D * D], sS[D * D], sTm[D * D], sG[D * D], sW[D * D], sMA[D * D];
    __shared__ double sQ[NS * D], smm[D], sPm[D], sv[NS];
    __shared__ double sdelta[NS], sdx[NS], sx[36], sxp[36], sJb[22], sAb[22], sLb[22], sBb[22];
    __shared__ int sperm[D], s_last;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int it = ctl->iter;
    const long long tk0 = clock64();
    long long tk1 = 0, tk2 = 0, tk3 = 0;

    // ---- (A)
    if (wid == W_MAN) {
        for (int q = lane; q < 36; q += 32) { sx[q] = f->x[q]; sxp[q] = f->xpred[q]; }
        __syncwarp();
        if (lane == 0) { const St x = st_load(sx), xp = st_load(sxp); st_boxminus(x, xp, sdelta); }
        __syncwarp();
        if (lane < 3) { jac_blocks(sJb, sdelta, sx + 33, sxp + 33, lane); inv_blocks(sJb, sAb, lane); }
    } else {
        for (int q = tid; q < NV * RED_CHUNKS; q += N_WORK) {
            const int v = q % NV, c = q / NV;
            const int b0 = (int)((long long)nblocks * c / RED_CHUNKS), b1 = (int)((long long)nblocks * (c + 1) / RED_CHUNKS);
            double t = 0.0;
#pragma unroll 8
            for (int b = b0; b < b1; b++) t += __ldcg(&partials[(size_t)b * PARTIAL_STRIDE + v]);
            sRed[c][v] = t;
        }
        for (int q = tid; q < NE; q += N_WORK) sP[q] = f->P[q];
    }
    __syncthreads();
    tk1 = clock64();
    // ---- (B) one warp, no block barriers
    if (wid == W_MAN) {
        for (int v = lane; v < NV; v += 32) {
            double t = sRed[0][v];
#pragma unroll
            for (int c = 1; c < RED_CHUNKS; c++) t += sRed[c][v];
            sHm[v] = t;
        }
        __syncwarp();
        for (int q = lane; q < D * D; q += 32) {             // M (symmetric, from the upper triangle)
            const int i = q / D, j = q % D, a = i < j ? i : j, c = i < j ? j : i;
            sM[q] = sHm[a * D - a * (a - 1) / 2 + (c - a)];
        }
        __syncwarp();
        for (int q = lane; q < D * D; q += 32) {             // MA = M A_D
            const int i = q / D, j = q % D;
            double s2 = 0.0;
            for (int k = 0; k < D; k++) s2 += sM[i * D + k] * bd_at(sAb, k, j);
            sMA[q] = s2;
        }
        if (lane < D) {                                      // mm = A_D^T m
            double s2 = 0.0;
            for (int k = 0; k < D; k++) s2 += bd_at(sAb, k, lane) * sHm[NH + k];
            smm[lane] = s2;
        }
        __syncwarp();
        for (int q = lane; q < D * D; q += 32) {             // S = A_D^T (M A_D)
            const int i = q / D, j = q % D;
            double s2 = 0.0;
            for (int k = 0; k < D; k++) s2 += bd_at(sAb, k, i) * sMA[k * D + j];
            sS[q] = s2;
        }
        if (lane < D) {                                      // Pm = P_DD mm
            double s2 = 0.0;
            for (int k = 0; k < D; k++) s2 += sP[lane * NS + k] * smm[k];
            sPm[lane] = s2;
        }
        __syncwarp();
        for (int q = lane; q < D * D; q += 32) {             // T = I + P_DD S
            const int i = q / D, j = q % D;
            double s2 = (i == j) ? 1.0 : 0.0;
            for (int k = 0; k < D; k++) s2 += sP[i * NS + k] * sS[k * D + j];
            sTm[q] = s2;
        }
        __syncwarp();
        warp_small_inverse<D>(sTm, sG, sperm);               // G = (I + P_DD S)^-1
        for (int q = lane; q < D * D; q += 32) {             // W = S G
            const int i = q / D, j = q % D;
            double s2 = 0.0;
            for (int k = 0; k < D; k++) s2 += sS[i * D + k] * sG[k * D + j];
            sW[q] = s2;
        }
        __syncwarp();
        for (int q = lane; q < NS * D; q += 32) {            // Q = P_{:,D} W   (23 x D)
            const int i = q / D, j = q % D;
            double s2 = 0.0;
            for (int k = 0; k < D; k++) s2 += sP[i * NS + k] * sW[k * D + j];
            sQ[q] = s2;
        }
        __syncwarp();
        if (lane < NS) {                                     // y = (delta - Q delta_D) + (P_{:,D} mm - Q Pm)
            double q1 = 0.0, q2 = 0.0, pm = 0.0;
            for (int k = 0; k < D; k++) { q1 += sQ[lane * D + k] * sdelta[k]; q2 += sQ[lane * D + k] * sPm[k]; pm += sP[lane * NS + k] * smm[k]; }
            sv[lane] = (sdelta[lane] - q1) + (pm - q2);
        }
        __syncwarp();
        if (lane < NS) {                                     // delta_x = -A y
            const int lo = lane < 3 ? lane : lane < 6 ? 3 : lane < 9 ? 6 : lane < 21 ? lane : 21;
            const int hi = lane < 3 ? lane + 1 : lane < 6 ? 6 : lane < 9 ? 9 : lane < 21 ? lane + 1 : 23;
            double s2 = 0.0;
            for (int k = lo; k < hi; k++) s2 += bd_at(sAb, lane, k) * sv[k];
            sdx[lane] = -s2;
        }
        __syncwarp();
        // x_ += delta, counters, convergence on the signed maximum (Q5)
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
        __syncwarp();
        // L blocks from the final delta and the updated state (ieskf.cpp:151-154), B = L A
        if (s_last && lane < 3) jac_blocks(sLb, sdx, sx + 33, sxp + 33, lane);
        __syncwarp();
        if (s_last && lane < 22) {
            if (lane < 18) {
                const int w = lane / 9, e = lane % 9, i = e / 3, j = e % 3;
                double s2 = 0.0;
                for (int k = 0; k < 3; k++) s2 += sLb[w * 9 + i * 3 + k] * sAb[w * 9 + k * 3 + j];
                sBb[lane] = s2;
            } else {
                const int e = lane - 18, i = e / 2, j = e % 2;
                sBb[lane] = sLb[18 + i * 2] * sAb[18 + j] + sLb[18 + i * 2 + 1] * sAb[20 + j];
            }
        }
    }
    __syncthreads();
    tk2 = clock64();
    if (s_last) {
        // ---- (C) P_ = B (P - Q P_{D,:}) B^T with B = L A block diagonal
        for (int q = tid; q < NE; q += SOLVE_THREADS) {
            const int i = q / NS, j = q % NS;
            double s2 = 0.0;
#pragma unroll
            for (int k = 0; k < D; k++) s2 += sQ[i * D + k] * sP[k * NS + j];
            sPn[q] = sP[q] - s2;
        }
        __syncthreads();
        for (int q = tid; q < NE; q += SOLVE_THREADS) {      // T1 = B Pn
            const int i = q / NS, j = q % NS;
            const int lo = i < 3 ? i : i < 6 ? 3 : i < 9 ? 6 : i < 21 ? i : 21;
            const int hi = i < 3 ? i + 1 : i < 6 ? 6 : i < 9 ? 9 : i < 21 ? i + 1 : 23;
            double s2 = 0.0;
            for (int k = lo; k < hi; k++) s2 += bd_at(sBb, i, k) * sPn[k * NS + j];
            sT1[q] = s2;
        }
        __syncthreads();
        for (int q = tid; q < NE; q += SOLVE_THREADS) {      // P = T1 B^T
            const int i = q / NS, j = q % NS;
            const int lo = j < 3 ? j : j < 6 ? 3 : j < 9 ? 6 : j < 21 ? j : 21;
            const int hi = j < 3 ? j + 1 : j < 6 ? 6 : j < 9 ? 9 : j < 21 ? j + 1 : 23;
            double s2 = 0.0;
            for (int k = lo; k < hi; k++) s2 += sT1[i * NS + k] * bd_at(sBb, j, k);
            f->P[q] = s2;
        }
    }
    tk3 = clock64();
    if (tid == 0 && it == 0) {      // phase cycles of the first iteration (debug counters): A, B, C
        ctl->dbg[3] = (int)(tk1 - tk0); ctl->dbg[4] = (int)(tk2 - tk1); ctl->dbg[5] = (int)(tk3 - tk2); ctl->dbg[6] = 0; ctl->dbg[7] = 0;
    }
}

