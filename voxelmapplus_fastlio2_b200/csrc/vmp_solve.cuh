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

// ---- manifold pieces: each runs on ONE lane of its own warp, so the three of them proceed in parallel ----
// J or L block of a rotation (ieskf.cpp:136-137 / 151-152)
__device__ __noinline__ void rot_jac(const double* d3, double* jb9) {
    const M3 j = right_jacobian(v3(d3[0], d3[1], d3[2]));
    for (int a = 0; a < 9; a++) jb9[a] = j.a[a];
}
// J or L block of the S2 gravity (ieskf.cpp:138-139 / 153-154)
__device__ __noinline__ void g_jac(const double* g_cur, const double* g_pred, const double* d2, double* jb4) {
    Mat<2, 1> dg; dg[0] = d2[0]; dg[1] = d2[1];
    const Mat<2, 2> jg = mul(st_Nx(v3(g_cur[0], g_cur[1], g_cur[2])), st_Mx_res(v3(g_pred[0], g_pred[1], g_pred[2]), dg));
    jb4[0] = jg(0, 0); jb4[1] = jg(0, 1); jb4[2] = jg(1, 0); jb4[3] = jg(1, 1);
}
__device__ __forceinline__ void inv3_cof(const double* m, double* o) {
    const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    o[0] = c00 / det; o[1] = (m[2] * m[7] - m[1] * m[8]) / det; o[2] = (m[1] * m[5] - m[2] * m[4]) / det;
    o[3] = c01 / det; o[4] = (m[0] * m[8] - m[2] * m[6]) / det; o[5] = (m[2] * m[3] - m[0] * m[5]) / det;
    o[6] = c02 / det; o[7] = (m[1] * m[6] - m[0] * m[7]) / det; o[8] = (m[0] * m[4] - m[1] * m[3]) / det;
}
// delta = Log(Rp^T R) (ieskf.cpp:41-47), J = right jacobian, A = J^-1
__device__ __noinline__ void man_rot_piece(const double* R, const double* Rp, double* delta3, double* jb9, double* ab9) {
    M3 a, b;
    for (int k = 0; k < 9; k++) { a.a[k] = R[k]; b.a[k] = Rp[k]; }
    const V3 t = so3_log(mul(tr(b), a));
    delta3[0] = t[0]; delta3[1] = t[1]; delta3[2] = t[2];
    rot_jac(delta3, jb9);
    inv3_cof(jb9, ab9);
}
__device__ __noinline__ void man_g_piece(const double* g, const double* gp, double* delta2, double* jb4, double* ab4) {
    st_boxminus_g(v3(g[0], g[1], g[2]), v3(gp[0], gp[1], gp[2]), delta2);
    g_jac(g, gp, delta2, jb4);
    const double det = jb4[0] * jb4[3] - jb4[1] * jb4[2];
    ab4[0] = jb4[3] / det; ab4[1] = -jb4[1] / det; ab4[2] = -jb4[2] / det; ab4[3] = jb4[0] / det;
}
// rot <- rot Exp(d) (ieskf.cpp:14-15)
__device__ __noinline__ void plus_rot_piece(double* R, const double* d3) {
    M3 a;
    for (int k = 0; k < 9; k++) a.a[k] = R[k];
    const M3 r = mul(a, so3_exp(v3(d3[0], d3[1], d3[2])));
    for (int k = 0; k < 9; k++) R[k] = r.a[k];
}
// g <- Exp(Bx(g) d) g (ieskf.cpp:20)
__device__ __noinline__ void plus_g_piece(double* g, const double* d2) {
    const V3 gv = v3(g[0], g[1], g[2]);
    Mat<2, 1> dg; dg[0] = d2[0]; dg[1] = d2[1];
    const V3 r = mul(so3_exp(mul(st_Bx(gv), dg)), gv);
    g[0] = r[0]; g[1] = r[1]; g[2] = r[2];
}
// entry (i, j) of a block-diagonal 23x23 matrix given by its three blocks (identity elsewhere):
// blk[0..8] rows 3-5, blk[9..17] rows 6-8, blk[18..21] rows 21-22
__device__ __forceinline__ double bd_at(const double* blk, int i, int j) {
    if (i >= 3 && i < 6 && j >= 3 && j < 6) return blk[(i - 3) * 3 + (j - 3)];
    if (i >= 6 && i < 9 && j >= 6 && j < 9) return blk[9 + (i - 6) * 3 + (j - 6)];
    if (i >= 21 && j >= 21) return blk[18 + (i - 21) * 2 + (j - 21)];
    return i == j ? 1.0 : 0.0;
}
// column range of the non-zeros of row i of such a matrix
__device__ __forceinline__ void bd_range(int i, int& lo, int& hi) {
    lo = i < 3 ? i : i < 6 ? 3 : i < 9 ? 6 : i < 21 ? i : 21;
    hi = i < 3 ? i + 1 : i < 6 ? 6 : i < 9 ? 9 : i < 21 ? i + 1 : 23;
}
// time stamp that cannot be issued before the preceding barrier has really released the warp
__device__ __forceinline__ long long stamp(const volatile int* flag) {
    long long t = 0;
    if (*flag == 0) t = clock64();
    return t;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// flag written by a stream memory operation / DMA copy of the host's upload stream (system scope)
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// G = T^-1 for a small D x D matrix in shared memory by one warp: Gauss-Jordan with partial pivoting on the augmented
// matrix [T | I], lane j < 2D keeps column j in registers (all indices static after unrolling).  One sweep, no
// substitution phase; a lone warp pays per instruction and per dependent division, and this has D of each.
template <int D>
__device__ void warp_small_inverse(const double* T, double* G) {
    const int lane = threadIdx.x & 31;
    double c[D];
#pragma unroll
    for (int i = 0; i < D; i++) c[i] = lane < D ? T[i * D + lane] : (lane - D == i ? 1.0 : 0.0);
#pragma unroll
    for (int k = 0; k < D; k++) {
        int p = k;
        double best = fabs(c[k]);
#pragma unroll
        for (int i = k + 1; i < D; i++) { const double v = fabs(c[i]); if (v > best) { best = v; p = i; } }
        p = __shfl_sync(0xffffffffu, p, k);                      // the pivot row is decided by the lane that owns column k
#pragma unroll
        for (int i = k + 1; i < D; i++) if (i == p) { const double t = c[i]; c[i] = c[k]; c[k] = t; }
        const double piv = __shfl_sync(0xffffffffu, c[k], k);
        double fct[D];
#pragma unroll
        for (int i = 0; i < D; i++) fct[i] = __shfl_sync(0xffffffffu, c[i], k);   // column k before it is eliminated
        c[k] = c[k] / piv;
#pragma unroll
        for (int i = 0; i < D; i++) if (i != k) c[i] = c[i] - fct[i] * c[k];
    }
    if (lane >= D && lane < 2 * D) {
#pragma unroll
        for (int i = 0; i < D; i++) G[i * D + (lane - D)] = c[i];
    }
    __syncwarp();
}

// The solver CTA of k_measure (block 0 of a launch with solve = 1, vmp_iekf.cu): THREADS = that kernel's CTA size (>= 128).
//   (A) while the other CTAs measure: x, x_pred, P -> shared; boxminus, J, A = J^-1 (three warps, one manifold piece each)
//   (W) wait for the nblocks measurement CTAs (ctl->ticket), then reduce their partial sums in a fixed order
//   (B) the D x D algebra, delta_x, boxplus, convergence
//   (C) only on the last executed iteration: P = (L A)(P - Q P_D)(L A)^T
// shared memory of the solver CTA; k_measure overlays it with the staging of its measurement CTAs (a CTA is one or the other)
template <bool EXT>
struct SolveShared {
    static constexpr int D = EXT ? 12 : 6;
    static constexpr int NH = D * (D + 1) / 2;
    static constexpr int NV = NH + D + 1;
    static constexpr int NVP = (NV + 31) / 32 * 32;           // values padded to whole warps
    static constexpr int RP = EXT ? 4 : 8;                   // block subsets of the partial-sum reduction
    double sP[NE], sPn[NE], sT1[NE], sA[NE], sB[NE];
    double sHm[NV], sRed[RP][NVP];
    double sMA[D * D], sS[D * D], sTm[D * D], sG[D * D], sW[D * D];
    double sQ[NS * D], smm[D], sPm[D], sv[NS];
    double sdelta[NS], sdx[NS], sx[36], sxp[36], sJb[22], sAb[22], sLb[22], sBb[22];
    int s_last, s_zero;
};

// publish != 0: called from the resident loop kernel; the value is released to ctl->iter_pub once x_ and the loop flags are written.
// first != null: the first iteration of a scan.  The solver then also opens IESKF::update (ieskf.cpp:127-130: predict_x = x_, iteration
// counters) and stages the prior named by the scan header into the filter (round 1 did that in a kernel of its own, k_set_scan).
template <bool EXT, int THREADS>
__device__ void ieskf_solve_cta(SolveShared<EXT>& S, DevFilter* f, DevCtl* ctl, const double* partials, int nblocks, const ScanIn* first, unsigned long long publish) {
    static_assert(THREADS >= 128, "the manifold pieces use four warps");
    constexpr int D = EXT ? 12 : 6;
    constexpr int NH = D * (D + 1) / 2;
    constexpr int NV = NH + D + 1;
    constexpr int NVP = SolveShared<EXT>::NVP;
    constexpr int RP = SolveShared<EXT>::RP;
    auto& sP = S.sP; auto& sPn = S.sPn; auto& sT1 = S.sT1; auto& sA = S.sA; auto& sB = S.sB;
    auto& sHm = S.sHm; auto& sRed = S.sRed;
    auto& sMA = S.sMA; auto& sS = S.sS; auto& sTm = S.sTm; auto& sG = S.sG; auto& sW = S.sW;
    auto& sQ = S.sQ; auto& smm = S.smm; auto& sPm = S.sPm; auto& sv = S.sv;
    auto& sdelta = S.sdelta; auto& sdx = S.sdx; auto& sx = S.sx; auto& sxp = S.sxp;
    auto& sJb = S.sJb; auto& sAb = S.sAb; auto& sLb = S.sLb; auto& sBb = S.sBb;
    int& s_last = S.s_last; int& s_zero = S.s_zero;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int it = first ? 0 : ctl->iter;
    const long long t0 = clock64();

    // ---- (A)
    if (first) {
        const int mode = first->mode;
        const bool staged = (mode & (SCAN_STATE_HDR | SCAN_STATE_DEV)) != 0;
        const double* xs = (mode & SCAN_STATE_HDR) ? first->x : (mode & SCAN_STATE_DEV) ? first->prior : f->x;
        const double* Ps = (mode & SCAN_STATE_HDR) ? first->P : (mode & SCAN_STATE_DEV) ? first->prior + 36 : f->P;
        for (int q = tid; q < 36; q += THREADS) { const double v = xs[q]; sx[q] = v; sxp[q] = v; f->xpred[q] = v; }
        for (int q = tid; q < NE; q += THREADS) { const double v = Ps[q]; sP[q] = v; if (staged) f->P[q] = v; }
        if (tid == 0) { ctl->n = first->n; ctl->seq = first->seq; ctl->converged = 0; }
        if (tid < 8) ctl->effect[tid] = 0;
    } else {
        for (int q = tid; q < 36; q += THREADS) { sx[q] = f->x[q]; sxp[q] = f->xpred[q]; }
        for (int q = tid; q < NE; q += THREADS) sP[q] = f->P[q];
    }
    if (tid == 0) s_zero = 0;
    __syncthreads();
    if (lane == 0) {
        if (wid == 0) man_rot_piece(sx + 3, sxp + 3, sdelta + 3, sJb, sAb);
        else if (wid == 1) man_rot_piece(sx + 12, sxp + 12, sdelta + 6, sJb + 9, sAb + 9);
        else if (wid == 2) man_g_piece(sx + 33, sxp + 33, sdelta + 21, sJb + 18, sAb + 18);
    }
    if (wid == 3 && lane < 15) {                    // the vector-space coordinates (ieskf.cpp:39-53)
        const int g = lane / 3, c = lane % 3;
        const int dst = g == 0 ? c : 9 + (g - 1) * 3 + c, src = g == 0 ? c : 21 + (g - 1) * 3 + c;
        sdelta[dst] = sx[src] - sxp[src];
    }
    __syncthreads();
    for (int q = tid; q < NE; q += THREADS) sA[q] = bd_at(sAb, q / NS, q % NS);
    const long long tA = stamp(&s_zero);

    // ---- (W)
    if (tid == 0) {
        while (ld_acquire_u32(&ctl->ticket) < (unsigned)nblocks) __nanosleep(40);
        ctl->ticket = 0;
    }
    __syncthreads();
    const long long tW = stamp(&s_zero);
    // fixed-order reduction of the per-CTA partial sums: RP interleaved block subsets per value; a subset is summed in
    // block order.  All loads of a subset are in flight at once (RED_INFLIGHT x RP blocks per L2 round trip: two trips for a whole B200 grid;
    // the loop used to take one round trip per eight blocks, five of them on the path of every iteration)
    constexpr int RED_INFLIGHT = 20;
    for (int q = tid; q < NVP * RP; q += THREADS) {
        const int v = q % NVP, part = q / NVP;
        double t = 0.0;
        if (v < NV) {
            for (int b = part; b < nblocks; b += RED_INFLIGHT * RP) {
                double xs[RED_INFLIGHT];
#pragma unroll
                for (int u = 0; u < RED_INFLIGHT; u++) {
                    const int bb = b + u * RP;
                    xs[u] = bb < nblocks ? __ldcg(&partials[(size_t)bb * PARTIAL_STRIDE + v]) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < RED_INFLIGHT; u++) t += xs[u];
            }
        }
        sRed[part][v] = t;
    }
    __syncthreads();
    for (int v = tid; v < NV; v += THREADS) {
        double t = sRed[0][v];
#pragma unroll
        for (int c = 1; c < RP; c++) t += sRed[c][v];
        sHm[v] = t;
    }
    __syncthreads();

    // ---- (B)
    for (int q = tid; q < D * D + D; q += THREADS) {
        if (q < D * D) {                                    // MA = M A_D (M symmetric, from its upper triangle)
            const int i = q / D, j = q % D;
            double s2 = 0.0;
#pragma unroll
            for (int k = 0; k < D; k++) {
                const int a = i < k ? i : k, c = i < k ? k : i;
                s2 += sHm[a * D - a * (a - 1) / 2 + (c - a)] * sA[k * NS + j];
            }
            sMA[q] = s2;
        } else {                                            // mm = A_D^T m
            const int l = q - D * D;
            double s2 = 0.0;
#pragma unroll
            for (int k = 0; k < D; k++) s2 += sA[k * NS + l] * sHm[NH + k];
            smm[l] = s2;
        }
    }
    __syncthreads();
    for (int q = tid; q < D * D + D; q += THREADS) {
        if (q < D * D) {                                    // S = A_D^T (M A_D)
            const int i = q / D, j = q % D;
            double s2 = 0.0;
#pragma unroll
            for (int k = 0; k < D; k++) s2 += sA[k * NS + i] * sMA[k * D + j];
            sS[q] = s2;
        } else {                                            // Pm = P_DD mm
            const int l = q - D * D;
            double s2 = 0.0;
#pragma unroll
            for (int k = 0; k < D; k++) s2 += sP[l * NS + k] * smm[k];
            sPm[l] = s2;
        }
    }
    __syncthreads();
    for (int q = tid; q < D * D; q += THREADS) {            // T = I + P_DD S
        const int i = q / D, j = q % D;
        double s2 = (i == j) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < D; k++) s2 += sP[i * NS + k] * sS[k * D + j];
        sTm[q] = s2;
    }
    __syncthreads();
    if (wid == 0) warp_small_inverse<D>(sTm, sG);           // G = (I + P_DD S)^-1
    __syncthreads();
    for (int q = tid; q < D * D; q += THREADS) {            // W = S G
        const int i = q / D, j = q % D;
        double s2 = 0.0;
#pragma unroll
        for (int k = 0; k < D; k++) s2 += sS[i * D + k] * sG[k * D + j];
        sW[q] = s2;
    }
    __syncthreads();
    for (int q = tid; q < NS * D; q += THREADS) {           // Q = P_{:,D} W   (23 x D)
        const int i = q / D, j = q % D;
        double s2 = 0.0;
#pragma unroll
        for (int k = 0; k < D; k++) s2 += sP[i * NS + k] * sW[k * D + j];
        sQ[q] = s2;
    }
    __syncthreads();
    if (tid < NS) {                                         // y = (delta - Q delta_D) + (P_{:,D} mm - Q Pm)
        double q1 = 0.0, q2 = 0.0, pm = 0.0;
#pragma unroll
        for (int k = 0; k < D; k++) { q1 += sQ[tid * D + k] * sdelta[k]; q2 += sQ[tid * D + k] * sPm[k]; pm += sP[tid * NS + k] * smm[k]; }
        sv[tid] = (sdelta[tid] - q1) + (pm - q2);
    }
    __syncthreads();
    if (tid < NS) {                                         // delta_x = -A y
        int lo, hi;
        bd_range(tid, lo, hi);
        double s2 = 0.0;
        for (int k = lo; k < hi; k++) s2 += sA[tid * NS + k] * sv[k];
        sdx[tid] = -s2;
    }
    __syncthreads();
    const long long tD = stamp(&s_zero);
    // x_ += delta_x (ieskf.cpp:11-21), counters, convergence on the signed maximum (Q5)
    if (lane == 0) {
        if (wid == 0) plus_rot_piece(sx + 3, sdx + 3);
        else if (wid == 1) plus_rot_piece(sx + 12, sdx + 6);
        else if (wid == 2) plus_g_piece(sx + 33, sdx + 21);
    }
    if (wid == 3) {
        if (lane < 15) {
            const int g = lane / 3, c = lane % 3;
            const int src = g == 0 ? c : 9 + (g - 1) * 3 + c, dst = g == 0 ? c : 21 + (g - 1) * 3 + c;
            sx[dst] = sx[dst] + sdx[src];
        } else if (lane == 15) {
            ctl->effect[it & 7] = (int)sHm[NH + D];
            const int nit = it + 1;
            ctl->iter = nit;
            double mx = sdx[0];
            for (int k = 1; k < NS; k++) if (sdx[k] > mx) mx = sdx[k];
            int last = 0;
            if (mx < 0.001) { ctl->converged = 1; last = 1; }
            if (nit >= ctl->max_iter) last = 1;
            ctl->done = last;                                   // (the first iteration of a scan clears the previous scan's flag)
            s_last = last;
        }
    }
    __syncthreads();
    const long long tX = stamp(&s_zero);
    if (tid < 36) f->x[tid] = sx[tid];
    if (publish) {      // k_iekf_loop: hand the new state (and ctl->done) to the measurement CTAs spinning on iter_pub
        __syncthreads();
        if (tid == 0) { __threadfence(); st_release_u64(&ctl->iter_pub, publish * 2ull + (s_last ? 1ull : 0ull)); }      // (the loop's stop flag rides in bit 0)
    }
    long long tC = tX;
    if (s_last) {
        // ---- (C) L blocks from the final delta_x and the updated state (ieskf.cpp:151-154), B = L A,
        //          P_ = B (P - Q P_{D,:}) B^T
        if (lane == 0 && wid < 3) {
            if (wid < 2) {
                rot_jac(sdx + 3 + 3 * wid, sLb + 9 * wid);
                for (int e = 0; e < 9; e++) {
                    const int i = e / 3, j = e % 3;
                    double s2 = 0.0;
                    for (int k = 0; k < 3; k++) s2 += sLb[wid * 9 + i * 3 + k] * sAb[wid * 9 + k * 3 + j];
                    sBb[wid * 9 + e] = s2;
                }
            } else {
                g_jac(sx + 33, sxp + 33, sdx + 21, sLb + 18);
                for (int e = 0; e < 4; e++) {
                    const int i = e / 2, j = e % 2;
                    sBb[18 + e] = sLb[18 + i * 2] * sAb[18 + j] + sLb[18 + i * 2 + 1] * sAb[20 + j];
                }
            }
        }
        for (int q = tid; q < NE; q += THREADS) {            // Pn = P - Q P_{D,:}
            const int i = q / NS, j = q % NS;
            double s2 = 0.0;
#pragma unroll
            for (int k = 0; k < D; k++) s2 += sQ[i * D + k] * sP[k * NS + j];
            sPn[q] = sP[q] - s2;
        }
        __syncthreads();
        for (int q = tid; q < NE; q += THREADS) sB[q] = bd_at(sBb, q / NS, q % NS);
        __syncthreads();
        for (int q = tid; q < NE; q += THREADS) {            // T1 = B Pn
            const int i = q / NS, j = q % NS;
            int lo, hi;
            bd_range(i, lo, hi);
            double s2 = 0.0;
            for (int k = lo; k < hi; k++) s2 += sB[i * NS + k] * sPn[k * NS + j];
            sT1[q] = s2;
        }
        __syncthreads();
        for (int q = tid; q < NE; q += THREADS) {            // P = T1 B^T
            const int i = q / NS, j = q % NS;
            int lo, hi;
            bd_range(j, lo, hi);
            double s2 = 0.0;
            for (int k = lo; k < hi; k++) s2 += sT1[i * NS + k] * sB[j * NS + k];
            f->P[q] = s2;
        }
        tC = clock64();
    }
    if (tid == 0) {      // phase cycles (debug counters): first iteration, and the posterior of the last one
        if (it == ctl->dbg_it) { ctl->dbg[3] = (int)(tA - t0); ctl->dbg[4] = (int)(tW - tA); ctl->dbg[5] = (int)(tD - tW); ctl->dbg[6] = (int)(tX - tD); }
        if (s_last) ctl->dbg[7] = (int)(tC - tX);
    }
}
