// vmp_math.cuh — fixed-size fp64 algebra shared by the device kernels and the host-side
// LIOBuilder of the B200 path.  Everything is __host__ __device__ and written so that
// host (g++/nvcc host pass, -ffp-contract=off) and device (nvcc -fmad=false) execute
// the same sequence of IEEE operations: products are accumulated left to right, the
// first product seeds the sum, no fused multiply-add anywhere.  That fixed evaluation
// order is what makes voxel keys, plane fits and gate decisions reproducible bit for
// bit against the CPU oracle (tests/), which restates the reference's Eigen/Sophus calls:
//   SelfAdjointEigenSolver<Matrix3d>   voxel_map.cpp:104-106      -> eig3_sym
//   Matrix23d::inverse()               ieskf.cpp:141,142,145,155  -> lu_inverse (host), k_ieskf_solve (device)
//   Sophus::SO3d exp/log/hat/leftJac   ieskf.cpp:8-40,89          -> so3_*
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define VMP_HD __host__ __device__ __forceinline__
#else
#define VMP_HD inline
#endif

namespace vmp {

template <int R, int C>
struct Mat {
    double a[R * C];
    VMP_HD double& operator()(int i, int j) { return a[i * C + j]; }
    VMP_HD const double& operator()(int i, int j) const { return a[i * C + j]; }
    VMP_HD double& operator[](int i) { return a[i]; }
    VMP_HD const double& operator[](int i) const { return a[i]; }
};
typedef Mat<3, 1> V3;
typedef Mat<3, 3> M3;
typedef Mat<6, 6> M6;

template <int R, int C>
VMP_HD Mat<R, C> zeros() { Mat<R, C> m; for (int i = 0; i < R * C; i++) m.a[i] = 0.0; return m; }
template <int N>
VMP_HD Mat<N, N> eye() { Mat<N, N> m = zeros<N, N>(); for (int i = 0; i < N; i++) m(i, i) = 1.0; return m; }
VMP_HD V3 v3(double x, double y, double z) { V3 v; v[0] = x; v[1] = y; v[2] = z; return v; }

template <int R, int K, int C>
VMP_HD Mat<R, C> mul(const Mat<R, K>& A, const Mat<K, C>& B) {
    Mat<R, C> o;
#pragma unroll
    for (int i = 0; i < R; i++)
#pragma unroll
        for (int j = 0; j < C; j++) {
            double s = A(i, 0) * B(0, j);
#pragma unroll
            for (int k = 1; k < K; k++) s += A(i, k) * B(k, j);
            o(i, j) = s;
        }
    return o;
}
template <int R, int C>
VMP_HD Mat<C, R> tr(const Mat<R, C>& A) {
    Mat<C, R> o;
#pragma unroll
    for (int i = 0; i < R; i++)
#pragma unroll
        for (int j = 0; j < C; j++) o(j, i) = A(i, j);
    return o;
}
template <int R, int C>
VMP_HD Mat<R, C> add(const Mat<R, C>& A, const Mat<R, C>& B) { Mat<R, C> o; for (int i = 0; i < R * C; i++) o.a[i] = A.a[i] + B.a[i]; return o; }
template <int R, int C>
VMP_HD Mat<R, C> sub(const Mat<R, C>& A, const Mat<R, C>& B) { Mat<R, C> o; for (int i = 0; i < R * C; i++) o.a[i] = A.a[i] - B.a[i]; return o; }
template <int R, int C>
VMP_HD Mat<R, C> scale(const Mat<R, C>& A, double s) { Mat<R, C> o; for (int i = 0; i < R * C; i++) o.a[i] = A.a[i] * s; return o; }
template <int R, int C>
VMP_HD Mat<R, C> divs(const Mat<R, C>& A, double s) { Mat<R, C> o; for (int i = 0; i < R * C; i++) o.a[i] = A.a[i] / s; return o; }
template <int R, int C>
VMP_HD Mat<R, C> neg(const Mat<R, C>& A) { Mat<R, C> o; for (int i = 0; i < R * C; i++) o.a[i] = -A.a[i]; return o; }

VMP_HD double dot(const V3& a, const V3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
VMP_HD double norm(const V3& a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
VMP_HD V3 normalized(const V3& a) { return divs(a, norm(a)); }
VMP_HD V3 cross(const V3& a, const V3& b) {
    return v3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
VMP_HD M3 outer(const V3& a, const V3& b) {
    M3 m;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) m(i, j) = a[i] * b[j];
    return m;
}
VMP_HD M3 hat(const V3& v) {
    M3 m;
    m(0, 0) = 0.0;   m(0, 1) = -v[2]; m(0, 2) = v[1];
    m(1, 0) = v[2];  m(1, 1) = 0.0;   m(1, 2) = -v[0];
    m(2, 0) = -v[1]; m(2, 1) = v[0];  m(2, 2) = 0.0;
    return m;
}
template <int BR, int BC, int R, int C>
VMP_HD void set_block(Mat<R, C>& dst, int r0, int c0, const Mat<BR, BC>& src) {
    for (int i = 0; i < BR; i++) for (int j = 0; j < BC; j++) dst(r0 + i, c0 + j) = src(i, j);
}

// ---- quaternion / SO(3) (Eigen Quaterniond + Sophus SO3d semantics) -----------------
struct Quat { double w, x, y, z; };

VMP_HD M3 quat_to_rot(const Quat& q) {
    const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    M3 r;
    r(0, 0) = 1.0 - (tyy + tzz); r(0, 1) = txy - twz;         r(0, 2) = txz + twy;
    r(1, 0) = txy + twz;         r(1, 1) = 1.0 - (txx + tzz); r(1, 2) = tyz - twx;
    r(2, 0) = txz - twy;         r(2, 1) = tyz + twx;         r(2, 2) = 1.0 - (txx + tyy);
    return r;
}

VMP_HD Quat rot_to_quat(const M3& m) {
    Quat q;
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0.0) {
        t = sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (m(2, 1) - m(1, 2)) * t;
        q.y = (m(0, 2) - m(2, 0)) * t;
        q.z = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if (m(1, 1) > m(0, 0)) i = 1;
        if (m(2, 2) > m(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
        double v[3];
        v[i] = 0.5 * t;
        t = 0.5 / t;
        q.w = (m(k, j) - m(j, k)) * t;
        v[j] = (m(j, i) + m(i, j)) * t;
        v[k] = (m(k, i) + m(i, k)) * t;
        q.x = v[0]; q.y = v[1]; q.z = v[2];
    }
    return q;
}

VMP_HD M3 so3_exp(const V3& w) {
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double im, re;
    if (th2 < 1e-10 * 1e-10) {
        const double th4 = th2 * th2;
        im = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
        re = 1.0 - (1.0 / 8.0) * th2 + (1.0 / 384.0) * th4;
    } else {
        const double th = sqrt(th2);
        const double half = 0.5 * th;
        double sh, ch;
        sincos(half, &sh, &ch);          // (one argument reduction for both; same values as sin() and cos())
        im = sh / th;
        re = ch;
    }
    Quat q; q.w = re; q.x = im * w[0]; q.y = im * w[1]; q.z = im * w[2];
    return quat_to_rot(q);
}

VMP_HD V3 so3_log(const M3& R) {
    Quat q = rot_to_quat(R);
    const double qn = sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    q.w /= qn; q.x /= qn; q.y /= qn; q.z /= qn;
    const double n2 = q.x * q.x + q.y * q.y + q.z * q.z;
    const double w = q.w;
    double k;
    if (n2 < 1e-10 * 1e-10) {
        k = 2.0 / w - (2.0 / 3.0) * n2 / (w * w * w);
    } else {
        const double n = sqrt(n2);
        const double at = (w < 0.0) ? atan2(-n, -w) : atan2(n, w);
        k = 2.0 * at / n;
    }
    return v3(k * q.x, k * q.y, k * q.z);
}

VMP_HD M3 so3_left_jacobian(const V3& w) {
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    const M3 Om = hat(w);
    const M3 I = eye<3>();
    if (th2 < 1e-10 * 1e-10) return add(I, scale(Om, 0.5));
    const double th = sqrt(th2);
    const M3 Om2 = mul(Om, Om);
    double st, ct;
    sincos(th, &st, &ct);
    const double c1 = (1.0 - ct) / th2;
    const double c2 = (th - st) / (th2 * th);
    return add(add(I, scale(Om, c1)), scale(Om2, c2));
}
VMP_HD M3 right_jacobian(const V3& w) { return tr(so3_left_jacobian(w)); }

VMP_HD M3 rot_from_two_vectors(const V3& a_in, const V3& b_in) {
    const V3 a = normalized(a_in), b = normalized(b_in);
    const double c = dot(a, b);
    Quat q;
    if (c < -1.0 + 1e-12) {
        V3 ax = cross(a, v3(1, 0, 0));
        if (norm(ax) < 1e-6) ax = cross(a, v3(0, 1, 0));
        ax = normalized(ax);
        q.w = 0.0; q.x = ax[0]; q.y = ax[1]; q.z = ax[2];
        return quat_to_rot(q);
    }
    const V3 axis = cross(a, b);
    const double s = sqrt((1.0 + c) * 2.0);
    const double invs = 1.0 / s;
    q.x = axis[0] * invs; q.y = axis[1] * invs; q.z = axis[2] * invs; q.w = s * 0.5;
    return quat_to_rot(q);
}

// ---- 3x3 symmetric eigen-solver --------------------------------------------------
// Tridiagonalise (closed form for 3x3) + implicit-shift QR, eigenvalues ascending,
// eigenvectors in columns; input = LOWER triangle a00 a10 a11 a20 a21 a22.
VMP_HD void eig3_sym(double a00, double a10, double a11, double a20, double a21, double a22,
                     double evals[3], M3& evecs) {
    double sc = fabs(a00);
    if (fabs(a10) > sc) sc = fabs(a10);
    if (fabs(a11) > sc) sc = fabs(a11);
    if (fabs(a20) > sc) sc = fabs(a20);
    if (fabs(a21) > sc) sc = fabs(a21);
    if (fabs(a22) > sc) sc = fabs(a22);
    if (sc == 0.0) sc = 1.0;
    a00 /= sc; a10 /= sc; a11 /= sc; a20 /= sc; a21 /= sc; a22 /= sc;

    // diag / sub / Q live in scalars with compile-time indices (the iteration only ever works on rows (0,1), (1,2) or (0,1,2)): with
    // arrays indexed by the loop variables they sat in local memory, and the eigen-solve - one lane per refit, the longest single step
    // of a voxel in k_fill - was a chain of local loads and stores.  Every expression is the one of Eigen's tridiagonal QR step.
    double d0, d1, d2, s0, s1;
    M3 Q;
    const double tiny = 2.2250738585072014e-308;
    d0 = a00;
    const double v1norm2 = a20 * a20;
    if (v1norm2 <= tiny) {
        d1 = a11; d2 = a22;
        s0 = a10; s1 = a21;
        Q = eye<3>();
    } else {
        const double beta = sqrt(a10 * a10 + v1norm2);
        const double invBeta = 1.0 / beta;
        const double m01 = a10 * invBeta;
        const double m02 = a20 * invBeta;
        const double q = 2.0 * m01 * a21 + m02 * (a22 - a11);
        d1 = a11 + m02 * q;
        d2 = a22 - m02 * q;
        s0 = beta;
        s1 = a21 - m01 * q;
        Q(0, 0) = 1; Q(0, 1) = 0;   Q(0, 2) = 0;
        Q(1, 0) = 0; Q(1, 1) = m01; Q(1, 2) = m02;
        Q(2, 0) = 0; Q(2, 1) = m02; Q(2, 2) = -m01;
    }
    int end = 2, start = 0, iter = 0;
    const double precision = 2.0 * 2.220446049250313e-16;
#define VMP_EIG_ROT(DK, DK1, SK, CK, CK1)                                                      \
    {                                                                                           \
        const double sdk = s * DK + c * SK;                                                     \
        const double dkp1 = s * SK + c * DK1;                                                   \
        const double ndk = c * (c * DK - s * SK) - s * (c * SK - s * DK1);                      \
        DK = ndk;                                                                               \
        DK1 = s * sdk + c * dkp1;                                                               \
        SK = c * sdk - s * dkp1;                                                                \
    }
#define VMP_EIG_QCOL(CK, CK1)                                                                   \
    for (int i = 0; i < 3; i++) {                                                               \
        const double xi = Q(i, CK), yi = Q(i, CK1);                                             \
        Q(i, CK) = c * xi - s * yi;                                                             \
        Q(i, CK1) = s * xi + c * yi;                                                            \
    }
    while (end > 0) {
        if (start <= 0 && 0 < end) { if (fabs(s0) <= (fabs(d0) + fabs(d1)) * precision || fabs(s0) <= tiny) s0 = 0.0; }
        if (start <= 1 && 1 < end) { if (fabs(s1) <= (fabs(d1) + fabs(d2)) * precision || fabs(s1) <= tiny) s1 = 0.0; }
        while (end > 0 && (end == 2 ? s1 : s0) == 0.0) end--;
        if (end <= 0) break;
        iter++;
        if (iter > 90) break;
        start = end - 1;
        if (start == 1 && s0 != 0.0) start = 0;
        const double dEm1 = end == 2 ? d1 : d0, dE = end == 2 ? d2 : d1;
        const double td = (dEm1 - dE) * 0.5;
        const double e = end == 2 ? s1 : s0;
        double mu = dE;
        if (td == 0.0) {
            mu -= fabs(e);
        } else if (e != 0.0) {
            const double e2 = e * e;
            const double ax = fabs(td), ay = fabs(e);
            double p, qp;
            if (ax > ay) { p = ax; qp = ay / p; } else { p = ay; qp = ax / p; }
            const double h = (p == 0.0) ? 0.0 : p * sqrt(1.0 + qp * qp);
            if (e2 == 0.0) mu -= e / ((td + (td > 0.0 ? h : -h)) / e);
            else mu -= e2 / (td + (td > 0.0 ? h : -h));
        }
        double x = (start == 0 ? d0 : d1) - mu;
        double z = start == 0 ? s0 : s1;
        for (int k = start; k < end && z != 0.0; ++k) {
            double c, s;
            if (z == 0.0) { c = x < 0.0 ? -1.0 : 1.0; s = 0.0; }
            else if (x == 0.0) { c = 0.0; s = z < 0.0 ? 1.0 : -1.0; }
            else if (fabs(x) > fabs(z)) {
                const double t = z / x; double u = sqrt(1.0 + t * t); if (x < 0.0) u = -u;
                c = 1.0 / u; s = -t * c;
            } else {
                const double t = x / z; double u = sqrt(1.0 + t * t); if (z < 0.0) u = -u;
                s = -1.0 / u; c = -t * s;
            }
            if (k == 0) {
                VMP_EIG_ROT(d0, d1, s0, 0, 1)
                x = s0;
                if (end == 2) { z = -s * s1; s1 = c * s1; }          // k < end - 1
                VMP_EIG_QCOL(0, 1)
            } else {
                VMP_EIG_ROT(d1, d2, s1, 1, 2)
                if (start == 0) s0 = c * s0 - s * z;                  // k > start
                x = s1;
                VMP_EIG_QCOL(1, 2)
            }
        }
    }
#undef VMP_EIG_ROT
#undef VMP_EIG_QCOL
    if (iter <= 90) {
        {   // i = 0: the smallest of (d0, d1, d2) to the front
            int k = 0; double best = d0;
            if (d1 < best) { best = d1; k = 1; }
            if (d2 < best) { best = d2; k = 2; }
            if (k == 1) { const double t = d0; d0 = d1; d1 = t; for (int r = 0; r < 3; r++) { const double u = Q(r, 0); Q(r, 0) = Q(r, 1); Q(r, 1) = u; } }
            if (k == 2) { const double t = d0; d0 = d2; d2 = t; for (int r = 0; r < 3; r++) { const double u = Q(r, 0); Q(r, 0) = Q(r, 2); Q(r, 2) = u; } }
        }
        if (d2 < d1) { const double t = d1; d1 = d2; d2 = t; for (int r = 0; r < 3; r++) { const double u = Q(r, 1); Q(r, 1) = Q(r, 2); Q(r, 2) = u; } }
    }
    double diag[3] = {d0, d1, d2};
    evals[0] = diag[0] * sc; evals[1] = diag[1] * sc; evals[2] = diag[2] * sc;
    evecs = Q;
}

// ---- LU (partial pivoting) inverse, N x N row-major; serial reference order --------
template <int N>
VMP_HD void lu_inverse(const double* Ain, double* inv, double* lu /*N*N scratch*/) {
    int perm[N];
    for (int i = 0; i < N * N; i++) lu[i] = Ain[i];
    for (int i = 0; i < N; i++) perm[i] = i;
    for (int k = 0; k < N; k++) {
        int piv = k; double best = fabs(lu[k * N + k]);
        for (int i = k + 1; i < N; i++) { const double v = fabs(lu[i * N + k]); if (v > best) { best = v; piv = i; } }
        if (piv != k) {
            for (int j = 0; j < N; j++) { const double t = lu[k * N + j]; lu[k * N + j] = lu[piv * N + j]; lu[piv * N + j] = t; }
            const int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
        }
        const double d = lu[k * N + k];
        for (int i = k + 1; i < N; i++) lu[i * N + k] = lu[i * N + k] / d;
        for (int i = k + 1; i < N; i++) {
            const double l = lu[i * N + k];
            for (int j = k + 1; j < N; j++) lu[i * N + j] = lu[i * N + j] - l * lu[k * N + j];
        }
    }
    for (int c = 0; c < N; c++) {
        double y[N];
        for (int i = 0; i < N; i++) {
            double s = (perm[i] == c) ? 1.0 : 0.0;
            for (int j = 0; j < i; j++) s = s - lu[i * N + j] * y[j];
            y[i] = s;
        }
        for (int i = N - 1; i >= 0; i--) {       // terms subtracted in the order the unknowns become available
            double s = y[i];
            for (int j = N - 1; j > i; j--) s = s - lu[i * N + j] * y[j];
            y[i] = s / lu[i * N + i];
        }
        for (int i = 0; i < N; i++) inv[i * N + c] = y[i];
    }
}

// ---- commons.cpp:18-45 calcBodyCov; sn2 = sin(angle_cov * 0.017453293)^2 and
// range_var = ranging_cov^2 are evaluated once on the host.  Mutates p.z 0 -> 0.001 (Q16).
VMP_HD void calc_body_cov(V3& pb, double range_var, double sn2, M3& cov) {
    if (pb[2] == 0) pb[2] = 0.001;
    const double range = sqrt(pb[0] * pb[0] + pb[1] * pb[1] + pb[2] * pb[2]);
    const V3 d = normalized(pb);
    const M3 dh = hat(d);
    const V3 b1 = normalized(v3(1.0, 1.0, -(d[0] + d[1]) / d[2]));
    const V3 b2 = normalized(cross(b1, d));
    Mat<3, 2> N;
    N(0, 0) = b1[0]; N(0, 1) = b2[0];
    N(1, 0) = b1[1]; N(1, 1) = b2[1];
    N(2, 0) = b1[2]; N(2, 1) = b2[2];
    const Mat<3, 2> A = mul(scale(dh, range), N);
    Mat<2, 2> dv; dv(0, 0) = sn2; dv(0, 1) = 0.0; dv(1, 0) = 0.0; dv(1, 1) = sn2;
    cov = add(mul(scale(d, range_var), tr(d)), mul(mul(A, dv), tr(A)));
}

// cov_world = R C R^T + [p]x Prr [p]x^T + Ppp   (lio_builder.cpp:199-203, 240-244, 264-267)
VMP_HD M3 world_cov(const M3& r_wl, const M3& cov_l, const V3& p_l, const M3& Prr, const M3& Ppp) {
    const M3 cm = hat(p_l);
    return add(add(mul(mul(r_wl, cov_l), tr(r_wl)), mul(mul(cm, Prr), tr(cm))), Ppp);
}

}  // namespace vmp
