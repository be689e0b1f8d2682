// oracle_math.h — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Small fixed-size linear algebra + Lie-group helpers for the CPU oracle.  The
// reference gets these from Eigen 3 / Sophus, neither of which exists in this image
//  and neither of which is version-pinned by the reference, so the
// evaluation order written here IS the oracle's definition ("parity unpinned").
// Compile with -ffp-contract=off: every a*b+c below is two roundings.
//
// Restated third-party algorithms (from their published sources, from memory):
//   * Eigen 3.3 SelfAdjointEigenSolver<Matrix3d>::compute()   -> eig3_sym()
//     (call site: voxel_map.cpp:104-106)
//   * Eigen PartialPivLU inverse for Matrix<double,23,23>      -> inverse()
//     (call sites: ieskf.cpp:141,142,145,155)
//   * Eigen Quaterniond(R), toRotationMatrix, FromTwoVectors   -> quat_*()
//   * Sophus SO3d::exp / log / hat / leftJacobian              -> so3_*()
#pragma once
#include <cmath>
#include <cstring>
#include <limits>

namespace orc {

template <int R, int C>
struct Mat {
    double a[R * C];
    double& operator()(int i, int j) { return a[i * C + j]; }
    const double& operator()(int i, int j) const { return a[i * C + j]; }
    double& operator[](int i) { return a[i]; }
    const double& operator[](int i) const { return a[i]; }
    static Mat zero() { Mat m; for (int i = 0; i < R * C; i++) m.a[i] = 0.0; return m; }
    static Mat identity() { Mat m = zero(); for (int i = 0; i < (R < C ? R : C); i++) m(i, i) = 1.0; return m; }
};
typedef Mat<3, 1> V3;
typedef Mat<3, 3> M3;
typedef Mat<6, 6> M6;
typedef Mat<12, 12> M12;
typedef Mat<12, 1> V12;
typedef Mat<23, 23> M23;
typedef Mat<23, 1> V23;

inline V3 v3(double x, double y, double z) { V3 v; v[0] = x; v[1] = y; v[2] = z; return v; }

// C = A*B, coefficient-wise: c_ij = a_i0 b_0j + a_i1 b_1j + ... (left to right, first product seeds the sum)
template <int R, int K, int C>
inline Mat<R, C> mul(const Mat<R, K>& A, const Mat<K, C>& B) {
    Mat<R, C> out;
    for (int i = 0; i < R; i++)
        for (int j = 0; j < C; j++) {
            double s = A(i, 0) * B(0, j);
            for (int k = 1; k < K; k++) s += A(i, k) * B(k, j);
            out(i, j) = s;
        }
    return out;
}
template <int R, int C>
inline Mat<C, R> tr(const Mat<R, C>& A) {
    Mat<C, R> out;
    for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) out(j, i) = A(i, j);
    return out;
}
template <int R, int C>
inline Mat<R, C> add(const Mat<R, C>& A, const Mat<R, C>& B) { Mat<R, C> o; for (int i = 0; i < R * C; i++) o.a[i] = A.a[i] + B.a[i]; return o; }
template <int R, int C>
inline Mat<R, C> sub(const Mat<R, C>& A, const Mat<R, C>& B) { Mat<R, C> o; for (int i = 0; i < R * C; i++) o.a[i] = A.a[i] - B.a[i]; return o; }
template <int R, int C>
inline Mat<R, C> scale(const Mat<R, C>& A, double s) { Mat<R, C> o; for (int i = 0; i < R * C; i++) o.a[i] = A.a[i] * s; return o; }
template <int R, int C>
inline Mat<R, C> divs(const Mat<R, C>& A, double s) { Mat<R, C> o; for (int i = 0; i < R * C; i++) o.a[i] = A.a[i] / s; return o; }
template <int R, int C>
inline Mat<R, C> neg(const Mat<R, C>& A) { Mat<R, C> o; for (int i = 0; i < R * C; i++) o.a[i] = -A.a[i]; return o; }

inline double dot(const V3& a, const V3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline double norm(const V3& a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
inline V3 normalized(const V3& a) { double n = norm(a); return divs(a, n); }
inline V3 cross(const V3& a, const V3& b) {
    return v3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
// p p^T
inline M3 outer(const V3& a, const V3& b) { M3 m; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m(i, j) = a[i] * b[j]; return m; }

template <int R, int C, int BR, int BC>
inline void set_block(Mat<R, C>& dst, int r0, int c0, const Mat<BR, BC>& src) {
    for (int i = 0; i < BR; i++) for (int j = 0; j < BC; j++) dst(r0 + i, c0 + j) = src(i, j);
}
template <int BR, int BC, int R, int C>
inline Mat<BR, BC> get_block(const Mat<R, C>& src, int r0, int c0) {
    Mat<BR, BC> o; for (int i = 0; i < BR; i++) for (int j = 0; j < BC; j++) o(i, j) = src(r0 + i, c0 + j); return o;
}

// Sophus::SO3d::hat
inline M3 hat(const V3& v) {
    M3 m;
    m(0, 0) = 0.0;   m(0, 1) = -v[2]; m(0, 2) = v[1];
    m(1, 0) = v[2];  m(1, 1) = 0.0;   m(1, 2) = -v[0];
    m(2, 0) = -v[1]; m(2, 1) = v[0];  m(2, 2) = 0.0;
    return m;
}

struct Quat { double w, x, y, z; };

// Eigen Quaternion::toRotationMatrix
inline M3 quat_to_rot(const Quat& q) {
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

// Eigen Quaternion(Matrix3) (trace branch method)
inline Quat rot_to_quat(const M3& m) {
    Quat q;
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0.0) {
        t = std::sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (m(2, 1) - m(1, 2)) * t;
        q.y = (m(0, 2) - m(2, 0)) * t;
        q.z = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if (m(1, 1) > m(0, 0)) i = 1;
        if (m(2, 2) > m(i, i)) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
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

// Sophus SO3::exp(omega).matrix()
inline M3 so3_exp(const V3& w) {
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double im, re;
    if (th2 < 1e-10 * 1e-10) {
        const double th4 = th2 * th2;
        im = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
        re = 1.0 - (1.0 / 8.0) * th2 + (1.0 / 384.0) * th4;
    } else {
        const double th = std::sqrt(th2);
        const double half = 0.5 * th;
        im = std::sin(half) / th;
        re = std::cos(half);
    }
    Quat q; q.w = re; q.x = im * w[0]; q.y = im * w[1]; q.z = im * w[2];
    return quat_to_rot(q);
}

// Sophus SO3(R).log(): quaternion from the matrix, then atan-based log
inline V3 so3_log(const M3& R) {
    Quat q = rot_to_quat(R);
    // Sophus normalises the quaternion it stores
    const double qn = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    q.w /= qn; q.x /= qn; q.y /= qn; q.z /= qn;
    const double n2 = q.x * q.x + q.y * q.y + q.z * q.z;
    const double w = q.w;
    double k;
    if (n2 < 1e-10 * 1e-10) {
        k = 2.0 / w - (2.0 / 3.0) * n2 / (w * w * w);
    } else {
        const double n = std::sqrt(n2);
        const double at = (w < 0.0) ? std::atan2(-n, -w) : std::atan2(n, w);
        k = 2.0 * at / n;
    }
    return v3(k * q.x, k * q.y, k * q.z);
}

// Sophus SO3::leftJacobian
inline M3 so3_left_jacobian(const V3& w) {
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    const M3 Om = hat(w);
    M3 I = M3::identity();
    if (th2 < 1e-10 * 1e-10) return add(I, scale(Om, 0.5));
    const double th = std::sqrt(th2);
    const M3 Om2 = mul(Om, Om);
    const double c1 = (1.0 - std::cos(th)) / th2;
    const double c2 = (th - std::sin(th)) / (th2 * th);
    return add(add(I, scale(Om, c1)), scale(Om2, c2));
}
// kf::rightJacobian (ieskf.cpp:6-9)
inline M3 right_jacobian(const V3& w) { return tr(so3_left_jacobian(w)); }

// Eigen Quaterniond::FromTwoVectors(a,b).matrix() (lio_builder.cpp:48). The c ~ -1
// branch (SVD in Eigen) is replaced by a deterministic orthogonal axis.
inline M3 rot_from_two_vectors(const V3& a_in, const V3& b_in) {
    V3 a = normalized(a_in), b = normalized(b_in);
    double c = dot(a, b);
    Quat q;
    if (c < -1.0 + 1e-12) {
        V3 ax = cross(a, v3(1, 0, 0));
        if (norm(ax) < 1e-6) ax = cross(a, v3(0, 1, 0));
        ax = normalized(ax);
        q.w = 0.0; q.x = ax[0]; q.y = ax[1]; q.z = ax[2];
        return quat_to_rot(q);
    }
    V3 axis = cross(a, b);
    double s = std::sqrt((1.0 + c) * 2.0);
    double invs = 1.0 / s;
    q.x = axis[0] * invs; q.y = axis[1] * invs; q.z = axis[2] * invs; q.w = s * 0.5;
    return quat_to_rot(q);
}

// ---- Eigen::SelfAdjointEigenSolver<Matrix3d>::compute(), iterative path ----------
// reads the LOWER triangle; eigenvalues ascending; eigenvectors in columns.
// returns false on non-convergence (never observed for 3x3).
inline bool eig3_sym(const M3& A, double evals[3], M3& evecs) {
    double m[3][3];
    m[0][0] = A(0, 0); m[1][0] = A(1, 0); m[1][1] = A(1, 1);
    m[2][0] = A(2, 0); m[2][1] = A(2, 1); m[2][2] = A(2, 2);
    double sc = std::fabs(m[0][0]);
    if (std::fabs(m[1][0]) > sc) sc = std::fabs(m[1][0]);
    if (std::fabs(m[1][1]) > sc) sc = std::fabs(m[1][1]);
    if (std::fabs(m[2][0]) > sc) sc = std::fabs(m[2][0]);
    if (std::fabs(m[2][1]) > sc) sc = std::fabs(m[2][1]);
    if (std::fabs(m[2][2]) > sc) sc = std::fabs(m[2][2]);
    if (sc == 0.0) sc = 1.0;
    m[0][0] /= sc; m[1][0] /= sc; m[1][1] /= sc; m[2][0] /= sc; m[2][1] /= sc; m[2][2] /= sc;

    double diag[3], sub[2];
    double Q[3][3];
    const double tiny = std::numeric_limits<double>::min();
    // closed-form 3x3 tridiagonalisation
    diag[0] = m[0][0];
    const double v1norm2 = m[2][0] * m[2][0];
    if (v1norm2 <= tiny) {
        diag[1] = m[1][1]; diag[2] = m[2][2];
        sub[0] = m[1][0]; sub[1] = m[2][1];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Q[i][j] = (i == j) ? 1.0 : 0.0;
    } else {
        const double beta = std::sqrt(m[1][0] * m[1][0] + v1norm2);
        const double invBeta = 1.0 / beta;
        const double m01 = m[1][0] * invBeta;
        const double m02 = m[2][0] * invBeta;
        const double q = 2.0 * m01 * m[2][1] + m02 * (m[2][2] - m[1][1]);
        diag[1] = m[1][1] + m02 * q;
        diag[2] = m[2][2] - m02 * q;
        sub[0] = beta;
        sub[1] = m[2][1] - m01 * q;
        Q[0][0] = 1; Q[0][1] = 0;   Q[0][2] = 0;
        Q[1][0] = 0; Q[1][1] = m01; Q[1][2] = m02;
        Q[2][0] = 0; Q[2][1] = m02; Q[2][2] = -m01;
    }
    // implicit symmetric QR with Wilkinson shift
    const int n = 3;
    int end = n - 1, start = 0, iter = 0;
    const int maxIter = 30;
    const double precision = 2.0 * std::numeric_limits<double>::epsilon();
    while (end > 0) {
        for (int i = start; i < end; ++i)
            if (std::fabs(sub[i]) <= (std::fabs(diag[i]) + std::fabs(diag[i + 1])) * precision || std::fabs(sub[i]) <= tiny)
                sub[i] = 0.0;
        while (end > 0 && sub[end - 1] == 0.0) end--;
        if (end <= 0) break;
        iter++;
        if (iter > maxIter * n) break;
        start = end - 1;
        while (start > 0 && sub[start - 1] != 0.0) start--;
        // one QR step on [start, end]
        double td = (diag[end - 1] - diag[end]) * 0.5;
        double e = sub[end - 1];
        double mu = diag[end];
        if (td == 0.0) {
            mu -= std::fabs(e);
        } else if (e != 0.0) {
            const double e2 = e * e;
            // Eigen numext::hypot
            double ax = std::fabs(td), ay = std::fabs(e), p, qp;
            if (ax > ay) { p = ax; qp = ay / p; } else { p = ay; qp = ax / p; }
            const double h = (p == 0.0) ? 0.0 : p * std::sqrt(1.0 + qp * qp);
            if (e2 == 0.0) mu -= e / ((td + (td > 0.0 ? h : -h)) / e);
            else mu -= e2 / (td + (td > 0.0 ? h : -h));
        }
        double x = diag[start] - mu;
        double z = sub[start];
        for (int k = start; k < end && z != 0.0; ++k) {
            // JacobiRotation::makeGivens(x, z)
            double c, s;
            if (z == 0.0) { c = x < 0.0 ? -1.0 : 1.0; s = 0.0; }
            else if (x == 0.0) { c = 0.0; s = z < 0.0 ? 1.0 : -1.0; }
            else if (std::fabs(x) > std::fabs(z)) {
                double t = z / x; double u = std::sqrt(1.0 + t * t); if (x < 0.0) u = -u;
                c = 1.0 / u; s = -t * c;
            } else {
                double t = x / z; double u = std::sqrt(1.0 + t * t); if (z < 0.0) u = -u;
                s = -1.0 / u; c = -t * s;
            }
            const double sdk = s * diag[k] + c * sub[k];
            const double dkp1 = s * sub[k] + c * diag[k + 1];
            diag[k] = c * (c * diag[k] - s * sub[k]) - s * (c * sub[k] - s * diag[k + 1]);
            diag[k + 1] = s * sdk + c * dkp1;
            sub[k] = c * sdk - s * dkp1;
            if (k > start) sub[k - 1] = c * sub[k - 1] - s * z;
            x = sub[k];
            if (k < end - 1) { z = -s * sub[k + 1]; sub[k + 1] = c * sub[k + 1]; }
            // Q = Q * G : columns k, k+1
            for (int i = 0; i < 3; i++) {
                const double xi = Q[i][k], yi = Q[i][k + 1];
                Q[i][k] = c * xi - s * yi;
                Q[i][k + 1] = s * xi + c * yi;
            }
        }
    }
    const bool ok = iter <= maxIter * n;
    // selection sort ascending, swapping eigenvector columns
    if (ok)
    for (int i = 0; i < n - 1; ++i) {
        int k = 0; double best = diag[i];
        for (int j = 1; j < n - i; ++j) if (diag[i + j] < best) { best = diag[i + j]; k = j; }
        if (k > 0) {
            double t = diag[i]; diag[i] = diag[k + i]; diag[k + i] = t;
            for (int r = 0; r < 3; r++) { double u = Q[r][i]; Q[r][i] = Q[r][k + i]; Q[r][k + i] = u; }
        }
    }
    for (int i = 0; i < 3; i++) evals[i] = diag[i] * sc;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) evecs(i, j) = Q[i][j];
    return ok;
}

// ---- general inverse via LU with partial (row) pivoting, then solve against I ----
// Unblocked right-looking LU; the first row of maximal |pivot| wins (Eigen picks the
// first maximum).  L has unit diagonal.  Forward/back substitution column by column.
template <int N>
inline Mat<N, N> inverse(const Mat<N, N>& Ain) {
    Mat<N, N> lu = Ain;
    int perm[N];
    for (int i = 0; i < N; i++) perm[i] = i;
    for (int k = 0; k < N; k++) {
        int piv = k; double best = std::fabs(lu(k, k));
        for (int i = k + 1; i < N; i++) { double v = std::fabs(lu(i, k)); if (v > best) { best = v; piv = i; } }
        if (piv != k) {
            for (int j = 0; j < N; j++) { double t = lu(k, j); lu(k, j) = lu(piv, j); lu(piv, j) = t; }
            int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
        }
        const double d = lu(k, k);
        for (int i = k + 1; i < N; i++) lu(i, k) = lu(i, k) / d;
        for (int i = k + 1; i < N; i++) {
            const double l = lu(i, k);
            for (int j = k + 1; j < N; j++) lu(i, j) = lu(i, j) - l * lu(k, j);
        }
    }
    Mat<N, N> inv;
    for (int c = 0; c < N; c++) {
        double y[N];
        // forward: L y = P e_c
        for (int i = 0; i < N; i++) {
            double s = (perm[i] == c) ? 1.0 : 0.0;
            for (int j = 0; j < i; j++) s = s - lu(i, j) * y[j];
            y[i] = s;
        }
        // backward: U x = y; terms subtracted in the order the unknowns become available (j = N-1 first)
        for (int i = N - 1; i >= 0; i--) {
            double s = y[i];
            for (int j = N - 1; j > i; j--) s = s - lu(i, j) * y[j];
            y[i] = s / lu(i, i);
        }
        for (int i = 0; i < N; i++) inv(i, c) = y[i];
    }
    return inv;
}

}  // namespace orc
