// vmp_state.cuh — kf::State manifold operations (reference ieskf.h:31-62, ieskf.cpp:11-95)
// and IESKF::predict (ieskf.cpp:101-123), host + device.  Flat layout = vmp_state:
// pos3 rot9 rot_ext9 pos_ext3 vel3 bg3 ba3 g3 (36 doubles).
#pragma once
#include "vmp_math.cuh"

namespace vmp {

constexpr double GRAVITY = 9.81;

struct St {
    V3 pos; M3 rot; M3 rot_ext; V3 pos_ext; V3 vel; V3 bg; V3 ba; V3 g;
};

VMP_HD St st_load(const double* x) {
    St s;
    for (int i = 0; i < 3; i++) { s.pos[i] = x[i]; s.pos_ext[i] = x[21 + i]; s.vel[i] = x[24 + i]; s.bg[i] = x[27 + i]; s.ba[i] = x[30 + i]; s.g[i] = x[33 + i]; }
    for (int i = 0; i < 9; i++) { s.rot.a[i] = x[3 + i]; s.rot_ext.a[i] = x[12 + i]; }
    return s;
}
VMP_HD void st_store(const St& s, double* x) {
    for (int i = 0; i < 3; i++) { x[i] = s.pos[i]; x[21 + i] = s.pos_ext[i]; x[24 + i] = s.vel[i]; x[27 + i] = s.bg[i]; x[30 + i] = s.ba[i]; x[33 + i] = s.g[i]; }
    for (int i = 0; i < 9; i++) { x[3 + i] = s.rot.a[i]; x[12 + i] = s.rot_ext.a[i]; }
}

// State::getBx  ieskf.cpp:69-77
VMP_HD Mat<3, 2> st_Bx(const V3& g) {
    Mat<3, 2> r;
    r(0, 0) = -g[1];
    r(0, 1) = -g[2];
    r(1, 0) = GRAVITY - g[1] * g[1] / (GRAVITY + g[0]);
    r(1, 1) = -g[2] * g[1] / (GRAVITY + g[0]);
    r(2, 0) = -g[2] * g[1] / (GRAVITY + g[0]);
    r(2, 1) = GRAVITY - g[2] * g[2] / (GRAVITY + g[0]);
    return divs(r, GRAVITY);
}
// State::getMx()  ieskf.cpp:79-83
VMP_HD Mat<3, 2> st_Mx(const V3& g) { return mul(neg(hat(g)), st_Bx(g)); }
// State::getMx(res)  ieskf.cpp:85-90
VMP_HD Mat<3, 2> st_Mx_res(const V3& g, const Mat<2, 1>& res) {
    const Mat<3, 2> bx = st_Bx(g);
    const V3 bu = mul(bx, res);
    const M3 a = mul(neg(so3_exp(bu)), hat(g));
    const M3 b = mul(a, tr(so3_left_jacobian(bu)));
    return mul(b, bx);
}
// State::getNx  ieskf.cpp:92-95
VMP_HD Mat<2, 3> st_Nx(const V3& g) {
    const double s = 1 / GRAVITY / GRAVITY;
    return mul(scale(tr(st_Bx(g)), s), hat(g));
}

// State::operator+=(Vector23d)  ieskf.cpp:11-21
VMP_HD void st_boxplus(St& s, const double* d) {
    s.pos = add(s.pos, v3(d[0], d[1], d[2]));
    s.rot = mul(s.rot, so3_exp(v3(d[3], d[4], d[5])));
    s.rot_ext = mul(s.rot_ext, so3_exp(v3(d[6], d[7], d[8])));
    s.pos_ext = add(s.pos_ext, v3(d[9], d[10], d[11]));
    s.vel = add(s.vel, v3(d[12], d[13], d[14]));
    s.bg = add(s.bg, v3(d[15], d[16], d[17]));
    s.ba = add(s.ba, v3(d[18], d[19], d[20]));
    Mat<2, 1> dg; dg[0] = d[21]; dg[1] = d[22];
    s.g = mul(so3_exp(mul(st_Bx(s.g), dg)), s.g);
}

// S2 part of State::operator-  ieskf.cpp:55-65   (ag = gravity of the left operand, bg = of the right operand)
VMP_HD void st_boxminus_g(const V3& ag, const V3& bg, double* out2) {
    const double v_sin = norm(mul(hat(ag), bg));
    const double v_cos = dot(ag, bg);
    const double theta = atan2(v_sin, v_cos);
    Mat<2, 1> res;
    if (v_sin < 1e-11) {
        if (fabs(theta) > 1e-11) { res[0] = 3.1415926; res[1] = 0.0; }
        else { res[0] = 0.0; res[1] = 0.0; }
    } else {
        const Mat<2, 3> p = scale(tr(st_Bx(bg)), theta / v_sin);
        const Mat<2, 3> q = mul(p, hat(bg));
        res = mul(q, ag);
    }
    out2[0] = res[0]; out2[1] = res[1];
}

// State::operator-  ieskf.cpp:35-67   (delta = a [-] b, 23 entries)
VMP_HD void st_boxminus(const St& a, const St& b, double* delta) {
    V3 t;
    t = sub(a.pos, b.pos);                         delta[0] = t[0];  delta[1] = t[1];  delta[2] = t[2];
    t = so3_log(mul(tr(b.rot), a.rot));            delta[3] = t[0];  delta[4] = t[1];  delta[5] = t[2];
    t = so3_log(mul(tr(b.rot_ext), a.rot_ext));    delta[6] = t[0];  delta[7] = t[1];  delta[8] = t[2];
    t = sub(a.pos_ext, b.pos_ext);                 delta[9] = t[0];  delta[10] = t[1]; delta[11] = t[2];
    t = sub(a.vel, b.vel);                         delta[12] = t[0]; delta[13] = t[1]; delta[14] = t[2];
    t = sub(a.bg, b.bg);                           delta[15] = t[0]; delta[16] = t[1]; delta[17] = t[2];
    t = sub(a.ba, b.ba);                           delta[18] = t[0]; delta[19] = t[1]; delta[20] = t[2];
    st_boxminus_g(a.g, b.g, delta + 21);
}

}  // namespace vmp
