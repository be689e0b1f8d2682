"""Synthetic driver that replaces the reference's ROS node for measurement and tests.

The reference is fed by a Livox Mid-360 through `lio_node.cpp` (`livox2pcl`, utils.cpp:3-28:
float32 xyz + per-point time offset in ms stored in `curvature`; `syncPackage`,
lio_node.cpp:137-165: one cloud + every IMU sample before `cloud_end_time`).  ROS is out of
scope (BASELINE.json north_star), so this module produces the same `SyncPackage` contents
procedurally: a Mid-360-style non-repetitive rosette scan of a planar scene, motion-distorted
along an analytic trajectory, with a 200 Hz IMU stream.  Everything is deterministic in
(seed, scan index) through counter-based Philox streams, so the CPU oracle and the CUDA path
see bit-identical inputs.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .ctypes_defs import IMU_DTYPE

G_WORLD = np.array([0.0, 0.0, -9.81])


# --------------------------------------------------------------------------- scene
@dataclass
class Scene:
    """A set of finite rectangles: centre c, unit normal n, in-plane unit axes u, v, half extents hu, hv."""
    c: np.ndarray
    n: np.ndarray
    u: np.ndarray
    v: np.ndarray
    hu: np.ndarray
    hv: np.ndarray

    @staticmethod
    def from_list(rects):
        c, n, u, v, hu, hv = [], [], [], [], [], []
        for (cc, uu, vv, a, b) in rects:
            uu = np.asarray(uu, float) / np.linalg.norm(uu)
            vv = np.asarray(vv, float)
            vv = vv - uu * (uu @ vv)
            vv /= np.linalg.norm(vv)
            c.append(np.asarray(cc, float)); u.append(uu); v.append(vv); n.append(np.cross(uu, vv)); hu.append(a); hv.append(b)
        return Scene(np.array(c), np.array(n), np.array(u), np.array(v), np.array(hu, float), np.array(hv, float))

    def near(self, o: np.ndarray, rmax: float) -> "Scene":
        """The rectangles that rays starting at the origins `o` can reach within rmax (bounding-sphere test).  Only used
        when a caller asks for culling (large scenes: C3 / C4 tools); the default path casts against every rectangle."""
        ctr = o.mean(axis=0)
        spread = float(np.linalg.norm(o - ctr, axis=1).max())
        keep = np.linalg.norm(self.c - ctr, axis=1) - np.sqrt(self.hu ** 2 + self.hv ** 2) <= rmax + spread + 1e-6
        return Scene(self.c[keep], self.n[keep], self.u[keep], self.v[keep], self.hu[keep], self.hv[keep])

    def cast(self, o: np.ndarray, d: np.ndarray, rmin: float, rmax: float) -> np.ndarray:
        """Range of the first hit of rays o + t d (N x 3 each); inf where nothing is hit in [rmin, rmax]."""
        denom = d @ self.n.T                                    # N x S
        num = np.einsum("sk,sk->s", self.c, self.n)[None, :] - o @ self.n.T
        with np.errstate(divide="ignore", invalid="ignore"):
            t = num / denom
        t[~np.isfinite(t)] = np.inf
        t[(t < rmin) | (t > rmax)] = np.inf
        best = np.full(o.shape[0], np.inf)
        # test in-rectangle only for candidate hits, surface by surface (S is small)
        for s in range(self.c.shape[0]):
            ts = t[:, s]
            m = ts < best
            if not m.any():
                continue
            hit = o[m] + ts[m, None] * d[m] - self.c[s]
            inside = (np.abs(hit @ self.u[s]) <= self.hu[s]) & (np.abs(hit @ self.v[s]) <= self.hv[s])
            idx = np.nonzero(m)[0][inside]
            best[idx] = ts[idx]
        return best


def _box_faces(lo, hi, inward=False, skip_bottom=False):
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    ctr, h = (lo + hi) / 2, (hi - lo) / 2
    faces = []
    ex = np.eye(3)
    for ax in range(3):
        a1, a2 = (ax + 1) % 3, (ax + 2) % 3
        for sgn in (-1, 1):
            if skip_bottom and ax == 2 and sgn == -1:
                continue
            cc = ctr.copy()
            cc[ax] += sgn * h[ax]
            faces.append((cc, ex[a1], ex[a2], h[a1], h[a2]))
    return faces


def scene_room(seed: int = 7, clutter: int = 14) -> Scene:
    """Scene A: 40 x 30 x 6 m hall, partitions, a pillar, tilted panels, box clutter."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    r = _box_faces([-20, -15, 0], [20, 15, 6])
    r += _box_faces([-2, -2, 0], [2, 2, 6], skip_bottom=True)[:4]                 # central pillar (4 side faces)
    # partitions attached to the outer walls (the trajectory stays inside |x|<13, |y|<9)
    for (x, y0, y1) in ((-10, -15, -10.5), (10, 10.5, 15), (4, -15, -11), (-5, 11, 15)):
        r.append(((x, (y0 + y1) / 2, 3), (0, 1, 0), (0, 0, 1), abs(y1 - y0) / 2, 3))
    for (y, x0, x1) in ((6, -20, -15), (-7, 15.5, 20)):
        r.append((((x0 + x1) / 2, y, 3), (1, 0, 0), (0, 0, 1), abs(x1 - x0) / 2, 3))
    # tilted panels (10..30 degrees off vertical / horizontal)
    for (c, yaw, tilt, a, b) in (((16, 3, 2.2), 0.2, 20, 3.0, 2.0), ((-16, -4, 2.5), 2.9, 25, 3.5, 2.2),
                                 ((3, 12.5, 2.0), 1.7, 15, 4.0, 1.8), ((-6, -12.5, 2.4), -1.5, 30, 3.0, 2.0),
                                 ((0, 0, 5.2), 0.5, 10, 6.0, 5.0)):
        t = np.deg2rad(tilt)
        u = np.array([-np.sin(yaw), np.cos(yaw), 0.0])
        if c[2] > 5:    # near-horizontal ceiling panel
            v = np.array([np.cos(yaw) * np.cos(t), np.sin(yaw) * np.cos(t), np.sin(t)])
        else:
            v = np.array([np.cos(yaw) * np.sin(t), np.sin(yaw) * np.sin(t), np.cos(t)])
        r.append((c, u, v, a, b))
    # box clutter on the floor, away from the trajectory ellipse
    for _ in range(clutter):
        while True:
            x, y = rng.uniform(-18, 18), rng.uniform(-13, 13)
            e = (x / 12.0) ** 2 + (y / 8.0) ** 2
            if (e > 1.9 or e < 0.35) and not (abs(x) < 3.5 and abs(y) < 3.5):
                break
        sx, sy, sz = rng.uniform(0.4, 1.5), rng.uniform(0.4, 1.5), rng.uniform(0.5, 2.5)
        r += _box_faces([x - sx, y - sy, 0], [x + sx, y + sy, sz], skip_bottom=True)
    return Scene.from_list(r)


def scene_city(blocks: int = 6, block: float = 200.0, street: float = 24.0, height: float = 30.0, pilasters: bool = False) -> Scene:
    """Scene B: Manhattan grid of facade planes + ground; streets between blocks.
    pilasters=True adds shallow boxes (0.6-1.0 m deep, ~1 m wide, 8-12 m high, every ~9.5 m) on both facades of the street
    y = blocks/2 * block, the one tools/bench_c3.py drives: between cross streets the bare scene is two parallel planes and the
    ground, which leaves the along-street direction unobservable for any point-to-plane matcher (the reference's estimator
    picks up a spurious velocity in the first scans and never loses it)."""
    r = []
    ext = blocks * block
    r.append(((ext / 2, ext / 2, 0.0), (1, 0, 0), (0, 1, 0), ext, ext))          # ground
    for i in range(blocks):
        for j in range(blocks):
            lo = [i * block + street / 2, j * block + street / 2, 0]
            hi = [(i + 1) * block - street / 2, (j + 1) * block - street / 2, height]
            r += _box_faces(lo, hi, skip_bottom=True)[:4]
    if pilasters:
        y0 = (blocks // 2) * block
        x, k = 6.0, 0
        while x < ext - 6.0:
            for side in (-1.0, 1.0):
                yf = y0 + side * street / 2                      # facade plane
                d = 0.6 + 0.2 * ((k + (side > 0)) % 3)
                w = 1.0 + 0.3 * (k % 2)
                r += _box_faces([x - w / 2, min(yf, yf - side * d), 0], [x + w / 2, max(yf, yf - side * d), 8.0 + 2.0 * (k % 3)], skip_bottom=True)
            x += 8.0 + 1.5 * (k % 3)
            k += 1
    return Scene.from_list(r)


# --------------------------------------------------------------------------- trajectory
def _euler_zyx(yaw, pitch, roll):
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    R = np.empty(yaw.shape + (3, 3))
    R[..., 0, 0] = cy * cp; R[..., 0, 1] = cy * sp * sr - sy * cr; R[..., 0, 2] = cy * sp * cr + sy * sr
    R[..., 1, 0] = sy * cp; R[..., 1, 1] = sy * sp * sr + cy * cr; R[..., 1, 2] = sy * sp * cr - cy * sr
    R[..., 2, 0] = -sp;     R[..., 2, 1] = cp * sr;                R[..., 2, 2] = cp * cr
    return R


@dataclass
class Trajectory:
    """Analytic body trajectory: static for `t_static` seconds, then eased onto a closed curve.

    kind="ellipse": radii (ax, ay) around `centre`, period `period`; kind="loop": rounded-rectangle-like
    Lissajous through a city grid (C3).  Body frame = IMU frame; yaw follows the motion."""
    kind: str = "ellipse"
    centre: tuple = (0.0, 0.0, 1.5)
    ax: float = 12.0
    ay: float = 8.0
    period: float = 60.0
    t_static: float = 2.0
    ease: float = 1.0
    wobble: float = 0.04

    def _tau(self, t):
        s = np.maximum(t - self.t_static, 0.0)
        k = self.ease
        e = np.exp(-k * s)
        tau = s - (1.0 - e) / k
        d1 = np.where(t > self.t_static, 1.0 - e, 0.0)
        d2 = np.where(t > self.t_static, k * e, 0.0)
        return tau, d1, d2

    def eval(self, t):
        """-> pos (..,3), R (..,3,3), vel world, acc world, omega body."""
        t = np.asarray(t, float)
        tau, d1, d2 = self._tau(t)
        w = 2 * np.pi / self.period
        c = np.array(self.centre)
        # position on the curve as a function of tau, with first/second derivatives w.r.t. tau
        p = np.stack([c[0] + self.ax * np.cos(w * tau), c[1] + self.ay * np.sin(w * tau),
                      c[2] + 0.3 * np.sin(2 * w * tau)], -1)
        p1 = np.stack([-self.ax * w * np.sin(w * tau), self.ay * w * np.cos(w * tau), 0.6 * w * np.cos(2 * w * tau)], -1)
        p2 = np.stack([-self.ax * w * w * np.cos(w * tau), -self.ay * w * w * np.sin(w * tau),
                       -1.2 * w * w * np.sin(2 * w * tau)], -1)
        vel = p1 * d1[..., None]
        acc = p2 * (d1 ** 2)[..., None] + p1 * d2[..., None]
        # attitude: yaw leads the motion, small roll/pitch wobble
        a = self.wobble
        yaw, yaw1 = w * tau + np.pi / 2 + 0.2 * np.sin(3 * w * tau), w + 0.6 * w * np.cos(3 * w * tau)
        pit, pit1 = a * np.sin(5 * w * tau), 5 * a * w * np.cos(5 * w * tau)
        rol, rol1 = a * np.sin(7 * w * tau + 0.3) - a * np.sin(0.3), 7 * a * w * np.cos(7 * w * tau + 0.3)
        R = _euler_zyx(yaw, pit, rol)
        yd, pd, rd = yaw1 * d1, pit1 * d1, rol1 * d1
        om = np.stack([rd - yd * np.sin(pit),
                       pd * np.cos(rol) + yd * np.sin(rol) * np.cos(pit),
                       -pd * np.sin(rol) + yd * np.cos(rol) * np.cos(pit)], -1)
        return p, R, vel, acc, om


# --------------------------------------------------------------------------- sensor + sequence
@dataclass
class SensorConfig:
    pts_per_scan: int = 20000
    scan_period: float = 0.1
    imu_rate: float = 200.0
    range_min: float = 0.5
    range_max: float = 40.0
    range_noise: float = 0.02
    bearing_noise_deg: float = 0.05
    gyro_noise: float = 1e-3
    acc_noise: float = 1e-2
    gyro_bias: tuple = (0.002, -0.001, 0.0015)
    acc_bias: tuple = (0.01, -0.02, 0.015)
    f_az: float = 317.3
    f_el: float = 73.19
    el_mid_deg: float = 22.5
    el_amp_deg: float = 29.5
    r_il: np.ndarray = field(default_factory=lambda: np.eye(3))
    p_il: np.ndarray = field(default_factory=lambda: np.zeros(3))


@dataclass
class Package:
    """The contents of lio::SyncPackage (commons.h:22-28) for one scan."""
    index: int
    imus: np.ndarray            # IMU_DTYPE
    cloud: np.ndarray           # N x 4 float32: x y z curvature[ms]
    t0: float
    t1: float
    gt_pos: np.ndarray          # ground truth body pose at t1
    gt_rot: np.ndarray


class Sequence:
    """Deterministic stream of SyncPackages. `package(i)` is a pure function of (seed, i)."""

    def __init__(self, scene: Scene | None = None, traj: Trajectory | None = None,
                 sensor: SensorConfig | None = None, seed: int = 0xC0FFEE, cull: bool = False):
        self.cull = cull                  # cast only against the rectangles within range of the scan (large scenes)
        self.scene = scene if scene is not None else scene_room()
        self.traj = traj if traj is not None else Trajectory()
        self.sensor = sensor if sensor is not None else SensorConfig()
        self.seed = int(seed)

    def _rng(self, stream: int, index: int):
        return np.random.Generator(np.random.Philox(key=self.seed + (stream << 40), counter=[index, 0, 0, 0]))

    def imu_between(self, k0: int, k1: int) -> np.ndarray:
        """IMU samples k0 <= k < k1 at t = k / imu_rate."""
        s = self.sensor
        k = np.arange(k0, k1)
        t = k / s.imu_rate
        _, R, _, acc, om = self.traj.eval(t)
        f_body = np.einsum("nji,nj->ni", R, acc - G_WORLD)            # R^T (a - g)
        out = np.zeros(k.shape[0], IMU_DTYPE)
        for j, kk in enumerate(k):
            rng = self._rng(1, int(kk))
            out["acc"][j] = f_body[j] + np.array(s.acc_bias) + rng.normal(0, s.acc_noise, 3)
            out["gyro"][j] = om[j] + np.array(s.gyro_bias) + rng.normal(0, s.gyro_noise, 3)
        out["timestamp"] = t
        return out

    def cloud(self, index: int):
        """Motion-distorted scan `index` in the lidar frame: N x 4 float32 (x,y,z,curvature ms)."""
        s = self.sensor
        n = s.pts_per_scan
        t0 = index * s.scan_period
        off = (np.arange(n) + 0.5) * (s.scan_period / n)            # seconds, strictly increasing
        t = t0 + off
        rng = self._rng(2, index)
        az = 2 * np.pi * np.mod(s.f_az * t + 0.37 * index, 1.0)
        el = np.deg2rad(s.el_mid_deg + s.el_amp_deg * np.sin(2 * np.pi * s.f_el * t + 0.11 * index))
        d_l = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], -1)
        p, R, _, _, _ = self.traj.eval(t)
        R_wl = R @ s.r_il
        o = p + R @ s.p_il
        d_w = np.einsum("nij,nj->ni", R_wl, d_l)
        rng_m = (self.scene.near(o, s.range_max) if self.cull else self.scene).cast(o, d_w, s.range_min, s.range_max)
        ok = np.isfinite(rng_m)
        # measurement noise: range + small bearing perturbation
        rr = rng_m + rng.normal(0, s.range_noise, n)
        b = np.deg2rad(s.bearing_noise_deg)
        d_n = d_l + rng.normal(0, b, (n, 3))
        d_n /= np.linalg.norm(d_n, axis=1, keepdims=True)
        pts = d_n * rr[:, None]
        cloud = np.empty((int(ok.sum()), 4), np.float32)
        cloud[:, :3] = pts[ok].astype(np.float32)
        cloud[:, 3] = (off[ok] * 1000.0).astype(np.float32)
        return cloud, t0

    def packages(self, n: int, start: int = 0, clouds: dict | None = None):
        """Generator over packages start..start+n-1 (keeps the IMU hand-over consistent and cheap).
        `clouds` may hold precomputed `cloud(i)` results (index -> (cloud, t0)), e.g. made by a process pool."""
        s = self.sensor
        prev_end = None
        for i in range(start, start + n):
            cloud, t0 = clouds[i] if clouds is not None and i in clouds else self.cloud(i)
            t1 = t0 + float(cloud[-1, 3]) / 1000.0
            if prev_end is None:
                if i == 0:
                    k0 = 0
                else:
                    pc, pt0 = self.cloud(i - 1)
                    k0 = int(np.ceil((pt0 + float(pc[-1, 3]) / 1000.0) * s.imu_rate - 1e-12))
            else:
                k0 = int(np.ceil(prev_end * s.imu_rate - 1e-12))
            k1 = int(np.ceil(t1 * s.imu_rate - 1e-12))
            imus = self.imu_between(k0, k1)
            p, R, _, _, _ = self.traj.eval(np.array([t1]))
            prev_end = t1
            yield Package(i, imus, cloud, t0, t1, p[0], R[0])


def rot_angle_deg(Ra: np.ndarray, Rb: np.ndarray) -> float:
    c = (np.trace(Ra.T @ Rb) - 1.0) / 2.0
    return float(np.degrees(np.arccos(np.clip(c, -1.0, 1.0))))
