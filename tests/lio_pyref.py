"""A free-running SECOND implementation of the whole per-scan path in Python + numpy / scipy, composed of the restatements that pin the oracle piece
by piece (tests/test_oracle_math.py, tests/test_oracle_independent.py, tests/voxelmap_pyref.py): LIOBuilder::process (lio_builder.cpp:175-248) =
initializeImu, undistortCloud with IESKF::predict (state AND covariance), calcBodyCov, IESKF::update around sharedUpdateFunc / buildResidual with the
residual records that persist across iterations and scans, lidarToWorld + pv_list, VoxelMap::build / update.  Test infrastructure only."""
import numpy as np

from test_oracle_independent import _LioNumpy, _body_cov_numpy, _hat as _hat3
from test_oracle_math import _Bx, _Nx, _boxminus, _boxplus, _exp, _hat, _jac, _right_jac
from voxelmap_pyref import VoxelMapPy


class LioPy:
    IMU_INIT, MAP_INIT, LIO_MAPPING = 0, 1, 2

    def __init__(self, cfg):
        self.cfg = cfg
        self.status = self.IMU_INIT
        self.pre = _LioNumpy(cfg)
        self.map = VoxelMapPy(cfg.max_point_thresh, cfg.update_size_thresh, cfg.plane_thresh, cfg.voxel_size, cfg.map_capacity)
        self.Q = np.zeros((12, 12))
        self.Q[0:3, 0:3] = np.eye(3) * cfg.ng; self.Q[3:6, 3:6] = np.eye(3) * cfg.na
        self.Q[6:9, 6:9] = np.eye(3) * cfg.nbg; self.Q[9:12, 9:12] = np.eye(3) * cfg.nba
        n = cfg.max_points_per_scan
        self.rec_valid = np.zeros(n, bool); self.rec_nrm = np.zeros((n, 3)); self.rec_res = np.zeros(n)       # ResidualData persists (lio_builder.h)
        self.x, self.P = None, None
        self.iters, self.effect = 0, []

    # IESKF::predict with the covariance (ieskf.cpp:101-123); the state part is the one _LioNumpy.undistort applies
    def _predict_P(self, x, acc, gyro, dt):
        w, a = gyro - x["bg"], acc - x["ba"]
        F = np.eye(23)
        F[0:3, 12:15] = np.eye(3) * dt
        F[3:6, 3:6] = _exp(-w * dt)
        F[3:6, 15:18] = -_right_jac(w * dt) * dt
        F[12:15, 3:6] = -x["rot"] @ _hat(a) * dt
        F[12:15, 18:21] = -x["rot"] * dt
        Mx = -_hat(x["g"]) @ _Bx(x["g"])
        F[12:15, 21:23] = Mx * dt
        F[21:23, 21:23] = _Nx(x["g"]) @ Mx
        G = np.zeros((23, 12))
        G[3:6, 0:3] = -_right_jac(w * dt) * dt
        G[12:15, 3:6] = -x["rot"] * dt
        G[15:18, 6:9] = np.eye(3) * dt
        G[18:21, 9:12] = np.eye(3) * dt
        self.P = F @ self.P @ F.T + G @ self.Q @ G.T

    def _propagate(self, imus, cloud, t0, t1):
        """undistortCloud: same walk over the IMU samples as _LioNumpy.undistort, with the covariance carried along"""
        pre, x = self.pre, self.x
        cache = [pre.last_imu] + [(np.array(i["acc"]), np.array(i["gyro"]), float(i["timestamp"])) for i in imus]
        xs = {k: v.copy() for k, v in x.items()}
        acc = gyro = None
        for head, tail in zip(cache[:-1], cache[1:]):                # covariance first (it needs the state BEFORE each step) ...
            if tail[2] < pre.last_end:
                continue
            gyro = 0.5 * (head[1] + tail[1]); acc = 0.5 * (head[0] + tail[0]) * 9.81 / pre.gravity_norm
            dt = tail[2] - pre.last_end if head[2] < pre.last_end else tail[2] - head[2]
            self._predict_P(xs, acc, gyro, dt)
            w, a = gyro - xs["bg"], acc - xs["ba"]
            xs["pos"], xs["vel"], xs["rot"] = xs["pos"] + xs["vel"] * dt, xs["vel"] + (xs["rot"] @ a + xs["g"]) * dt, xs["rot"] @ _exp(w * dt)
        self._predict_P(xs, acc, gyro, t1 - cache[-1][2])
        return pre.undistort(x, imus, cloud, t0, t1)                 # ... then states, pose list and compensation (advances x in place)

    def _pv_list(self, cloud):
        x, P = self.x, self.P
        r_wl, p_wl = x["rot"] @ x["rot_ext"], x["rot"] @ x["pos_ext"] + x["pos"]
        m, t = r_wl.astype(np.float32), p_wl.astype(np.float32)
        px, py, pz = cloud[:, 0], cloud[:, 1], cloud[:, 2]
        w = np.stack([m[r, 0] * px + (m[r, 1] * py + (m[r, 2] * pz + t[r])) for r in range(3)], 1).astype(np.float64)
        covs = np.zeros((len(cloud), 3, 3))
        for i in range(len(cloud)):
            pl, cl = _body_cov_numpy(cloud[i, :3].astype(np.float64))
            covs[i] = r_wl @ cl @ r_wl.T + _hat3(pl) @ P[3:6, 3:6] @ _hat3(pl).T + P[0:3, 0:3]
        return w, covs

    def _measure(self, x, pls, cls):
        """sharedUpdateFunc + buildResidual on the persistent records"""
        cfg, P = self.cfg, self.P
        R, Rext, pext = x["rot"], x["rot_ext"], x["pos_ext"]
        r_wl, p_wl = R @ Rext, R @ pext + x["pos"]
        H = np.zeros((12, 12)); b = np.zeros(12); eff = 0
        for i, (pl, cl) in enumerate(zip(pls, cls)):
            pw = r_wl @ pl + p_wl
            g = self.map.feat.get(self.map.index(pw))
            if g is not None:
                self.rec_valid[i] = False
                if g.is_plane:
                    cw = r_wl @ cl @ r_wl.T + _hat3(pl) @ P[3:6, 3:6] @ _hat3(pl).T + P[0:3, 0:3]
                    self.rec_nrm[i] = g.norm
                    self.rec_res[i] = g.norm @ (pw - g.mean)
                    self.rec_valid[i] = abs(self.rec_res[i]) < 3.0 * np.sqrt(g.norm @ cw @ g.norm)
            if not self.rec_valid[i]:
                continue
            eff += 1
            nrm, res = self.rec_nrm[i], self.rec_res[i]
            r_cov = nrm @ r_wl @ cl @ r_wl.T @ nrm
            wgt = 5000.0 if r_cov < 0.0002 else 1.0 / r_cov
            J = np.zeros(12)
            J[0:3] = nrm
            J[3:6] = -nrm @ R @ _hat3(Rext @ pl + pext)
            if cfg.estimate_ext:
                J[6:9] = -nrm @ r_wl @ _hat3(pl)
                J[9:12] = nrm @ R
            H += np.outer(J, J) * wgt
            b += J * wgt * res
        return H, b, eff

    def _update(self, cloud):
        """IESKF::update (ieskf.cpp:125-156); max_iter = opti_max_iter, eps = 0.001"""
        body = [_body_cov_numpy(p.astype(np.float64)) for p in cloud[:, :3]]
        pls, cls = [b[0] for b in body], [b[1] for b in body]
        x_pred = {k: v.copy() for k, v in self.x.items()}
        Pinv = np.linalg.inv(self.P)
        x = self.x
        self.iters, self.effect = 0, []
        for _ in range(self.cfg.opti_max_iter):
            Hm, bm, eff = self._measure(x, pls, cls)
            self.effect.append(eff)
            delta = _boxminus(x, x_pred)
            J = _jac(x, x_pred, delta)
            b_ = J.T @ Pinv @ delta
            H_ = J.T @ Pinv @ J
            H_[:12, :12] += Hm; b_[:12] += bm
            step = -np.linalg.solve(H_, b_)
            x = _boxplus(x, step)
            self.iters += 1
            if step.max() < 0.001:
                break
        L = _jac(x, x_pred, step)
        self.x, self.P = x, L @ np.linalg.inv(H_) @ L.T

    def process(self, imus, cloud, t0, t1):
        """cloud: n x 4 float32 (x y z, time offset in ms), in time order.  Returns the compensated cloud (None while the IMU initialises)."""
        if self.status == self.IMU_INIT:
            x = self.pre.initialize(imus, t1)
            if x is not None:
                self.x = x
                self.P = np.eye(23)
                self.P[6:9, 6:9] = np.eye(3) * 0.00001; self.P[9:12, 9:12] = np.eye(3) * 0.00001
                self.P[15:18, 15:18] = np.eye(3) * 0.0001; self.P[18:21, 18:21] = np.eye(3) * 0.0001
                self.P[21:23, 21:23] = np.eye(2) * 0.00001
                self.status = self.MAP_INIT
            return None
        comp = self._propagate(imus, cloud, t0, t1)
        if self.status == self.MAP_INIT:
            w, covs = self._pv_list(comp)
            self.map.build(w, covs)
            self.status = self.LIO_MAPPING
            return comp
        lidar = comp
        if self.cfg.scan_resolution > 0.0:                           # scan_filter.filter(*lidar_cloud), lio_builder.cpp:215-219 (MAP_INIT builds from the unfiltered cloud)
            from test_downsample import np_voxel_grid
            lidar = np_voxel_grid(comp, self.cfg.scan_resolution)
        self._update(lidar)
        w, covs = self._pv_list(lidar)
        self.map.update(w, covs)
        return comp
