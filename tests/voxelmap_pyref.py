"""A second, independent reading of the reference's voxel map (voxel_plus/src/map_builder/voxel_map.cpp:11-256) in plain Python + numpy,
used ONLY to pin the C++ oracle (tests/test_oracle_independent.py): dict + OrderedDict instead of unordered_map + list, numpy.linalg.eigh
(LAPACK) instead of a hand-written 3x3 solver, matrix products by numpy.  Written from the reference's source, not from oracle/oracle.cpp.
plane->cov starts at zero (the reference leaves it uninitialised; Q7)."""
from collections import OrderedDict

import numpy as np

MERGE_ANGLE, MERGE_DIST = 0.1, 0.04          # VoxelGrid::merge_thresh_for_angle / _for_distance (voxel_map.cpp:8-9)


class Grid:
    def __init__(self, vmap, key, gid):
        self.map, self.key, self.group_id = vmap, key, gid
        self.merged = self.is_init = self.is_plane = False
        self.update_enable = True
        self.temp = []                       # (point, cov 3x3)
        self.newly = 0
        self.mean, self.ppt, self.norm, self.cov, self.n = np.zeros(3), np.zeros((3, 3)), np.zeros(3), np.zeros((6, 6)), 0
        self.center = np.zeros(3)

    def add_to_plane(self, p):                                      # :29-34
        self.mean = self.mean + (p - self.mean) / (self.n + 1.0)
        self.ppt = self.ppt + np.outer(p, p)
        self.n += 1

    def push_point(self, p, c):                                     # :42-95
        m = self.map
        if not self.is_init:
            self.add_to_plane(p); self.temp.append((p, c)); self.update_plane()
            return
        if self.is_plane and not self.update_enable:
            self.merge()
            return
        if self.update_enable:
            self.add_to_plane(p); self.temp.append((p, c)); self.newly += 1
            if self.newly >= m.update_thresh:
                self.update_plane(); self.newly = 0
            if len(self.temp) >= m.max_thresh:
                self.update_enable = False; self.temp = []

    def update_plane(self):                                         # :97-136
        m = self.map
        assert len(self.temp) == self.n
        if self.n < m.update_thresh:
            return
        self.is_init = True
        cov = self.ppt / float(self.n) - np.outer(self.mean, self.mean)
        evals, evecs = np.linalg.eigh(cov)                          # ascending, as SelfAdjointEigenSolver
        if evals[0] > m.plane_thresh:
            self.is_plane = False
            return
        self.is_plane = True
        nrm = evecs[:, 0].copy()
        JQ = np.eye(3) / float(self.n)
        for p, c in self.temp:
            F = np.zeros((3, 3))
            for k in (1, 2):
                F[k] = (p - self.mean) / (self.n * (evals[0] - evals[k])) @ (np.outer(evecs[:, k], nrm) + np.outer(nrm, evecs[:, k]))
            J = np.vstack([evecs @ F, JQ])
            self.cov = self.cov + J @ c @ J.T
        if -(self.mean @ nrm) < 0.0:
            nrm = -nrm
        self.norm = nrm
        self.center = self.mean.copy()

    def merge(self):                                                # :138-186
        x, y, z = self.key
        for k in ((x - 1, y, z), (x, y - 1, z), (x, y, z - 1), (x + 1, y, z), (x, y + 1, z), (x, y, z + 1)):
            o = self.map.feat.get(k)
            if o is None or o.group_id == self.group_id or o.update_enable or not o.is_plane:
                continue
            norm_distance = 1.0 - o.norm @ self.norm
            axis_distance = abs(o.norm @ o.mean - self.norm @ self.mean)
            if norm_distance > MERGE_ANGLE or axis_distance > MERGE_DIST:
                continue
            tn0, tm0 = np.trace(self.cov[:3, :3]), np.trace(self.cov[3:, 3:])
            tn1, tm1 = np.trace(o.cov[:3, :3]), np.trace(o.cov[3:, 3:])
            tc0, tc1 = tn0 + tm0, tn1 + tm1
            new_mean = tm0 * o.mean + tm1 * self.mean / (tm0 + tm1)            # the reference's expression, operator precedence included
            new_norm = tn0 * o.norm + tn1 * self.norm / (tn0 + tn1)
            new_cov = (tc0 * tc0 * o.cov + tc1 * tc1 * self.cov) / ((tc0 + tc1) * (tc0 + tc1))
            o.group_id = self.group_id
            self.merged = o.merged = True
            if -(new_mean @ new_norm) < 0.0:
                new_norm = -new_norm
            self.mean, self.norm, self.cov = new_mean.copy(), new_norm.copy(), new_cov.copy()
            o.mean, o.norm, o.cov = new_mean.copy(), new_norm.copy(), new_cov.copy()


class VoxelMapPy:
    def __init__(self, max_point_thresh, update_size_thresh, plane_thresh, voxel_size, capacity):
        self.max_thresh, self.update_thresh, self.plane_thresh, self.voxel_size, self.capacity = max_point_thresh, update_size_thresh, plane_thresh, voxel_size, capacity
        self.feat = {}
        self.cache = OrderedDict()           # last = most recently inserted-into (the reference's list front)
        self.count = 0
        self.evicted = []

    def index(self, p):
        return tuple(int(v) for v in np.floor(p / self.voxel_size))

    def _touch(self, p):                                            # :206-222 / :236-251
        k = self.index(p)
        g = self.feat.get(k)
        if g is None:
            g = self.feat[k] = Grid(self, k, self.count)
            self.count += 1
            self.cache[k] = True
            if len(self.cache) > self.capacity:
                old, _ = self.cache.popitem(last=False)
                del self.feat[old]
                self.evicted.append(old)
        else:
            self.cache.move_to_end(k)
        return g

    def build(self, pts, covs):
        self.evicted = []
        for p, c in zip(pts, covs):
            g = self._touch(p)
            g.add_to_plane(p); g.temp.append((p, c))
        for g in self.feat.values():
            g.update_plane()

    def update(self, pts, covs):
        self.evicted = []
        for p, c in zip(pts, covs):
            self._touch(p).push_point(p, c)
