"""Synthetic arena collision meshes in RocketSim's ``.cmf`` format.

The game's real collision meshes are not distributable (reference ``README.md:25``,
``.gitignore:35-36``) and are absent from this environment, so every
non-VOID configuration -- on BOTH the oracle and the CUDA path -- runs on these
deterministic stand-ins, fed through ``RocketSim::InitFromMem`` / ``rsb_init_from_mem``.

File format (reference ``src/CollisionMeshFile/CollisionMeshFile.cpp:11-61``), little endian:
``int32 numTris, int32 numVerts, numTris x int32[3], numVerts x float32[3]``; vertices are
already in Bullet units (1 bt = 50 uu) and used unscaled (``CollisionMeshFile.cpp:63-73``).

Geometry (uu): Soccar shell = 4 quarter-pipe ramps (R=256) along the x=+-4096 walls (floor
and ceiling), two back walls at y=+-5120 with 1786x643 goal openings, and 880-deep goal
boxes.  Hoops shell = the same ramps at Hoops extents, an elliptical rim band per hoop and a
505-vertex "net" funnel per side (the reference detects the net by that exact vertex count,
``src/Sim/Arena/Arena.cpp:1016-1019``).
"""
from __future__ import annotations

import struct
from typing import Dict, List

import numpy as np

UU_TO_BT = 1.0 / 50.0

MODE_SOCCAR, MODE_HOOPS, MODE_HEATSEEKER, MODE_SNOWDAY, MODE_DROPSHOT, MODE_THE_VOID = range(6)


class _Mesh:
    def __init__(self):
        self.verts: List[np.ndarray] = []
        self.tris: List[np.ndarray] = []
        self.nv = 0

    def add_grid(self, pts: np.ndarray, inward: np.ndarray):
        """pts: [nu, nv, 3] grid (uu); triangles wound so normals point along ``inward`` side."""
        nu, nv, _ = pts.shape
        idx = self.nv + np.arange(nu * nv).reshape(nu, nv)
        a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
        p = pts
        n = np.cross(p[1, 0] - p[0, 0], p[1, 1] - p[0, 0])
        flip = float(np.dot(n, inward)) < 0
        if flip:
            t1 = np.stack([a, c, b], -1).reshape(-1, 3)
            t2 = np.stack([a, d, c], -1).reshape(-1, 3)
        else:
            t1 = np.stack([a, b, c], -1).reshape(-1, 3)
            t2 = np.stack([a, c, d], -1).reshape(-1, 3)
        # interleave the two triangles of each quad
        tris = np.empty((t1.shape[0] * 2, 3), dtype=np.int32)
        tris[0::2] = t1
        tris[1::2] = t2
        self.verts.append(pts.reshape(-1, 3).astype(np.float32))
        self.tris.append(tris)
        self.nv += nu * nv

    def to_cmf(self) -> bytes:
        v = (np.concatenate(self.verts).astype(np.float64) * UU_TO_BT).astype(np.float32)
        t = np.concatenate(self.tris).astype(np.int32)
        return struct.pack("<ii", t.shape[0], v.shape[0]) + t.tobytes() + v.tobytes()

    @property
    def counts(self):
        return sum(t.shape[0] for t in self.tris), self.nv


def _quarter_pipes(m: _Mesh, ext_x: float, ext_y: float, height: float, R=256.0, n_arc=6, n_len=20):
    ys = np.linspace(-ext_y, ext_y, n_len + 1)
    th = np.linspace(0.0, np.pi / 2, n_arc + 1)
    for sx in (-1.0, 1.0):
        for top in (False, True):
            cx = sx * (ext_x - R)
            cz = (height - R) if top else R
            x = cx + sx * R * np.sin(th)
            z = cz + (R * np.cos(th) if top else -R * np.cos(th))
            pts = np.zeros((n_arc + 1, n_len + 1, 3))
            pts[:, :, 0] = x[:, None]
            pts[:, :, 1] = ys[None, :]
            pts[:, :, 2] = z[:, None]
            inward = np.array([-sx, 0.0, -1.0 if top else 1.0])
            m.add_grid(pts, inward)


def _wall_xz(m: _Mesh, y: float, x0: float, x1: float, z0: float, z1: float, nx: int, nz: int, inward_y: float):
    xs = np.linspace(x0, x1, nx + 1)
    zs = np.linspace(z0, z1, nz + 1)
    pts = np.zeros((nx + 1, nz + 1, 3))
    pts[:, :, 0] = xs[:, None]
    pts[:, :, 1] = y
    pts[:, :, 2] = zs[None, :]
    m.add_grid(pts, np.array([0.0, inward_y, 0.0]))


def _wall_yz(m: _Mesh, x: float, y0: float, y1: float, z0: float, z1: float, ny: int, nz: int, inward_x: float):
    ys = np.linspace(y0, y1, ny + 1)
    zs = np.linspace(z0, z1, nz + 1)
    pts = np.zeros((ny + 1, nz + 1, 3))
    pts[:, :, 0] = x
    pts[:, :, 1] = ys[:, None]
    pts[:, :, 2] = zs[None, :]
    m.add_grid(pts, np.array([inward_x, 0.0, 0.0]))


def _wall_xy(m: _Mesh, z: float, x0: float, x1: float, y0: float, y1: float, nx: int, ny: int, inward_z: float):
    xs = np.linspace(x0, x1, nx + 1)
    ys = np.linspace(y0, y1, ny + 1)
    pts = np.zeros((nx + 1, ny + 1, 3))
    pts[:, :, 0] = xs[:, None]
    pts[:, :, 1] = ys[None, :]
    pts[:, :, 2] = z
    m.add_grid(pts, np.array([0.0, 0.0, inward_z]))


GOAL_HALF_W = 893.0
GOAL_H = 643.0
GOAL_DEPTH = 880.0


def soccar_meshes() -> List[bytes]:
    """Three meshes: ramps, back walls (with goal openings), goal boxes."""
    EX, EY, H = 4096.0, 5120.0, 2048.0
    ramps = _Mesh()
    _quarter_pipes(ramps, EX, EY, H)
    walls = _Mesh()
    for sy in (-1.0, 1.0):
        y = sy * EY
        _wall_xz(walls, y, -EX, -GOAL_HALF_W, 0.0, H, 4, 4, -sy)
        _wall_xz(walls, y, GOAL_HALF_W, EX, 0.0, H, 4, 4, -sy)
        _wall_xz(walls, y, -GOAL_HALF_W, GOAL_HALF_W, GOAL_H, H, 2, 3, -sy)
    goals = _Mesh()
    for sy in (-1.0, 1.0):
        y0, y1 = sy * EY, sy * (EY + GOAL_DEPTH)
        _wall_yz(goals, -GOAL_HALF_W, y0, y1, 0.0, GOAL_H, 2, 2, 1.0)
        _wall_yz(goals, GOAL_HALF_W, y0, y1, 0.0, GOAL_H, 2, 2, -1.0)
        _wall_xz(goals, y1, -GOAL_HALF_W, GOAL_HALF_W, 0.0, GOAL_H, 2, 2, -sy)
        _wall_xy(goals, GOAL_H, -GOAL_HALF_W, GOAL_HALF_W, min(y0, y1), max(y0, y1), 2, 2, -1.0)
    return [ramps.to_cmf(), walls.to_cmf(), goals.to_cmf()]


def hoops_meshes() -> List[bytes]:
    """Ramps at Hoops extents, a rim band per hoop, and one 505-vertex net mesh."""
    EX, EY, H = 8900.0 / 3.0, 3581.0, 1820.0
    ramps = _Mesh()
    _quarter_pipes(ramps, EX, EY, H, n_len=16)
    rims = _Mesh()
    ang = np.linspace(0.0, 2 * np.pi, 25)
    for sy in (-1.0, 1.0):
        cy = sy * 2770.0 / 0.9
        zs = np.array([270.0, 320.0, 365.0])
        pts = np.zeros((25, 3, 3))
        pts[:, :, 0] = (800.0 * np.cos(ang))[:, None]
        pts[:, :, 1] = cy + (800.0 / 0.9 * np.sin(ang))[:, None]
        pts[:, :, 2] = zs[None, :]
        # outward-facing band (cars outside the hoop hit it)
        rims.add_grid(pts, np.array([1.0, 0.0, 0.0]) if pts[0, 0, 0] > 0 else np.array([-1.0, 0.0, 0.0]))
    # net: 5 rings x 101 points = 505 vertices; two funnels would be 1010, so the -y hoop only
    # carries the net mesh the reference singles out; the +y net is a second 505-vertex mesh.
    nets = []
    ang = np.linspace(0.0, 2 * np.pi, 101)
    for sy in (-1.0, 1.0):
        net = _Mesh()
        cy = sy * 2770.0 / 0.9
        radii = np.array([790.0, 700.0, 600.0, 500.0, 420.0])
        zs = np.array([265.0, 210.0, 150.0, 90.0, 40.0])
        pts = np.zeros((101, 5, 3))
        pts[:, :, 0] = np.cos(ang)[:, None] * radii[None, :]
        pts[:, :, 1] = cy + np.sin(ang)[:, None] * radii[None, :] / 0.9
        pts[:, :, 2] = zs[None, :]
        net.add_grid(pts, np.array([-1.0, 0.0, 0.0]))
        assert net.nv == 505
        nets.append(net.to_cmf())
    return [ramps.to_cmf(), rims.to_cmf()] + nets


def dropshot_meshes() -> List[bytes]:
    """Dropshot shell: the arena has no side planes (reference ``Arena.cpp:1056-1061``), so the mesh carries all six walls of
    the hexagonal field (circumradius 5400 uu: encloses the 13..7-tile rows of ``DropshotTiles.cpp``) -- a 45-degree skirt from
    the floor up to 250 uu in front of a vertical wall from the floor to the 2024 uu ceiling -- plus the solid floor strip between the two tile fields,
    in three meshes (walls / skirts / centre strip)."""
    RC, H, SK = 5400.0, 2024.0, 250.0
    ang = np.deg2rad(np.arange(7) * 60.0)
    cx, cy = RC * np.cos(ang), RC * np.sin(ang)
    walls, skirts = _Mesh(), _Mesh()
    for k in range(6):
        a = np.array([cx[k], cy[k]]); b = np.array([cx[k + 1], cy[k + 1]])
        mid = 0.5 * (a + b)
        inward = -mid / np.linalg.norm(mid)
        ts = np.linspace(0.0, 1.0, 7)
        edge = a[None, :] + ts[:, None] * (b - a)[None, :]
        # vertical part, down to the floor: the skirts of two neighbouring walls leave a wedge open at every corner, and a
        # wall that only started at the skirt's top edge would let a car slip out underneath (the reference does let it)
        zs = np.linspace(0.0, H, 5)
        pts = np.zeros((7, 5, 3))
        pts[:, :, 0] = edge[:, 0][:, None]; pts[:, :, 1] = edge[:, 1][:, None]; pts[:, :, 2] = zs[None, :]
        walls.add_grid(pts, np.array([inward[0], inward[1], 0.0]))
        # skirt: from (edge moved inward by SK) at z = 0 up to the wall foot at z = SK
        pts = np.zeros((7, 3, 3))
        for j, f in enumerate((1.0, 0.5, 0.0)):
            pts[:, j, 0] = edge[:, 0] + inward[0] * SK * f
            pts[:, j, 1] = edge[:, 1] + inward[1] * SK * f
            pts[:, j, 2] = SK * (1.0 - f)
        skirts.add_grid(pts, np.array([inward[0], inward[1], 1.0]))
    # the tile fields of the two teams start at |y| = 127.4 uu (TILE_OFFSET_Y): the strip between them is solid floor
    strip = _Mesh()
    _wall_xy(strip, 0.0, -RC, RC, -130.0, 130.0, 12, 2, 1.0)
    return [walls.to_cmf(), skirts.to_cmf(), strip.to_cmf()]


def far_triangle_mesh() -> bytes:
    """A single far-away triangle: turns any arena into a planes-only arena for staged bring-up."""
    m = _Mesh()
    pts = np.array([[[5000.0, 5900.0, 2100.0], [5010.0, 5900.0, 2100.0]],
                    [[5000.0, 5910.0, 2100.0], [5010.0, 5910.0, 2100.0]]])
    m.add_grid(pts, np.array([0.0, 0.0, 1.0]))
    return m.to_cmf()


def parse_cmf(data: bytes):
    nt, nv = struct.unpack_from("<ii", data, 0)
    tris = np.frombuffer(data, dtype="<i4", count=nt * 3, offset=8).reshape(nt, 3)
    verts = np.frombuffer(data, dtype="<f4", count=nv * 3, offset=8 + nt * 12).reshape(nv, 3)
    return tris, verts


def mesh_sets(planes_only: bool = False) -> Dict[int, List[bytes]]:
    """Mesh lists per game mode.  The Soccar list is also registered under HEATSEEKER and
    SNOWDAY, because this reference snapshot looks meshes up by the arena's own GameMode key
    (``src/RocketSim.cpp:50-53``)."""
    if planes_only:
        s = [far_triangle_mesh()]
        h = [far_triangle_mesh()]
        d = [far_triangle_mesh()]
    else:
        s = soccar_meshes()
        h = hoops_meshes()
        d = dropshot_meshes()
    return {MODE_SOCCAR: s, MODE_HOOPS: h, MODE_HEATSEEKER: s, MODE_SNOWDAY: s, MODE_DROPSHOT: d}
