"""Deterministic synthetic instanced scene of BASELINE.json configs[4] / SURVEY 8d config 5: `n_blas` meshes of
`grid x grid` displaced quads (2*grid^2 triangles each), one instance per mesh like the reference
(src/ray/mod.rs:122), placed on a jittered cubic lattice with random rotation and uniform scale (PCG-style
hashing seeded by 0xB200 + instance index), 8 materials cycling diffuse / rough metal / one emitter."""
import numpy as np

from . import scene as _scene


def _hash01(seed, n):
    x = (np.arange(n, dtype=np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
    x ^= x >> np.uint64(29)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(32)
    return ((x & np.uint64(0xFFFFFF)).astype(np.float64) / float(1 << 24)).astype(np.float32)


def make_scene(n_blas=1000, grid=100, seed=0xB200):
    side = int(round(n_blas ** (1.0 / 3.0)))
    while side ** 3 < n_blas:
        side += 1
    g = grid + 1
    u, v = np.meshgrid(np.linspace(-0.5, 0.5, g, dtype=np.float32), np.linspace(-0.5, 0.5, g, dtype=np.float32), indexing="ij")
    ii, jj = np.meshgrid(np.arange(grid), np.arange(grid), indexing="ij")
    a = (ii * g + jj).ravel()
    idx = np.stack([a, a + 1, a + g, a + 1, a + g + 1, a + g], axis=1).astype(np.uint32).ravel()
    materials = np.zeros((8, 12), dtype=np.float32)
    for m in range(8):
        col = 0.25 + 0.7 * _hash01(seed + 1000 + m, 3)
        materials[m, 0:3] = col
        materials[m, 3] = 1.0
        materials[m, 8] = 1.0 if m % 3 == 1 else 0.0   # metallic
        materials[m, 9] = 0.35 if m % 3 == 1 else 0.9  # roughness
    materials[7, 4:7] = 12.0  # one emitter
    meshes = []
    for b in range(n_blas):
        r = _hash01(seed + 17 * b, 16)
        # displaced grid: two sine bumps with per-mesh phase, so every BLAS has unique triangles
        h = 0.08 * np.sin((u * (3 + 4 * r[0]) + r[1]) * 6.2831853) * np.cos((v * (2 + 5 * r[2]) + r[3]) * 6.2831853)
        pos = np.stack([u, h.astype(np.float32), v], axis=-1).reshape(-1, 3)
        du = np.gradient(h, axis=0) * grid
        dv = np.gradient(h, axis=1) * grid
        nrm = np.stack([-du, np.ones_like(du), -dv], axis=-1).reshape(-1, 3)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        verts = np.zeros((g * g, 16), dtype=np.float32)
        verts[:, 0:3] = pos
        verts[:, 3] = 1.0
        verts[:, 4:8] = [0.6 + 0.4 * r[4], 0.6 + 0.4 * r[5], 0.6 + 0.4 * r[6], 1.0]
        verts[:, 8:11] = nrm
        verts[:, 11] = 1.0
        # instance transform: lattice cell + jitter, rotation about a random axis, uniform scale
        cx, cy, cz = b % side, (b // side) % side, b // (side * side)
        t = (np.array([cx, cy, cz], dtype=np.float32) + 0.5 + 0.3 * (r[7:10] - 0.5)) / side * 10.0 - 5.0
        axis = r[10:13] - 0.5
        axis /= max(np.linalg.norm(axis), 1e-6)
        ang = 6.2831853 * r[13]
        c, s_ = np.cos(ang), np.sin(ang)
        x, y, z = axis
        R = np.array([[c + x * x * (1 - c), x * y * (1 - c) - z * s_, x * z * (1 - c) + y * s_],
                      [y * x * (1 - c) + z * s_, c + y * y * (1 - c), y * z * (1 - c) - x * s_],
                      [z * x * (1 - c) - y * s_, z * y * (1 - c) + x * s_, c + z * z * (1 - c)]], dtype=np.float32)
        sc = (0.7 + 0.5 * r[14]) * 10.0 / side
        M = np.eye(4, dtype=np.float32)
        M[:3, :3] = R * sc
        M[:3, 3] = t
        transform = M.T.reshape(16).copy()  # column-major
        sec = _scene.PrimitiveSection(0, 0, g * g, 0, idx.shape[0], b % 8)
        meshes.append(_scene.Mesh("blas%d" % b, verts, idx, transform, [sec]))
    cam = _scene.Camera((1920, 1080))
    cam.look_at((9.0, 7.0, 11.0), (0.0, 0.0, 0.0), (0.0, -1.0, 0.0))
    return _scene.Scene(meshes, materials, cam)
