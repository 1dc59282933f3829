"""Deterministic synthetic meshes (fp32 positions [V,3], uint32 faces [T,3]).

The reference ships no usable mesh (objects/bunny.obj is missing from the checkout, only a 2-triangle quad
exists), so every configuration in BASELINE.json runs on seeded generators. Trigonometry is
evaluated in float64 and rounded once to float32, so the meshes do not depend on the libm/numpy SIMD build.
"""
import numpy as np


def _grid_faces(nu, nv, wrap_u):
    """two triangles per quad of a (nv+1) x nu vertex grid (rows j, columns i); columns wrap when wrap_u"""
    j, i = np.meshgrid(np.arange(nv, dtype=np.int64), np.arange(nu if wrap_u else nu - 1, dtype=np.int64),
                       indexing="ij")
    i1 = (i + 1) % nu if wrap_u else i + 1
    a = j * nu + i
    b = j * nu + i1
    c = (j + 1) * nu + i
    d = (j + 1) * nu + i1
    f = np.stack([np.stack([a, b, c], -1), np.stack([b, d, c], -1)], axis=2)  # [nv, ni, 2, 3]
    return f.reshape(-1, 3).astype(np.uint32)


def uv_sphere(nu, nv=None, radius=1.0, center=(0.0, 0.0, 0.0)):
    """nu columns x nv rows UV sphere, T = 2*nu*nv, V = (nv+1)*nu (pole rows are duplicated vertices, so the
    pole triangles are zero-area on purpose: degenerate input the narrow phase must treat like the reference)."""
    nv = nu if nv is None else nv
    th = np.pi * np.arange(nv + 1, dtype=np.float64) / nv
    ph = 2.0 * np.pi * np.arange(nu, dtype=np.float64) / nu
    st, ct = np.sin(th)[:, None], np.cos(th)[:, None]
    x = st * np.cos(ph)[None, :]
    y = np.broadcast_to(ct, x.shape)
    z = st * np.sin(ph)[None, :]
    pos = np.stack([x, y, z], -1).reshape(-1, 3) * radius + np.asarray(center, np.float64)
    return pos.astype(np.float32), _grid_faces(nu, nv, True)


def blob(nu, nv=None, seed=1234, amp=0.15, radius=1.0, center=(0.0, 0.0, 0.0)):
    """bunny stand-in: UV sphere whose radius is modulated by a few seeded low-frequency lobes.
    T = 2*nu*nv (nu=1024, nv=512 -> 2^20 faces)."""
    nv = nu if nv is None else nv
    rng = np.random.default_rng(seed)
    th = np.pi * np.arange(nv + 1, dtype=np.float64) / nv
    ph = 2.0 * np.pi * np.arange(nu, dtype=np.float64) / nu
    TH, PH = np.meshgrid(th, ph, indexing="ij")
    d = np.stack([np.sin(TH) * np.cos(PH), np.cos(TH), np.sin(TH) * np.sin(PH)], -1)
    r = np.ones_like(TH)
    for _ in range(6):
        k = rng.normal(size=3)
        k /= np.linalg.norm(k)
        freq = rng.integers(2, 7)
        phase = rng.uniform(0, 2 * np.pi)
        r += (amp / 6.0) * np.sin(freq * np.arccos(np.clip(d @ k, -1, 1)) + phase)
    pos = d * (radius * r)[..., None]
    pos = pos.reshape(-1, 3) + np.asarray(center, np.float64)
    return pos.astype(np.float32), _grid_faces(nu, nv, True)


def icosphere(subdiv, radius=1.0, center=(0.0, 0.0, 0.0)):
    """closed genus-0 mesh, T = 20*4^subdiv, shared vertices"""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11],
                  [6, 2, 10], [8, 6, 7], [9, 8, 1]], np.int64)
    for _ in range(subdiv):
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)
        es = np.sort(e, axis=1)
        uniq, inv = np.unique(es, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        mid = v[uniq[:, 0]] + v[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = len(v)
        v = np.concatenate([v, mid], 0)
        n = len(f)
        m01, m12, m20 = base + inv[:n], base + inv[n:2 * n], base + inv[2 * n:]
        f = np.concatenate([np.stack([f[:, 0], m01, m20], 1), np.stack([f[:, 1], m12, m01], 1),
                            np.stack([f[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)], 0)
    pos = v * radius + np.asarray(center, np.float64)
    return pos.astype(np.float32), f.astype(np.uint32)


def terrain(nx, ny, seed=1234, size=(8.0, 8.0), height=0.4, octaves=5):
    """height field over an (ny+1) x (nx+1) vertex grid, T = 2*nx*ny, y up; seeded sum-of-sines fBm"""
    rng = np.random.default_rng(seed)
    xs = np.linspace(-size[0] / 2, size[0] / 2, nx + 1)
    zs = np.linspace(-size[1] / 2, size[1] / 2, ny + 1)
    Z, X = np.meshgrid(zs, xs, indexing="ij")
    h = np.zeros_like(X)
    for o in range(octaves):
        ang = rng.uniform(0, 2 * np.pi)
        fr = (2.0 ** o) * 0.8
        h += (0.5 ** o) * np.sin(fr * (np.cos(ang) * X + np.sin(ang) * Z) + rng.uniform(0, 2 * np.pi))
    pos = np.stack([X, height * h, Z], -1).reshape(-1, 3)
    return pos.astype(np.float32), _grid_faces(nx + 1, ny, False)


def cloth_positions(base_pos, frame, amp=0.05, freq=3.0):
    """per-frame sinusoidal deformation of a base mesh (config 3: refit only, topology fixed)"""
    p = np.asarray(base_pos, np.float64)
    d = amp * np.sin(freq * p[:, [1, 2, 0]] + 0.37 * frame)
    return (p + d).astype(np.float32)


def cube(half=0.5, center=(0.0, 0.0, 0.0)):
    """12-triangle box"""
    s = half
    v = np.array([[-s, -s, s], [s, -s, s], [s, s, s], [-s, s, s], [-s, -s, -s], [s, -s, -s], [s, s, -s], [-s, s, -s]],
                 np.float64) + np.asarray(center, np.float64)
    f = np.array([[0, 1, 2], [0, 2, 3], [1, 5, 6], [1, 6, 2], [5, 4, 7], [5, 7, 6], [4, 0, 3], [4, 3, 7], [3, 2, 6],
                  [3, 6, 7], [4, 5, 1], [4, 1, 0]], np.uint32)
    return v.astype(np.float32), f


def quad():
    """the reference's objects/cube.obj content: a 2-triangle planar quad in z = 0 (objects/cube.obj:8-15)"""
    v = np.array([[0.5, 0.5, 0], [0.5, -0.5, 0], [-0.5, -0.5, 0], [-0.5, 0.5, 0]], np.float32)
    f = np.array([[0, 1, 3], [1, 2, 3]], np.uint32)
    return v, f


def shuffle_faces(faces, seed=7):
    """arbitrary input order (asset pipelines do not emit spatially sorted faces)"""
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(faces[rng.permutation(len(faces))])


def truncate_faces(pos, faces, T):
    """first T faces (to hit non-power-of-two / odd primitive counts)"""
    return pos, np.ascontiguousarray(faces[:T])
