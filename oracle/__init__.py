"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the CPU oracle.

* ``port``  -- oracle/liboibvh_oracle.so, the C restatement (oracle/oibvh_oracle.c) of the reference path.
* ``ref``   -- oracle/_ref/liboibvh_ref.so, the UNMODIFIED reference CPU classes (SimpleBVH/SimpleCollide/Mesh)
               behind oracle/ref_shim.cpp; exists only after ``make -C oracle ref`` ran where /root/reference is.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs import this
package. The product package ``oibvh_b200`` must never do so.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_PATH = os.path.join(_HERE, "liboibvh_oracle.so")
REF_PATH = os.path.join(_HERE, "_ref", "liboibvh_ref.so")
# the reference CPU classes with `int a[19]` -> `int a[64]` (oracle/Makefile target refdeep): timing arm of bench.py only
REF_DEEP_PATH = os.path.join(_HERE, "_ref", "liboibvh_ref_deep.so")

u32p = C.POINTER(C.c_uint32)
f32p = C.POINTER(C.c_float)
u64p = C.POINTER(C.c_uint64)


def build(ref=True):
    """(Re)build the oracle libraries; `ref` only where /root/reference exists."""
    subprocess.check_call(["make", "-s", "-C", _HERE, os.path.join(_HERE, "liboibvh_oracle.so")])
    if ref and os.path.isdir("/root/reference/src/cpu"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(f32p)


def _u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(u32p)


# =====================================================================================================
# port
# =====================================================================================================
class Port:
    def __init__(self, path=PORT_PATH):
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        for name in ("orc_next_pow2", "orc_ilog2", "orc_get_size"):
            getattr(L, name).restype = C.c_uint32
            getattr(L, name).argtypes = [C.c_uint32]
        for name in ("orc_implicit_to_real", "orc_real_to_implicit", "orc_most_right_valid", "orc_level_real_count"):
            getattr(L, name).restype = C.c_uint32
            getattr(L, name).argtypes = [C.c_uint32] * 3
        L.orc_have_rchild.restype = C.c_int
        L.orc_have_rchild.argtypes = [C.c_uint32] * 3
        L.orc_mesh_aabb.argtypes = [f32p, C.c_uint32, f32p]
        L.orc_leaf_aabbs.argtypes = [f32p, u32p, C.c_uint32, f32p]
        L.orc_morton_keys.argtypes = [f32p, u32p, C.c_uint32, f32p, u32p]
        L.orc_stable_sort_perm.argtypes = [u32p, C.c_uint32, u32p]
        L.orc_tree_from_faces.argtypes = [f32p, u32p, C.c_uint32, f32p]
        L.orc_build.argtypes = [f32p, u32p, C.c_uint32, f32p, f32p, u32p, u32p, u32p]
        L.orc_tri_tri.restype = C.c_int
        L.orc_tri_tri.argtypes = [f32p, f32p]
        L.orc_aabb_overlap.restype = C.c_int
        L.orc_aabb_overlap.argtypes = [f32p, f32p]
        L.orc_detect.restype = C.c_uint64
        L.orc_detect.argtypes = [C.c_uint32, C.POINTER(f32p), C.POINTER(u32p), C.POINTER(f32p), u32p, u32p,
                                 C.c_uint64, u64p, u64p]
        L.orc_candidates_bruteforce.restype = C.c_uint64
        L.orc_candidates_bruteforce.argtypes = [f32p, C.c_uint32, f32p, C.c_uint32, u32p, C.c_uint64]
        L.orc_transform_positions.argtypes = [f32p, C.c_uint32, f32p]
        L.orc_gen_uv_sphere.argtypes = [C.c_uint32, f32p, u32p]

    # ---- layout ----
    def get_size(self, t):
        return self.lib.orc_get_size(t)

    def implicit_to_real(self, i, leaf_lev, vl):
        return self.lib.orc_implicit_to_real(i, leaf_lev, vl)

    def real_to_implicit(self, r, leaf_lev, vl):
        return self.lib.orc_real_to_implicit(r, leaf_lev, vl)

    def have_rchild(self, i, leaf_lev, vl):
        return bool(self.lib.orc_have_rchild(i, leaf_lev, vl))

    def most_right_valid(self, level, leaf_lev, vl):
        return self.lib.orc_most_right_valid(level, leaf_lev, vl)

    def level_real_count(self, level, leaf_lev, vl):
        return self.lib.orc_level_real_count(level, leaf_lev, vl)

    # ---- geometry ----
    def mesh_aabb(self, pos):
        pos, pp = _f32(pos)
        out = np.empty(6, np.float32)
        self.lib.orc_mesh_aabb(pp, pos.shape[0], out.ctypes.data_as(f32p))
        return out

    def leaf_aabbs(self, pos, faces):
        pos, pp = _f32(pos)
        faces, fp = _u32(faces)
        out = np.empty((faces.shape[0], 6), np.float32)
        self.lib.orc_leaf_aabbs(pp, fp, faces.shape[0], out.ctypes.data_as(f32p))
        return out

    def morton_keys(self, pos, faces, mesh_aabb):
        pos, pp = _f32(pos)
        faces, fp = _u32(faces)
        mesh_aabb, mp = _f32(mesh_aabb)
        keys = np.empty(faces.shape[0], np.uint32)
        self.lib.orc_morton_keys(pp, fp, faces.shape[0], mp, keys.ctypes.data_as(u32p))
        return keys

    def stable_sort_perm(self, keys):
        keys, kp = _u32(keys)
        perm = np.empty(keys.shape[0], np.uint32)
        self.lib.orc_stable_sort_perm(kp, keys.shape[0], perm.ctypes.data_as(u32p))
        return perm

    def tree_from_faces(self, pos, faces):
        pos, pp = _f32(pos)
        faces, fp = _u32(faces)
        T = faces.shape[0]
        nodes = np.empty((self.get_size(T), 6), np.float32)
        self.lib.orc_tree_from_faces(pp, fp, T, nodes.ctypes.data_as(f32p))
        return nodes

    refit = tree_from_faces

    def build(self, pos, faces, mesh_aabb=None):
        """-> dict(nodes[N,6], faces[T,3] sorted, perm[T] sorted->original, keys[T] sorted)"""
        pos, pp = _f32(pos)
        faces, fp = _u32(faces)
        if mesh_aabb is None:
            mesh_aabb = self.mesh_aabb(pos)
        mesh_aabb, mp = _f32(mesh_aabb)
        T = faces.shape[0]
        nodes = np.empty((self.get_size(T), 6), np.float32)
        sfaces = np.empty((T, 3), np.uint32)
        perm = np.empty(T, np.uint32)
        keys = np.empty(T, np.uint32)
        self.lib.orc_build(pp, fp, T, mp, nodes.ctypes.data_as(f32p), sfaces.ctypes.data_as(u32p),
                           perm.ctypes.data_as(u32p), keys.ctypes.data_as(u32p))
        return dict(nodes=nodes, faces=sfaces, perm=perm, keys=keys)

    def tri_tri(self, p, q):
        p, pp = _f32(p)
        q, qp = _f32(q)
        return bool(self.lib.orc_tri_tri(pp, qp))

    def aabb_overlap(self, a, b):
        a, ap = _f32(a)
        b, bp = _f32(b)
        return bool(self.lib.orc_aabb_overlap(ap, bp))

    def detect(self, trees, want_hist=False):
        """trees: list of (nodes[N,6], faces[T,3], pos[V,3]) -> (pairs[H,4] unsorted, n_candidates[, hist])"""
        n = len(trees)
        keep = []
        nodes_arr = (f32p * n)()
        faces_arr = (u32p * n)()
        pos_arr = (f32p * n)()
        Ts = np.empty(n, np.uint32)
        for k, (nodes, faces, pos) in enumerate(trees):
            a, nodes_arr[k] = _f32(nodes)
            b, faces_arr[k] = _u32(faces)
            c, pos_arr[k] = _f32(pos)
            keep += [a, b, c]
            Ts[k] = b.shape[0]
        ncand = C.c_uint64(0)
        hist = np.zeros(64, np.uint64)
        cap = 1 << 16
        while True:
            out = np.empty((cap, 4), np.uint32)
            cnt = self.lib.orc_detect(n, nodes_arr, faces_arr, pos_arr, Ts.ctypes.data_as(u32p),
                                      out.ctypes.data_as(u32p), cap, C.byref(ncand), hist.ctypes.data_as(u64p))
            if cnt <= cap:
                break
            cap = int(cnt)
        res = (out[:cnt].copy(), int(ncand.value))
        return res + (hist,) if want_hist else res

    def candidates_bruteforce(self, leaves_a, leaves_b):
        la, lap = _f32(leaves_a)
        lb, lbp = _f32(leaves_b)
        cap = 1 << 16
        while True:
            out = np.empty((cap, 2), np.uint32)
            cnt = self.lib.orc_candidates_bruteforce(lap, la.shape[0], lbp, lb.shape[0], out.ctypes.data_as(u32p), cap)
            if cnt <= cap:
                return out[:cnt].copy()
            cap = int(cnt)

    def transform_positions(self, pos, M):
        """M: 16 floats column-major (glm layout). Returns transformed copy."""
        pos = np.array(pos, dtype=np.float32, order="C", copy=True)
        M, mp = _f32(np.asarray(M).reshape(16))
        self.lib.orc_transform_positions(pos.ctypes.data_as(f32p), pos.shape[0], mp)
        return pos

    def gen_uv_sphere(self, n):
        pos = np.empty(((n + 1) * n, 3), np.float32)
        faces = np.empty((2 * n * n, 3), np.uint32)
        self.lib.orc_gen_uv_sphere(n, pos.ctypes.data_as(f32p), faces.ctypes.data_as(u32p))
        return pos, faces


# =====================================================================================================
# ref (unmodified reference classes)
# =====================================================================================================
def ref_available():
    return os.path.exists(REF_PATH)


def ref_deep_available():
    return os.path.exists(REF_DEEP_PATH)


class Ref:
    """Thin handle-based wrapper; see oracle/ref_shim.cpp."""

    def __init__(self, path=REF_PATH):
        L = self.lib = C.CDLL(path)
        vp = C.c_void_p
        L.ref_mesh_create.restype = vp
        L.ref_mesh_create.argtypes = [f32p, C.c_uint32, u32p, C.c_uint32]
        L.ref_mesh_clone.restype = vp
        L.ref_mesh_clone.argtypes = [vp]
        L.ref_mesh_destroy.argtypes = [vp]
        L.ref_mesh_translate.argtypes = [vp, C.c_float, C.c_float, C.c_float]
        L.ref_mesh_rotate.argtypes = [vp, C.c_float, C.c_float, C.c_float, C.c_float]
        L.ref_mesh_transform.argtypes = [vp, f32p]
        L.ref_mesh_set_positions.argtypes = [vp, f32p]
        L.ref_mesh_get_positions.argtypes = [vp, f32p]
        L.ref_mesh_get_aabb.argtypes = [vp, f32p]
        L.ref_mesh_get_center.argtypes = [vp, f32p]
        L.ref_glm_translate.argtypes = [C.c_float] * 3 + [f32p]
        L.ref_glm_rotate_about.argtypes = [f32p] + [C.c_float] * 4 + [f32p]
        L.ref_bvh_create.restype = vp
        L.ref_bvh_create.argtypes = [vp]
        L.ref_bvh_destroy.argtypes = [vp]
        L.ref_bvh_build.argtypes = [vp]
        L.ref_bvh_refit.argtypes = [vp]
        L.ref_bvh_node_count.restype = C.c_uint32
        L.ref_bvh_node_count.argtypes = [vp]
        L.ref_bvh_depth.restype = C.c_uint32
        L.ref_bvh_depth.argtypes = [vp]
        L.ref_bvh_dump_bfs.restype = C.c_uint32
        L.ref_bvh_dump_bfs.argtypes = [vp, f32p, C.POINTER(C.c_int32)]
        L.ref_collide_create.restype = vp
        L.ref_collide_destroy.argtypes = [vp]
        L.ref_collide_add.argtypes = [vp, vp]
        L.ref_collide_detect.restype = C.c_uint32
        L.ref_collide_detect.argtypes = [vp]
        L.ref_collide_get_pairs.argtypes = [vp, u32p]
        L.ref_collide_vertex_array.restype = C.c_uint32
        L.ref_collide_vertex_array.argtypes = [vp, f32p]
        L.ref_box_wireframe.argtypes = [f32p, C.c_uint32, f32p, u32p]
        L.ref_triangle_intersect.restype = C.c_int
        L.ref_triangle_intersect.argtypes = [f32p, f32p]
        L.ref_aabb_overlap.restype = C.c_int
        L.ref_aabb_overlap.argtypes = [f32p, f32p]
        L.ref_oibvh_get_size.restype = C.c_uint32
        L.ref_oibvh_get_size.argtypes = [C.c_uint32]
        for name in ("ref_oibvh_implicit_to_real", "ref_oibvh_real_to_implicit", "ref_oibvh_most_right_valid",
                     "ref_oibvh_level_real_count"):
            getattr(L, name).restype = C.c_uint32
            getattr(L, name).argtypes = [C.c_uint32] * 3
        L.ref_oibvh_have_rchild.restype = C.c_int
        L.ref_oibvh_have_rchild.argtypes = [C.c_uint32] * 3

    # ---- Mesh ----
    def mesh_create(self, pos, faces):
        pos, pp = _f32(pos)
        faces, fp = _u32(faces)
        return C.c_void_p(self.lib.ref_mesh_create(pp, pos.shape[0], fp, faces.shape[0]))

    def mesh_clone(self, m):
        return C.c_void_p(self.lib.ref_mesh_clone(m))

    def mesh_destroy(self, m):
        self.lib.ref_mesh_destroy(m)

    def mesh_translate(self, m, t):
        self.lib.ref_mesh_translate(m, float(t[0]), float(t[1]), float(t[2]))

    def mesh_rotate(self, m, axis, angle_deg):
        self.lib.ref_mesh_rotate(m, float(axis[0]), float(axis[1]), float(axis[2]), float(angle_deg))

    def mesh_transform(self, m, M):
        M, mp = _f32(np.asarray(M).reshape(16))
        self.lib.ref_mesh_transform(m, mp)

    def mesh_set_positions(self, m, pos):
        pos, pp = _f32(pos)
        self.lib.ref_mesh_set_positions(m, pp)

    def mesh_positions(self, m, V):
        out = np.empty((V, 3), np.float32)
        self.lib.ref_mesh_get_positions(m, out.ctypes.data_as(f32p))
        return out

    def mesh_aabb(self, m):
        out = np.empty(6, np.float32)
        self.lib.ref_mesh_get_aabb(m, out.ctypes.data_as(f32p))
        return out

    def mesh_center(self, m):
        out = np.empty(3, np.float32)
        self.lib.ref_mesh_get_center(m, out.ctypes.data_as(f32p))
        return out

    def glm_translate(self, t):
        out = np.empty(16, np.float32)
        self.lib.ref_glm_translate(float(t[0]), float(t[1]), float(t[2]), out.ctypes.data_as(f32p))
        return out

    def glm_rotate_about(self, center, axis, angle_deg):
        c, cp = _f32(center)
        out = np.empty(16, np.float32)
        self.lib.ref_glm_rotate_about(cp, float(axis[0]), float(axis[1]), float(axis[2]), float(angle_deg),
                                      out.ctypes.data_as(f32p))
        return out

    # ---- SimpleBVH ----
    def bvh_create(self, m):
        return C.c_void_p(self.lib.ref_bvh_create(m))

    def bvh_destroy(self, b):
        self.lib.ref_bvh_destroy(b)

    def bvh_build(self, b):
        self.lib.ref_bvh_build(b)

    def bvh_refit(self, b):
        self.lib.ref_bvh_refit(b)

    def bvh_dump_bfs(self, b):
        n = self.lib.ref_bvh_node_count(b)
        aabbs = np.empty((n, 6), np.float32)
        tri = np.empty(n, np.int32)
        got = self.lib.ref_bvh_dump_bfs(b, aabbs.ctypes.data_as(f32p), tri.ctypes.data_as(C.POINTER(C.c_int32)))
        assert got == n
        return aabbs, tri

    # ---- SimpleCollide ----
    def collide_create(self):
        return C.c_void_p(self.lib.ref_collide_create())

    def collide_destroy(self, c):
        self.lib.ref_collide_destroy(c)

    def collide_add(self, c, b):
        self.lib.ref_collide_add(c, b)

    def collide_detect(self, c):
        n = self.lib.ref_collide_detect(c)
        out = np.empty((n, 4), np.uint32)
        if n:
            self.lib.ref_collide_get_pairs(c, out.ctypes.data_as(u32p))
        return out

    def collide_vertex_array(self, c):
        """unmodified SimpleCollide::convertToVertexArray after a detect: [pairs*6, 3]"""
        n = self.lib.ref_collide_vertex_array(c, None)
        out = np.empty((n, 3), np.float32)
        if n:
            self.lib.ref_collide_vertex_array(c, out.ctypes.data_as(f32p))
        return out

    def box_wireframe(self, nodes):
        nodes, npx = _f32(np.asarray(nodes, np.float32).reshape(-1, 6))
        n = len(nodes)
        verts = np.empty((n * 8, 3), np.float32)
        idx = np.empty(n * 24, np.uint32)
        if n:
            self.lib.ref_box_wireframe(npx, n, verts.ctypes.data_as(f32p), idx.ctypes.data_as(u32p))
        return verts, idx

    # ---- free functions ----
    def tri_tri(self, p, q):
        p, pp = _f32(p)
        q, qp = _f32(q)
        return bool(self.lib.ref_triangle_intersect(pp, qp))

    def aabb_overlap(self, a, b):
        a, ap = _f32(a)
        b, bp = _f32(b)
        return bool(self.lib.ref_aabb_overlap(ap, bp))

    # convenience: whole reference CPU pipeline on a list of (pos, faces) meshes
    def detect_meshes(self, meshes):
        """meshes: list of (pos[V,3], faces[T,3]) -> pairs[H,4] (bvhA,bvhB,faceA,faceB) in the given face order"""
        ms, bs = [], []
        col = self.collide_create()
        for pos, faces in meshes:
            m = self.mesh_create(pos, faces)
            b = self.bvh_create(m)
            self.bvh_build(b)
            self.collide_add(col, b)
            ms.append(m)
            bs.append(b)
        pairs = self.collide_detect(col)
        self.collide_destroy(col)
        for b in bs:
            self.bvh_destroy(b)
        for m in ms:
            self.mesh_destroy(m)
        return pairs


def pair_vertices(pairs, trees):
    """Restates Scene::convertToVertexArray (src/cuda/scene.cu:68-93) == SimpleCollide::convertToVertexArray
    (src/cpu/simpleCollide.cpp:191-216): pair i -> rows 6i..6i+5 = positions of triangle A's three vertices, then
    triangle B's. pairs: [H,4] (bvhA, bvhB, triA, triB) with tri indexing trees[k] = (faces[T,3], pos[V,3])."""
    p = np.asarray(pairs, np.uint32).reshape(-1, 4)
    out = np.empty((len(p) * 6, 3), np.float32)
    for i, (a, b, ta, tb) in enumerate(p):
        fa, pa = trees[a]
        fb, pb = trees[b]
        out[6 * i:6 * i + 3] = pa[fa[ta]]
        out[6 * i + 3:6 * i + 6] = pb[fb[tb]]
    return out


_CUBE_SIGNS = np.array([[-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1],
                        [-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1]], np.float32)  # utils.cpp:21-31
_CUBE_EDGES = np.array([0, 1, 1, 2, 2, 3, 3, 0, 4, 5, 5, 6, 6, 7, 7, 4, 1, 5, 0, 4, 3, 7, 2, 6], np.uint32)  # utils.cpp:33-69


def box_wireframe(nodes, max_nodes=256, n_prims=None):
    """Restates OibvhTree::convertToVertexArray (src/cuda/oibvhTree.cu:69-124) + makeCube (src/utils/utils.cpp:15-70)
    in fp32: boxes of the first min(internal, max_nodes) nodes; h = 0.5f * (max - min); corner = (+-h) + (min - (-h))."""
    nodes = np.asarray(nodes, np.float32).reshape(-1, 6)
    internal = len(nodes) - n_prims if n_prims is not None else len(nodes)
    n = min(internal, max_nodes)
    mn, mx = nodes[:n, :3], nodes[:n, 3:]
    h = (np.float32(0.5) * (mx - mn)).astype(np.float32)
    diff = (mn - (-h)).astype(np.float32)                       # aabb.m_minimum - cubeVertices[4]
    verts = (_CUBE_SIGNS[None, :, :] * h[:, None, :]).astype(np.float32) + diff[:, None, :]
    idx = (_CUBE_EDGES[None, :] + (np.arange(n, dtype=np.uint32) * 8)[:, None]).astype(np.uint32)
    return verts.reshape(-1, 3).astype(np.float32), idx.reshape(-1)


def canonical_pairs(pairs, perms=None):
    """Canonical pair set: (objA, objB, origFaceA, origFaceB), rows sorted lexicographically.
    perms: per-object arrays sorted-position -> original face id (None = ids are already original)."""
    p = np.array(pairs, dtype=np.uint32).reshape(-1, 4).copy()
    if perms is not None and len(p):
        for k in range(len(perms)):
            if perms[k] is None:
                continue
            sel = p[:, 0] == k
            p[sel, 2] = perms[k][p[sel, 2]]
            sel = p[:, 1] == k
            p[sel, 3] = perms[k][p[sel, 3]]
    if len(p):
        order = np.lexsort((p[:, 3], p[:, 2], p[:, 1], p[:, 0]))
        p = p[order]
    return p
