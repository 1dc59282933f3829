"""oibvh_b200 -- B200-native oibvh collision path (build, refit, broad phase, narrow phase).

Python host mirror of the reference's operator surface on top of the C ABI (include/oibvh_b200.h):

    Mesh        include/utils/mesh.h:71-187       (positions, indices, m_aabb, translate/rotate/transform)
    OibvhTree   include/cuda/oibvhTree.cuh:44-93  (build, refit, getDepth, getPrimCount, copy-constructor)
    Scene       include/cuda/scene.cuh:26-67      (addOibvhTree, detectCollision, getIntTriPairCount)
    DeviceType  include/cuda/scene.cuh:12-24

The C++ facade with the same names lives in include/oibvh/oibvh.hpp. All compute runs in
liboibvh_b200.so (hand-written sm_100a CUDA); there is NO CPU fallback: if the library is missing the import
fails, and without a GPU every compute call raises OibvhError.
"""
import ctypes as C
import enum
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# OIBVH_B200_LIB selects another build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("OIBVH_B200_LIB") or os.path.join(_HERE, "liboibvh_b200.so")

STAGES = ("build", "refit", "broad", "narrow")
# kernels inside the stages (oibvh_ctx_stage_ms also reports them): keys + sort + emit = build
SUBSTAGES = ("keys", "sort", "emit", "transform")


class OibvhError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"oibvh_b200 error {code}: {msg}")
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C oibvh_b200/csrc`). oibvh_b200 has no CPU fallback.")
    return C.CDLL(LIB_PATH)


_lib = _load()

_vp = C.c_void_p
_u32 = C.c_uint32
_u32p = C.POINTER(C.c_uint32)
_f32p = C.POINTER(C.c_float)

# every symbol include/oibvh_b200.h declares, with its signature (tests/test_abi.py checks this list against the header)
_SIGNATURES = {
    "oibvh_last_error": (C.c_char_p, []),
    "oibvh_version": (C.c_int, []),
    "oibvh_device_count": (C.c_int, []),
    "oibvh_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "oibvh_ctx_create_on_stream": (C.c_int, [C.c_int, _vp, C.POINTER(_vp)]),
    "oibvh_ctx_destroy": (C.c_int, [_vp]),
    "oibvh_ctx_synchronize": (C.c_int, [_vp]),
    "oibvh_ctx_get_stream": (C.c_int, [_vp, C.POINTER(_vp)]),
    "oibvh_ctx_launch_count": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "oibvh_ctx_enable_timing": (C.c_int, [_vp, C.c_int]),
    "oibvh_ctx_stage_ms": (C.c_int, [_vp, _f32p]),
    "oibvh_ctx_capture_begin": (C.c_int, [_vp]),
    "oibvh_ctx_capture_end": (C.c_int, [_vp, C.POINTER(_vp)]),
    "oibvh_ctx_capture_abort": (C.c_int, [_vp]),
    "oibvh_graph_launch": (C.c_int, [_vp]),
    "oibvh_graph_destroy": (C.c_int, [_vp]),
    "oibvh_tree_create": (C.c_int, [_vp, _vp, _u32, _vp, _u32, _f32p, C.POINTER(_vp)]),
    "oibvh_tree_create_from_device": (C.c_int, [_vp, _vp, _u32, _vp, _u32, _f32p, C.POINTER(_vp)]),
    "oibvh_tree_clone": (C.c_int, [_vp, C.POINTER(_vp)]),
    "oibvh_tree_replicate": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "oibvh_tree_sync_replica": (C.c_int, [_vp, _vp]),
    "oibvh_tree_destroy": (C.c_int, [_vp]),
    "oibvh_tree_set_positions": (C.c_int, [_vp, _vp]),
    "oibvh_tree_set_positions_from_device": (C.c_int, [_vp, _vp]),
    "oibvh_tree_transform": (C.c_int, [_vp, _f32p]),
    "oibvh_tree_build": (C.c_int, [_vp]),
    "oibvh_tree_build_many": (C.c_int, [C.POINTER(_vp), _u32]),
    "oibvh_tree_refit_many": (C.c_int, [C.POINTER(_vp), _u32]),
    "oibvh_tree_transform_refit_many": (C.c_int, [C.POINTER(_vp), _u32, _f32p, C.c_char_p]),
    "oibvh_tree_transform_many": (C.c_int, [C.POINTER(_vp), _u32, _f32p]),
    "oibvh_tree_transform_many_from_device": (C.c_int, [C.POINTER(_vp), _u32, _vp]),
    "oibvh_tree_refit": (C.c_int, [_vp]),
    "oibvh_tree_get_info": (C.c_int, [_vp, _u32p, _u32p, _u32p, _u32p]),
    "oibvh_tree_is_built": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "oibvh_tree_download": (C.c_int, [_vp, _vp, _vp, _vp]),
    "oibvh_tree_download_positions": (C.c_int, [_vp, _vp]),
    "oibvh_tree_download_keys": (C.c_int, [_vp, _vp]),
    "oibvh_tree_device_views": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "oibvh_scene_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "oibvh_scene_destroy": (C.c_int, [_vp]),
    "oibvh_scene_add_tree": (C.c_int, [_vp, _vp]),
    "oibvh_scene_set_shard": (C.c_int, [_vp, _u32, _u32]),
    "oibvh_scene_reserve": (C.c_int, [_vp, _u32, _u32, _u32]),
    "oibvh_scene_set_self_collision": (C.c_int, [_vp, C.c_int]),
    "oibvh_scene_set_coherence": (C.c_int, [_vp, C.c_int, _u32]),
    "oibvh_mgpu_export": (C.c_int, [_vp, _vp]),
    "oibvh_mgpu_attach": (C.c_int, [_vp, _vp]),
    "oibvh_mgpu_detach": (C.c_int, [_vp]),
    "oibvh_mgpu_open_frame": (C.c_int, [_vp]),
    "oibvh_scene_detect": (C.c_int, [_vp, _u32, _u32, _u32p, _u32p]),
    "oibvh_scene_detect_async": (C.c_int, [_vp, _u32, _u32]),
    "oibvh_scene_get_counts": (C.c_int, [_vp, _u32p, _u32p]),
    "oibvh_scene_get_pairs": (C.c_int, [_vp, _vp]),
    "oibvh_scene_device_pairs": (C.c_int, [_vp, C.POINTER(_vp), _u32p]),
    "oibvh_scene_get_phase_cycles": (C.c_int, [_vp, _u32p, _u32, _u32p]),
    "oibvh_scene_device_counters": (C.c_int, [_vp, C.POINTER(_vp)]),
    "oibvh_scene_pair_capacity": (C.c_int, [_vp, _u32p]),
    "oibvh_scene_get_round_stats": (C.c_int, [_vp, _u32p, _u32, _u32p]),
    "oibvh_scene_pair_vertices_device": (C.c_int, [_vp, _vp, _u32]),
    "oibvh_scene_pair_vertices": (C.c_int, [_vp, _vp]),
    "oibvh_tree_box_wireframe": (C.c_int, [_vp, _u32, _vp, _vp, _u32p]),
}
for _name, (_res, _args) in _SIGNATURES.items():
    _fn = getattr(_lib, _name)  # AttributeError here = the library does not export what the header declares
    _fn.restype = _res
    _fn.argtypes = _args


def _check(rc):
    if rc != 0:
        raise OibvhError(rc, _lib.oibvh_last_error().decode(errors="replace"))


def device_count():
    return _lib.oibvh_device_count()


def version():
    return _lib.oibvh_version()


def _ptr(a):
    return a.ctypes.data_as(_vp)


# =====================================================================================================
# glm-compatible 4x4 helpers (column-major float32[16]); operation order follows third/glm/ext/matrix_transform.inl
# so that matrices equal the ones Mesh::translate / Mesh::rotate build (src/utils/mesh.cpp:171-185)
# =====================================================================================================
_f = np.float32


def mat_identity():
    return np.eye(4, dtype=_f).reshape(16).copy()


def mat_translate(m, v):
    """glm::translate(m, v): Result[3] = m[0]*v[0] + m[1]*v[1] + m[2]*v[2] + m[3]  (matrix_transform.inl:10-15)"""
    m = np.asarray(m, _f).reshape(4, 4).copy()  # rows of this array = glm columns
    v = np.asarray(v, _f)
    m[3] = ((m[0] * v[0] + m[1] * v[1]) + m[2] * v[2]) + m[3]
    return m.reshape(16)


def mat_rotate(m, angle_rad, axis):
    """glm::rotate(m, angle, axis)  (matrix_transform.inl:18-46), float32 arithmetic in glm's order"""
    m = np.asarray(m, _f).reshape(4, 4)
    a = _f(angle_rad)
    c, s = _f(np.cos(a)), _f(np.sin(a))
    v = np.asarray(axis, _f)
    d = _f(_f(v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])
    ax = v * _f(_f(1) / np.sqrt(d))
    t = _f(_f(1) - c) * ax
    R = np.zeros((3, 3), _f)
    R[0, 0] = c + t[0] * ax[0]
    R[0, 1] = t[0] * ax[1] + s * ax[2]
    R[0, 2] = t[0] * ax[2] - s * ax[1]
    R[1, 0] = t[1] * ax[0] - s * ax[2]
    R[1, 1] = c + t[1] * ax[1]
    R[1, 2] = t[1] * ax[2] + s * ax[0]
    R[2, 0] = t[2] * ax[0] + s * ax[1]
    R[2, 1] = t[2] * ax[1] - s * ax[0]
    R[2, 2] = c + t[2] * ax[2]
    out = np.empty((4, 4), _f)
    for k in range(3):
        out[k] = (m[0] * R[k, 0] + m[1] * R[k, 1]) + m[2] * R[k, 2]
    out[3] = m[3]
    return out.reshape(16)


def mat_apply_point(m, p):
    """glm mat4 * vec4(p, 1): (m0*x + m1*y) + (m2*z + m3*w)  (type_mat4x4.inl:561-572)"""
    m = np.asarray(m, _f).reshape(4, 4)
    p = np.asarray(p, _f)
    r = (m[0] * p[0] + m[1] * p[1]) + (m[2] * p[2] + m[3] * _f(1))
    return r[:3].astype(_f)


# =====================================================================================================
# Context
# =====================================================================================================
class Context:
    """one device + one stream (oibvh_ctx). `stream` = raw cudaStream_t int (e.g. torch stream.cuda_stream)."""

    def __init__(self, device=0, stream=None):
        h = _vp()
        if stream is None:
            _check(_lib.oibvh_ctx_create(int(device), C.byref(h)))
        else:
            _check(_lib.oibvh_ctx_create_on_stream(int(device), _vp(int(stream)), C.byref(h)))
        self._h = h
        self.device = int(device)

    def close(self):
        if self._h:
            _lib.oibvh_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        _check(_lib.oibvh_ctx_synchronize(self._h))

    @property
    def stream(self):
        s = _vp()
        _check(_lib.oibvh_ctx_get_stream(self._h, C.byref(s)))
        return s.value or 0

    def launch_count(self):
        n = C.c_uint64()
        _check(_lib.oibvh_ctx_launch_count(self._h, C.byref(n)))
        return n.value

    def enable_timing(self, on=True):
        _check(_lib.oibvh_ctx_enable_timing(self._h, 1 if on else 0))

    def stage_ms(self):
        ms = (C.c_float * 8)()
        _check(_lib.oibvh_ctx_stage_ms(self._h, ms))
        return dict(zip(STAGES + SUBSTAGES, [float(x) for x in ms]))

    def capture_begin(self):
        _check(_lib.oibvh_ctx_capture_begin(self._h))

    def capture_end(self):
        g = _vp()
        _check(_lib.oibvh_ctx_capture_end(self._h, C.byref(g)))
        return Graph(self, g)

    def capture_abort(self):
        _check(_lib.oibvh_ctx_capture_abort(self._h))


class Graph:
    def __init__(self, ctx, h):
        self._ctx = ctx
        self._h = h

    def launch(self):
        _check(_lib.oibvh_graph_launch(self._h))

    def close(self):
        if self._h:
            _lib.oibvh_graph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


# =====================================================================================================
# Mesh (host-side input type; GL / texture members of the reference are out of scope)
# =====================================================================================================
class Mesh:
    """positions [V,3] float32 + indices [T,3] uint32; m_aabb is fixed at construction like Mesh::setupAABB
    (src/utils/mesh.cpp:91-98) and m_center like Mesh::setupCenter (:145-153)."""

    def __init__(self, positions, indices):
        self.m_positions = np.array(positions, dtype=np.float32, order="C").reshape(-1, 3)
        self.m_indices = np.array(indices, dtype=np.uint32, order="C").reshape(-1, 3)
        self.m_verticesCount = self.m_positions.shape[0]
        self.m_facesCount = self.m_indices.shape[0]
        p = self.m_positions
        # exact: min/max are selections, so numpy's reduction equals the sequential glm::min/max chain up to the
        # sign of zero; canonicalise -0 -> the value the sequential chain keeps (first occurrence wins on ties for
        # neither min nor max in glm: glm::min(vertex, cur) returns `vertex` on ties, i.e. the LAST tied element)
        self.m_aabb = np.concatenate([_seq_min(p), _seq_max(p)]).astype(np.float32)
        c = np.zeros(3, np.float32)
        # m_center: sequential float32 accumulation then divide (mesh.cpp:147-152)
        c = np.add.accumulate(p, axis=0, dtype=np.float32)[-1] / np.float32(self.m_verticesCount)
        self.m_center = c.astype(np.float32)
        self._version = 0

    def copy(self):
        m = Mesh.__new__(Mesh)
        m.m_positions = self.m_positions.copy()
        m.m_indices = self.m_indices.copy()
        m.m_verticesCount = self.m_verticesCount
        m.m_facesCount = self.m_facesCount
        m.m_aabb = self.m_aabb.copy()
        m.m_center = self.m_center.copy()
        m._version = 0
        return m

    # -- transforms: matrices built exactly like mesh.cpp:155-185; the vertex update itself runs on the GPU inside
    #    OibvhTree (device-resident positions) and is mirrored on the host copy here for callers that read it.
    def transform_matrix_translate(self, t):
        return mat_translate(mat_identity(), t)

    def transform_matrix_rotate(self, axis, angle_deg):
        m = mat_translate(mat_identity(), self.m_center)
        m = mat_rotate(m, np.float32(angle_deg) * np.float32(0.01745329251994329576923690768489), axis)
        m = mat_translate(m, -self.m_center)
        return m

    def transform(self, M):
        """Mesh::transform (mesh.cpp:187-213) on the host copy: p = M * (p,1) with glm's operation order."""
        M = np.asarray(M, np.float32).reshape(4, 4)
        p = self.m_positions
        x, y, z = p[:, 0:1], p[:, 1:2], p[:, 2:3]
        r = (M[0][None, :] * x + M[1][None, :] * y) + (M[2][None, :] * z + M[3][None, :] * np.float32(1))
        self.m_center = mat_apply_point(M, self.m_center)
        self.m_positions = np.ascontiguousarray(r[:, :3], dtype=np.float32)
        self._version += 1

    def translate(self, t):
        self.transform(self.transform_matrix_translate(t))

    def rotate(self, axis, angle_deg):
        self.transform(self.transform_matrix_rotate(axis, angle_deg))

    def rotateX(self, angle=1.0):
        self.rotate((1.0, 0.0, 0.0), angle)

    def rotateY(self, angle=1.0):
        self.rotate((0.0, 1.0, 0.0), angle)

    def rotateZ(self, angle=1.0):
        self.rotate((0.0, 0.0, 1.0), angle)


def _seq_min(p):
    """column-wise result of folding glm::min(vertex, cur) = (cur < vertex) ? cur : vertex from +FLT_MAX"""
    out = np.empty(3, np.float32)
    for a in range(3):
        col = p[:, a]
        m = col.min() if len(col) else np.float32(np.finfo(np.float32).max)
        if m == 0:
            # ties between +0 and -0: the fold keeps the LAST zero seen
            zeros = col[col == 0]
            m = zeros[-1]
        out[a] = m
    return out


def _seq_max(p):
    """column-wise result of folding glm::max(vertex, cur) = (vertex < cur) ? cur : vertex from -FLT_MAX"""
    out = np.empty(3, np.float32)
    for a in range(3):
        col = p[:, a]
        m = col.max() if len(col) else -np.float32(np.finfo(np.float32).max)
        if m == 0:
            zeros = col[col == 0]
            m = zeros[-1]
        out[a] = m
    return out


# =====================================================================================================
# OibvhTree
# =====================================================================================================
class OibvhTree:
    """OibvhTree(mesh) or OibvhTree(other_tree, mesh) like the reference's two constructors
    (src/cuda/oibvhTree.cu:9-33)."""

    def __init__(self, a, b=None, ctx=None):
        if isinstance(a, OibvhTree):
            other, mesh = a, b
            assert mesh is not None, "OibvhTree(other, mesh)"
            self.ctx = other.ctx
            self.m_mesh = mesh
            h = _vp()
            _check(_lib.oibvh_tree_clone(other._h, C.byref(h)))
            self._h = h
            self.m_buildDone = other.m_buildDone
        else:
            mesh = a
            self.ctx = ctx or default_context()
            self.m_mesh = mesh
            h = _vp()
            aabb = np.ascontiguousarray(mesh.m_aabb, np.float32)
            _check(_lib.oibvh_tree_create(self.ctx._h, _ptr(mesh.m_positions), mesh.m_verticesCount,
                                          _ptr(mesh.m_indices), mesh.m_facesCount, aabb.ctypes.data_as(_f32p),
                                          C.byref(h)))
            self._h = h
            self.m_buildDone = False

    def replicate(self, ctx):
        """a replica of this tree on another context / device of this process (oibvh_tree_replicate)"""
        t = OibvhTree.__new__(OibvhTree)
        t.ctx, t.m_mesh, t.m_buildDone = ctx, self.m_mesh, self.m_buildDone
        h = _vp()
        _check(_lib.oibvh_tree_replicate(self._h, ctx._h, C.byref(h)))
        t._h = h
        return t

    def sync_replica(self, src):
        """refresh this replica after `src` changed (positions, nodes, face order)"""
        _check(_lib.oibvh_tree_sync_replica(self._h, src._h))
        self.m_buildDone = src.m_buildDone

    def close(self):
        if getattr(self, "_h", None):
            _lib.oibvh_tree_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference methods --
    def build(self):
        _check(_lib.oibvh_tree_build(self._h))
        self.m_buildDone = True

    def refit(self, upload=True):
        """re-reads the mesh positions (oibvhTree.cu:196-199) unless upload=False (positions already on device)"""
        if upload:
            _check(_lib.oibvh_tree_set_positions(self._h, _ptr(self.m_mesh.m_positions)))
        _check(_lib.oibvh_tree_refit(self._h))

    def getPrimCount(self):
        return self.info()[0]

    def getDepth(self):
        return self.info()[3]

    # -- extensions --
    def info(self):
        T, V, N, D = _u32(), _u32(), _u32(), _u32()
        _check(_lib.oibvh_tree_get_info(self._h, C.byref(T), C.byref(V), C.byref(N), C.byref(D)))
        return T.value, V.value, N.value, D.value

    def set_positions(self, positions):
        p = np.ascontiguousarray(positions, np.float32)
        assert p.size == 3 * self.info()[1]
        _check(_lib.oibvh_tree_set_positions(self._h, _ptr(p)))
        self._keep = p  # async H2D from pageable memory is staged by the runtime, but keep a ref anyway

    def set_positions_from_host_ptr(self, host_ptr):
        """raw host pointer (e.g. a pinned torch tensor's data_ptr()); the caller keeps it alive until the next sync"""
        _check(_lib.oibvh_tree_set_positions(self._h, _vp(int(host_ptr))))

    def set_positions_from_device(self, dev_ptr):
        _check(_lib.oibvh_tree_set_positions_from_device(self._h, _vp(int(dev_ptr))))

    def transform(self, M):
        m = np.ascontiguousarray(M, np.float32).reshape(16)
        _check(_lib.oibvh_tree_transform(self._h, m.ctypes.data_as(_f32p)))

    def convertToVertexArray(self, max_nodes=256):
        """OibvhTree::convertToVertexArray (oibvhTree.cu:69-124): wireframe boxes of the first min(internal, 256)
        nodes -> (vertices [n*8, 3] float32, indices [n*24] uint32)"""
        T, V, N, _ = self.info()
        n = min(N - T, int(max_nodes))
        verts = np.empty((n * 8, 3), np.float32)
        idx = np.empty(n * 24, np.uint32)
        got = _u32()
        _check(_lib.oibvh_tree_box_wireframe(self._h, int(max_nodes), _ptr(verts) if n else None,
                                             _ptr(idx) if n else None, C.byref(got)))
        assert got.value == n
        return verts, idx

    def download(self, nodes=True, faces=True, perm=True):
        T, V, N, _ = self.info()
        out = {}
        an = np.empty((N, 6), np.float32) if nodes else None
        af = np.empty((T, 3), np.uint32) if faces else None
        ap = np.empty(T, np.uint32) if perm else None
        _check(_lib.oibvh_tree_download(self._h, _ptr(an) if nodes else None, _ptr(af) if faces else None,
                                        _ptr(ap) if perm else None))
        if nodes:
            out["nodes"] = an
        if faces:
            out["faces"] = af
        if perm:
            out["perm"] = ap
        return out

    @property
    def m_aabbTree(self):
        return self.download(True, False, False)["nodes"]

    @property
    def m_faces(self):
        return self.download(False, True, False)["faces"]

    @property
    def m_positions(self):
        V = self.info()[1]
        p = np.empty((V, 3), np.float32)
        _check(_lib.oibvh_tree_download_positions(self._h, _ptr(p)))
        return p

    def sorted_keys(self):
        k = np.empty(self.info()[0], np.uint32)
        _check(_lib.oibvh_tree_download_keys(self._h, _ptr(k)))
        return k

    def device_views(self):
        a, b, c = _vp(), _vp(), _vp()
        _check(_lib.oibvh_tree_device_views(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value


class TreeBatch:
    """a fixed list of trees with its handle array prepared once: pass it to build_many / refit_many /
    transform_many every frame (a many-body scene has thousands of trees; rebuilding the array costs more than the
    launch)"""

    def __init__(self, trees):
        self.trees = list(trees)
        self._arr = (_vp * len(self.trees))(*[t._h for t in self.trees])

    def __len__(self):
        return len(self.trees)

    def __iter__(self):
        return iter(self.trees)


def _handles(trees):
    if isinstance(trees, TreeBatch):
        return trees._arr
    return (_vp * len(trees))(*[t._h for t in trees])


def build_many(trees):
    """build several trees together (oibvh_tree_build_many): every tree of <= 4096 triangles in one launch, the keys
    of 2..4 larger trees in one cooperative sort launch"""
    _check(_lib.oibvh_tree_build_many(_handles(trees), len(trees)))
    for t in trees:
        t.m_buildDone = True


def refit_many(trees):
    """refit several trees on their device-resident positions (oibvh_tree_refit_many): small trees in one launch"""
    _check(_lib.oibvh_tree_refit_many(_handles(trees), len(trees)))


def transform_refit_many(trees, mats, apply=None):
    """Mesh::transform of the selected trees followed by the refit of ALL of them, in one call: the transform of one
    body overlaps the refit of another (oibvh_tree_transform_refit_many). mats: [n, 4, 4] or [n, 16] column-major
    (row i is ignored when apply[i] is false)."""
    h, n = _handles(trees), len(trees)
    m = np.ascontiguousarray(np.asarray(mats, np.float32).reshape(n, 16))
    flags = None if apply is None else bytes(bytearray(1 if a else 0 for a in apply))
    _check(_lib.oibvh_tree_transform_refit_many(h, n, m.ctypes.data_as(_f32p), flags))


def transform_many(trees, mats, device_ptr=None):
    """one rigid transform per tree in ONE launch (oibvh_tree_transform_many). mats: (n, 4, 4) glm-order matrices
    (mats[i][column][row], as Mesh.transform_matrix_*) on the host, or device_ptr = address of n x 16 floats."""
    if device_ptr is not None:
        _check(_lib.oibvh_tree_transform_many_from_device(_handles(trees), len(trees), _vp(device_ptr)))
        return
    m = np.ascontiguousarray(mats, np.float32).reshape(len(trees), 16)
    _check(_lib.oibvh_tree_transform_many(_handles(trees), len(trees), m.ctypes.data_as(_f32p)))


# =====================================================================================================
# Scene
# =====================================================================================================
class DeviceType(enum.IntEnum):
    CPU = -1
    GPU0 = 0
    GPU1 = 1
    GPU2 = 2
    GPU3 = 3
    GPU4 = 4
    GPU5 = 5
    GPU6 = 6
    GPU7 = 7
    GPU8 = 8


class Scene:
    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()
        h = _vp()
        _check(_lib.oibvh_scene_create(self.ctx._h, C.byref(h)))
        self._h = h
        self.m_oibvhTrees = []
        self.m_intTriPairCount = 0
        self._candidates = 0

    def close(self):
        if getattr(self, "_h", None):
            _lib.oibvh_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference methods --
    def addOibvhTree(self, tree):
        _check(_lib.oibvh_scene_add_tree(self._h, tree._h))
        self.m_oibvhTrees.append(tree)

    def detectCollision(self, deviceType=DeviceType.GPU0, entryLevel=0, expandLevels=1):
        """Scene::detectCollision (src/cuda/scene.cu:157-185). DeviceType.CPU is an empty TODO in the reference
        (scene.cu:187-190) and is rejected here: this package has no CPU path."""
        if int(deviceType) < 0:
            raise OibvhError(-1, "DeviceType.CPU: the reference's CPU detect is an empty TODO; no CPU path exists here")
        if int(deviceType) != self.ctx.device:
            raise OibvhError(-1, f"scene lives on GPU{self.ctx.device}, asked for GPU{int(deviceType)}")
        n, c = _u32(), _u32()
        _check(_lib.oibvh_scene_detect(self._h, int(entryLevel), int(expandLevels), C.byref(n), C.byref(c)))
        self.m_intTriPairCount, self._candidates = n.value, c.value

    def getIntTriPairCount(self):
        return self.m_intTriPairCount

    @property
    def m_intTriPairs(self):
        """[H,4] uint32 rows {bvhA, bvhB, triA, triB}; tri = index into that tree's Morton-sorted faces"""
        n = _u32()
        _check(_lib.oibvh_scene_get_counts(self._h, C.byref(n), None))
        out = np.empty((n.value, 4), np.uint32)
        _check(_lib.oibvh_scene_get_pairs(self._h, _ptr(out) if n.value else None))
        return out

    def convertToVertexArray(self):
        """Scene::convertToVertexArray (scene.cu:68-93): [n_pairs * 6, 3] float32 -- per pair the three vertices of
        the A triangle, then of the B triangle, gathered on the device"""
        n, _ = self.counts()
        out = np.empty((n * 6, 3), np.float32)
        _check(_lib.oibvh_scene_pair_vertices(self._h, _ptr(out) if n else None))
        self.m_vertices = out
        return out

    def pair_vertices_device(self, dev_ptr, capacity_pairs):
        """enqueue the same gather into a caller-provided device buffer (no sync, graph-capturable)"""
        _check(_lib.oibvh_scene_pair_vertices_device(self._h, _vp(int(dev_ptr)), int(capacity_pairs)))

    def get_pairs_into(self, host_ptr):
        """oibvh_scene_get_pairs into a caller-owned host buffer (e.g. pinned); returns the pair count"""
        n, _ = self.counts()
        if n:
            _check(_lib.oibvh_scene_get_pairs(self._h, _vp(int(host_ptr))))
        return n

    # -- extensions --
    def set_shard(self, rank, world):
        _check(_lib.oibvh_scene_set_shard(self._h, int(rank), int(world)))

    def detect_async(self, entryLevel=0, expandLevels=0):
        _check(_lib.oibvh_scene_detect_async(self._h, int(entryLevel), int(expandLevels)))

    def set_self_collision(self, enable=True):
        """also test every object against itself (non-adjacent triangle pairs of one mesh); opt-in extension"""
        _check(_lib.oibvh_scene_set_self_collision(self._h, 1 if enable else 0))

    def set_coherence(self, enable=True, cut_depth=0):
        """temporal coherence: start detections from a recorded BVTT cut while the trees are only refitted; opt-in"""
        _check(_lib.oibvh_scene_set_coherence(self._h, 1 if enable else 0, int(cut_depth)))

    def reserve(self, front_records=0, candidate_records=0, pair_records=0):
        """size the work queues up front (a multi-GPU scene cannot regrow them)"""
        _check(_lib.oibvh_scene_reserve(self._h, int(front_records), int(candidate_records), int(pair_records)))

    # -- multi-GPU: every rank appends its hits to rank 0's pair list through a peer mapping (oibvh_mgpu_*) --
    MGPU_HANDLE_BYTES = 160

    def mgpu_export(self):
        """rank 0: returns the 160-byte handle the other ranks attach to (move it with any transport)"""
        buf = C.create_string_buffer(self.MGPU_HANDLE_BYTES)
        _check(_lib.oibvh_mgpu_export(self._h, C.cast(buf, _vp)))
        return bytes(buf.raw)

    def mgpu_attach(self, handle):
        assert len(handle) == self.MGPU_HANDLE_BYTES
        buf = C.create_string_buffer(bytes(handle), self.MGPU_HANDLE_BYTES)
        _check(_lib.oibvh_mgpu_attach(self._h, C.cast(buf, _vp)))

    def mgpu_detach(self):
        _check(_lib.oibvh_mgpu_detach(self._h))

    def mgpu_open_frame(self):
        _check(_lib.oibvh_mgpu_open_frame(self._h))

    def counts(self):
        n, c = _u32(), _u32()
        _check(_lib.oibvh_scene_get_counts(self._h, C.byref(n), C.byref(c)))
        self.m_intTriPairCount, self._candidates = n.value, c.value
        return n.value, c.value

    def getCandidateCount(self):
        return self._candidates

    def device_pairs(self):
        p, n = _vp(), _u32()
        _check(_lib.oibvh_scene_device_pairs(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def phase_cycles(self):
        buf = (C.c_uint32 * 64)()
        n = _u32()
        _check(_lib.oibvh_scene_get_phase_cycles(self._h, buf, 64, C.byref(n)))
        return [int(buf[i]) for i in range(n.value)]

    def device_counters(self):
        p = _vp()
        _check(_lib.oibvh_scene_device_counters(self._h, C.byref(p)))
        return p.value

    def pair_capacity(self):
        n = _u32()
        _check(_lib.oibvh_scene_pair_capacity(self._h, C.byref(n)))
        return n.value

    def round_stats(self):
        buf = (C.c_uint32 * 64)()
        n = _u32()
        _check(_lib.oibvh_scene_get_round_stats(self._h, buf, 64, C.byref(n)))
        return [int(buf[i]) for i in range(n.value)]

    def canonical_pairs(self):
        """(bvhA, bvhB, origFaceA, origFaceB) rows sorted lexicographically"""
        p = self.m_intTriPairs.copy()
        perms = [t.download(False, False, True)["perm"] for t in self.m_oibvhTrees]
        for k, perm in enumerate(perms):
            sel = p[:, 0] == k
            p[sel, 2] = perm[p[sel, 2]]
            sel = p[:, 1] == k
            p[sel, 3] = perm[p[sel, 3]]
        if len(p):
            p = p[np.lexsort((p[:, 3], p[:, 2], p[:, 1], p[:, 0]))]
        return p


# shard bookkeeping shared by bench.py and the gloo tests (host logic of the multi-GPU path)
def shard_prefix(counts):
    """exclusive prefix of per-rank pair counts -> (offsets, total)"""
    counts = [int(c) for c in counts]
    off, s = [], 0
    for c in counts:
        off.append(s)
        s += c
    return off, s
