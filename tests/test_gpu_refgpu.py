"""The UNMODIFIED reference GPU path (oracle/_ref/ref_gpu_bench: src/cuda/*.cu recompiled for sm_100a) against this
library on the same scene: the colliding pair sets must be identical. Skipped where the binary was not prebuilt."""
import os

import numpy as np
import pytest

import bench
import oibvh_b200 as ob
from oibvh_b200 import meshgen


def rows_by_vertices(pairs, faces_a, faces_b):
    """(bvhA, bvhB, triA, triB) -> rows (bvhA, bvhB, A's 3 vertex ids, B's 3 vertex ids), sorted"""
    r = np.concatenate([pairs[:, :2], faces_a[pairs[:, 2]], faces_b[pairs[:, 3]]], axis=1).astype(np.uint32)
    return r[np.lexsort(r.T[::-1])]


@pytest.mark.gpu
@pytest.mark.parametrize("nu,nv", [(64, 48), (256, 192)])
def test_reference_gpu_path_pair_set(ctx, nu, nv):
    if not os.path.exists(bench.REF_GPU_EXE):
        pytest.skip("oracle/_ref/ref_gpu_bench not prebuilt (needs /root/reference: make -C oracle refgpu)")
    pos, faces = meshgen.blob(nu, nv, seed=1234)
    faces = meshgen.shuffle_faces(faces)
    keep = {}
    res = bench.reference_gpu_baseline(pos, faces, frames=3, keep=keep)
    assert res is not None and "unavailable" not in res, res
    ref_rows = keep["pairs"]
    ref_rows = ref_rows[np.lexsort(ref_rows.T[::-1])]
    # same final positions of body B as the reference produced (its Mesh::transform runs on the GPU)
    ta = ob.OibvhTree(ob.Mesh(pos, faces), ctx=ctx)
    tb = ob.OibvhTree(ob.Mesh(keep["posB"], faces), ctx=ctx)
    ta.build()
    tb.build()
    sc = ob.Scene(ctx)
    sc.addOibvhTree(ta)
    sc.addOibvhTree(tb)
    sc.detectCollision(ob.DeviceType.GPU0, 4, 3)
    ours = rows_by_vertices(sc.m_intTriPairs, ta.download()["faces"], tb.download()["faces"])
    assert res["pairs"] == len(ref_rows) == len(ours) > 0
    assert np.array_equal(ours, ref_rows)
    for t in (ta, tb):
        t.close()
