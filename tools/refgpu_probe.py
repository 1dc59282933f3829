import os, sys, subprocess, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench
for nu, nv in [(384, 256), (512, 256)]:
    pos, faces = bench.make_meshes(nu, nv)
    d = tempfile.mkdtemp()
    mesh, out = os.path.join(d, "mesh.bin"), os.path.join(d, "pairs.bin")
    with open(mesh, "wb") as f:
        f.write(np.array([len(pos), len(faces)], np.uint32).tobytes()); f.write(pos.astype(np.float32).tobytes())
        f.write(faces.astype(np.uint32).tobytes()); f.write(np.array(list(bench.OFFSET_B) + [0, 0, 1, 1], np.float32).tobytes())
    env = dict(os.environ, REF_GPU_VERBOSE="1", REF_GPU_COUT="1")
    try:
        r = subprocess.run([bench.REF_GPU_EXE, mesh, "3", out], capture_output=True, text=True, timeout=12, env=env)
        print(len(faces), "rc", r.returncode, r.stdout[-1500:], r.stderr)
    except subprocess.TimeoutExpired as e:
        print(len(faces), "TIMEOUT", ((e.stdout or b"").decode() if isinstance(e.stdout, bytes) else str(e.stdout))[-2500:], (e.stderr or b"").decode() if isinstance(e.stderr, bytes) else e.stderr)
