"""Development helper: time the unmodified reference GPU path on the bench meshes (see bench.reference_gpu_baseline)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pos, faces = bench.make_meshes()
print(json.dumps(bench.reference_gpu_baseline(pos, faces, frames=12), indent=1))
