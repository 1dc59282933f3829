"""Host-only checks of the mesh input layer (include/oibvh/model.hpp: OBJ reader, Loop subdivision, generators,
Model deep copy) -- SURVEY.md §8 row f2. The checks live in tests/cpp/model_test.cpp; no GPU work."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "model_test")


def test_model_io_checks():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")])
    res = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    lines = res.stdout.splitlines()
    assert res.returncode == 0 and lines, res.stdout + res.stderr
    assert all(l.startswith("ok ") for l in lines), res.stdout
    assert len(lines) >= 19


def test_headless_driver_mirrors_reference_main():
    """the driver walks the reference's set-up and frame loop (main.cpp:127-151, 240-319) through the facade"""
    text = open(os.path.join(ROOT, "apps", "oibvh_headless.cpp")).read()
    for name in ["Model body1", "Model body2(body1)", "std::make_shared<OibvhTree>(tree1, body2.m_meshes[0])",
                 "->translate(", "->refit()", "scene.addOibvhTree", "detectCollision(DeviceType::GPU0", "rotateZ"]:
        assert name in text, name


def test_facade_compiles_in_glm_mode_against_reference_glm(tmp_path):
    """-DOIBVH_FACADE_USE_GLM: every C++ translation unit of this repo compiles with the
    reference's own glm as the vector/matrix types. Needs /root/reference/third/glm (absent on the GPU box)."""
    import pytest
    glm = "/root/reference/third"
    if not os.path.exists(os.path.join(glm, "glm", "glm.hpp")):
        pytest.skip("reference glm not available")
    for src in ("tests/cpp/glm_dropin.cpp", "apps/oibvh_headless.cpp", "tests/cpp/facade_demo.cpp",
                "tests/cpp/model_test.cpp"):
        res = subprocess.run(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-DOIBVH_FACADE_USE_GLM",
                              "-I", os.path.join(ROOT, "include"), "-I", glm, os.path.join(ROOT, src)],
                             capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, src + "\n" + res.stderr[:2000]
    # and the drop-in unit links against the library
    exe = tmp_path / "glm_dropin"
    res = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-DOIBVH_FACADE_USE_GLM", "-I", os.path.join(ROOT, "include"),
                          "-I", glm, os.path.join(ROOT, "tests/cpp/glm_dropin.cpp"), "-o", str(exe),
                          "-L", os.path.join(ROOT, "oibvh_b200"), "-loibvh_b200",
                          "-Wl,-rpath," + os.path.join(ROOT, "oibvh_b200")], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[:2000]
