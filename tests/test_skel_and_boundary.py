"""CPU tests: the .skel model compiler, the C-ABI boundary (library loads and exports every
symbol include/dartb.h declares), and the host-side lowering (needs no GPU)."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

from dart_env_b200 import capi
from dart_env_b200.cstructs import CModel, CTask, pack_model, pack_task
from dart_env_b200.skel import (JOINT_PRISMATIC, JOINT_REVOLUTE, JOINT_WELD, SHAPE_CAPSULE, Model, parse_skel,
                                shape_inertia, euler_xyz_to_matrix, SkelError)
from dart_env_b200.tasks import SPECS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_hopper_model_matches_survey_appendix_a(models):
    m = models["DartHopper-v1"]
    assert [b.name for b in m.bodies] == ["h_pelvis_aux2", "h_pelvis_aux", "h_pelvis", "h_thigh", "h_shin", "h_foot"]
    assert [b.joint_type for b in m.bodies] == [JOINT_PRISMATIC, JOINT_PRISMATIC] + [JOINT_REVOLUTE] * 4
    assert np.allclose(m.bodies[2].axis, [0, 0, -1]) and np.allclose(m.bodies[3].axis, [0, 0, 1])
    assert [b.limit_enforced for b in m.bodies] == [False, False, False, True, True, True]
    assert m.bodies[3].q_lo == pytest.approx(-2.61799) and m.bodies[5].q_hi == pytest.approx(0.785398)
    assert [b.damping for b in m.bodies] == [0, 0, 0, 1.0, 1.0, 1.0]
    assert m.bodies[5].mass == pytest.approx(5.0893801) and np.allclose(m.bodies[5].com, [0.065, 0, 0])
    # parent->joint of thigh: pelvis at y=1.25, thigh at y=1.05
    assert np.allclose(m.bodies[3].T_parent_joint[:3, 3], [0, -0.2, 0])
    assert len(m.shapes) == 4 and all(s.type == SHAPE_CAPSULE for s in m.shapes)
    assert len(m.ground) == 1 and np.allclose(m.ground[0].size, [500, 0.05, 5])
    assert m.dt == 0.002 and np.allclose(m.gravity, [0, -9.81, 0])


def test_capsule_inertia_rule():
    """B.1: moment from the first shape in the shape's OWN frame (capsule axis = z)."""
    r, h, mass = 0.05, 0.4, 3.53429174
    I = shape_inertia(SHAPE_CAPSULE, np.array([r, h, 0]), mass)
    vc, vs = math.pi * r * r * h, 4.0 / 3.0 * math.pi * r ** 3
    mc, ms = mass * vc / (vc + vs), mass * vs / (vc + vs)
    assert I[2, 2] == pytest.approx(mc * r * r / 2 + ms * 0.4 * r * r)
    assert I[0, 0] == pytest.approx(mc * (h * h / 12 + r * r / 4) + ms * (0.4 * r * r + h * h / 4 + 3 * h * r / 8))
    assert I[0, 0] == I[1, 1] and I[0, 1] == 0


def test_models_shapes(models):
    assert models["DartWalker2d-v1"].n_dofs == 9 and len(models["DartWalker2d-v1"].shapes) == 7
    ch = models["DartHalfCheetah-v1"]
    assert ch.n_bodies == 10 and ch.n_dofs == 9 and ch.dt == 0.01
    assert ch.bodies[3].joint_type == JOINT_WELD and ch.bodies[3].dof == -1
    assert ch.bodies[4].spring_k == 240.0 and ch.bodies[4].damping == 6.0
    sn = models["DartSnake7Link-v1"]
    assert all(b.friction_coeff == 0.0 for b in sn.bodies)
    assert np.allclose(sn.bodies[2].axis, [0, 1, 0])


def test_euler_xyz_order():
    R = euler_xyz_to_matrix(0.3, -0.5, 1.1)
    Rx = euler_xyz_to_matrix(0.3, 0, 0); Ry = euler_xyz_to_matrix(0, -0.5, 0); Rz = euler_xyz_to_matrix(0, 0, 1.1)
    assert np.allclose(R, Rx @ Ry @ Rz)


def test_parse_own_skel_file_order_and_errors(tmp_path):
    """joint file order with a forward reference: the parent's joint is created first."""
    txt = """<?xml version="1.0" ?><skel version="1.0"><world name="w"><physics><time_step>0.001</time_step>
    <gravity>0 -9.81 0</gravity></physics>
    <skeleton name="arm">
      <body name="b2"><transformation>0 -1 0 0 0 0</transformation><inertia><mass>1</mass><offset>0 -0.5 0</offset></inertia>
        <visualization_shape><geometry><box><size>0.1 1 0.1</size></box></geometry></visualization_shape></body>
      <body name="b1"><transformation>0 0 0 0 0 0</transformation><inertia><mass>2</mass><offset>0 -0.5 0</offset>
         <moment_of_inertia><ixx>1</ixx><iyy>2</iyy><izz>3</izz><ixy>0</ixy><ixz>0</ixz><iyz>0</iyz></moment_of_inertia></inertia></body>
      <joint type="revolute" name="j2"><parent>b1</parent><child>b2</child><axis><xyz>0 0 2</xyz>
          <limit><lower>-1</lower><upper>1</upper></limit></axis></joint>
      <joint type="revolute" name="j1"><parent>world</parent><child>b1</child><axis><xyz>0 0 1</xyz></axis><init_pos>0.25</init_pos></joint>
    </skeleton></world></skel>"""
    p = tmp_path / "arm.skel"
    p.write_text(txt)
    m = parse_skel(str(p))
    assert [b.name for b in m.bodies] == ["b1", "b2"] and m.bodies[1].parent == 0
    assert np.allclose(m.bodies[1].axis, [0, 0, 1]) and m.bodies[1].has_limit and not m.bodies[1].limit_enforced
    assert np.allclose(np.diag(m.bodies[0].inertia), [1, 2, 3]) and m.bodies[0].q_init == 0.25
    assert m.bodies[1].inertia[1, 1] == pytest.approx(1 / 12 * (0.01 + 0.01))  # box, shape frame
    assert m.dt == 0.001 and parse_skel(str(p), 0.01).dt == 0.01
    with pytest.raises(IOError):
        parse_skel(str(tmp_path / "nope.skel"))
    (tmp_path / "ball.skel").write_text(txt.replace('type="revolute" name="j2"', 'type="ball" name="j2"'))
    with pytest.raises(SkelError):
        parse_skel(str(tmp_path / "ball.skel"))


def test_model_json_roundtrip(models, tmp_path):
    m = models["DartHalfCheetah-v1"]
    p = tmp_path / "m.json"
    m.save_json(str(p))
    m2 = Model.load_json(str(p))
    a, b = pack_model(m), pack_model(m2)
    assert bytes(a) == bytes(b)


def test_bundled_models_match_skel_files(models):
    """dart_env_b200/assets/*.model.json (what the GPU box loads) == the parsed reference .skel."""
    from dart_env_b200.skel import parse_skel
    ref_assets = "/root/reference/gym/envs/dart/assets"   # tests may read the reference where it exists (not on the GPU box)
    for env_id, spec in SPECS.items():
        if not os.path.exists(os.path.join(ref_assets, spec.skel)):
            pytest.skip("reference assets not on this machine")
        ms = parse_skel(os.path.join(ref_assets, spec.skel), spec.dt)
        ms.enforce_limits()
        if spec.friction_all is not None:
            for b in ms.bodies:
                b.friction_coeff = spec.friction_all
        # `models` (what the product loads by default) comes from the bundled compiled models
        assert bytes(pack_model(ms)) == bytes(pack_model(models[env_id]))


# ------------------------------------------------------------------ C-ABI
def test_library_exports_every_declared_symbol():
    L = capi.load()
    hdr = open(os.path.join(ROOT, "include", "dartb.h")).read()
    names = set(re.findall(r"\b(dartb_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    for n in sorted(names):
        assert hasattr(L, n), "libdartb.so does not export %s" % n
    assert set(capi._EXPORTS) == names
    assert b"sm_100a" in L.dartb_version()


def test_struct_sizes_match_header():
    """sizeof() of the ctypes mirrors must equal the C structs (compiled probe)."""
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include "dartb.h"\nint main(){printf("%zu %zu %zu %zu", sizeof(dartb_body_t), sizeof(dartb_shape_t), sizeof(dartb_model_t), sizeof(dartb_task_t));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "p.c")
        open(c, "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", os.path.join(d, "p")])
        out = subprocess.check_output([os.path.join(d, "p")]).decode().split()
    from dart_env_b200.cstructs import CBody, CShape
    assert [int(x) for x in out] == [ctypes.sizeof(CBody), ctypes.sizeof(CShape), ctypes.sizeof(CModel), ctypes.sizeof(CTask)]


@pytest.mark.parametrize("env_id,expect", [("DartHopper-v1", "planar-xy/static:hopper6/f32 nd=6 max_contacts=4"),
                                           ("DartWalker2d-v1", "planar-xy/static:walker9/f32 nd=9 max_contacts=7"),
                                           ("DartHalfCheetah-v1", "planar-xy/static:cheetah9/f32 nd=9 max_contacts=8"),
                                           ("DartSnake7Link-v1", "planar-zx/static:snake9/f32 nd=9 max_contacts=1")])
def test_lowering_picks_the_static_kernel(models, env_id, expect):
    assert capi.describe(models[env_id], SPECS[env_id].task) == expect


def test_lowering_rejects_non_planar_and_unknown(models):
    import copy
    m = copy.deepcopy(models["DartHopper-v1"])
    m.bodies[4].axis = np.array([0.0, 1.0, 0.0])  # shin joint leaves the plane
    with pytest.raises(capi.DartbError, match="non-planar"):
        capi.describe(m, SPECS["DartHopper-v1"].task)
    # a planar skeleton without a dedicated instantiation runs on the topology-generic loop kernel
    m = copy.deepcopy(models["DartHopper-v1"])
    m.shapes = m.shapes[:3]
    assert capi.describe(m, SPECS["DartHopper-v1"].task) == "planar-xy/loop:generic/f32 nd=6 max_contacts=3"
    # weld to world / unsupported statics are still rejected loudly
    m = copy.deepcopy(models["DartHopper-v1"])
    m.ground = m.ground + m.ground
    with pytest.raises(capi.DartbError, match="more than one static"):
        capi.describe(m, SPECS["DartHopper-v1"].task)


def test_no_gpu_fails_loudly(models):
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    from dart_env_b200.envs import make
    with pytest.raises(capi.DartbError, match="no CPU fallback"):
        make("DartHopper-v1")
    # and straight through the C-ABI
    L = capi.load()
    h = ctypes.c_void_p()
    cm, ct = pack_model(models["DartHopper-v1"]), pack_task(SPECS["DartHopper-v1"].task)
    assert L.dartb_create(ctypes.byref(cm), ctypes.byref(ct), 4, 0, 0, 0, ctypes.byref(h)) != 0
    assert b"no CPU fallback" in L.dartb_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dart_env_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "dart_oracle" not in txt.replace("oracle/dart_oracle.c", ""), f


def test_spaces_and_specs():
    from dart_env_b200.spaces import Box, batch_space
    b = Box(np.array([-1.0] * 3), np.array([1.0] * 3))
    assert b.dtype == np.float32 and b.shape == (3,)
    b.seed(0)
    s = b.sample()
    assert s in b and s.dtype == np.float32
    assert batch_space(b, 5).shape == (5, 3)
    hi = np.inf * np.ones(11)
    o = Box(-hi, hi)
    assert np.zeros(11) in o and o.sample().shape == (11,)


def test_output_pool_hands_out_only_unreferenced_slots():
    """VectorEnv(copy=True) semantics of the batched host path without copies (dart_env._PinnedOutPool): a slot is
    reused only when the caller holds no array or view of it.  (Host logic: plain memory stands in for pinned memory.)"""
    import ctypes as C

    from dart_env_b200.dart_env import _PinnedOutPool
    blocks = []

    def alloc(nbytes):
        buf = (C.c_char * nbytes)()
        blocks.append(buf)
        return buf, C.addressof(buf)

    pool = _PinnedOutPool(None, 16, 5, True, max_slots=4, alloc=alloc)
    s = pool.take()
    o, r, d, t = s["obs"], s["rew"], s["done"], s["trunc"]
    assert o.shape == (16, 5) and o.dtype == np.float32 and r.dtype == np.float64 and d.dtype == np.bool_ and t.dtype == np.bool_
    o[:] = 1.0
    del s
    s2 = pool.take()                      # `o` is still held: a different slot
    assert s2["obs"] is not o and len(pool.slots) == 2
    s2["obs"][:] = 2.0
    assert (o == 1.0).all()
    del s2
    view = o[3]                           # a VIEW keeps the slot out of circulation
    del o, r, d, t
    for _ in range(10):
        s3 = pool.take()
        assert s3["obs"].base is not view.base
        s3["obs"][:] = 3.0
        del s3
    assert (view == 1.0).all() and len(pool.slots) == 2
    del view
    seen = set()
    for _ in range(4):
        s4 = pool.take(); seen.add(id(s4)); del s4
    assert len(seen) == 2 and len(pool.slots) == 2    # both slots circulate again
    held = [pool.take()["obs"] for _ in range(4)]
    assert len({id(h) for h in held}) == 4 and pool.take() is None    # max_slots all held: caller falls back to copies


def test_planar_body_limit_matches_the_kernel_header():
    """cstructs.PM_MAXB is the index base of the per-reset dynamics draws (kernels.cuh::redraw_dynamics)"""
    from dart_env_b200.cstructs import PM_MAXB
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(here, "dart_env_b200", "csrc", "planar_model.h")).read()
    assert int(re.search(r"#define PM_MAXB (\d+)", txt).group(1)) == PM_MAXB


def test_header_is_plain_c(tmp_path):
    """include/dartb.h is the drop-in boundary: it must compile as C on its own (a cgo / ctypes / JNI binding includes
    nothing else)"""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(here, "include", "dartb.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_integration_doc_lists_every_export():
    """INTEGRATION.md maps each C-ABI entry point to the reference call it replaces: none may be missing"""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(here, "include", "dartb.h")).read()
    doc = open(os.path.join(here, "INTEGRATION.md")).read()
    names = set(re.findall(r"^(?:int|int32_t|int64_t|const char\*)\s+(dartb_[a-z0-9_]+)\(", hdr, flags=re.M))
    assert len(names) > 30
    missing = sorted(n for n in names if n not in doc)
    assert not missing, missing


def test_library_staleness_is_decided_by_content(tmp_path, monkeypatch):
    """build.is_stale(): a copy of the tree that does not preserve modification times (the GPU box's snapshot) must not
    trigger a rebuild; a changed source must."""
    from dart_env_b200 import build as b
    if not os.path.exists(b.STAMP):
        pytest.skip("library built without a stamp")
    assert not b.is_stale()
    src = os.path.join(b.CSRC, "lower.h")
    st = os.stat(src)
    try:
        os.utime(src)                       # newer than the library, same content
        assert not b.is_stale()
    finally:
        os.utime(src, (st.st_atime, st.st_mtime))
    monkeypatch.setattr(b, "source_hash", lambda: "0" * 64)
    assert b.is_stale()
