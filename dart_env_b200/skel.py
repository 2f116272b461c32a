"""`.skel` model compiler (host side).

Parses a DART `.skel` XML file into a flat, topologically ordered :class:`Model`
that the C-ABI (`include/dartb.h: dartb_model_t`) and the CPU oracle both consume.

What it replaces in the reference: the ``pydart.World(dt, skel_path)`` call at
``gym/envs/dart/dart_env.py:54-55`` (reference) which hands the file to DART's C++
``utils::SkelParser``.  The semantics restated here:

* ``<transformation>`` = ``x y z rx ry rz`` with ``R = Rx(rx) * Ry(ry) * Rz(rz)``.
* body transforms are world poses at q = 0 (pre-multiplied by the skeleton frame).
* joint ``<transformation>`` = child-body -> joint frame (default identity);
  ``parent_to_joint = parentWorld^-1 * childWorld * child_to_joint``.
* bodies / dofs are created by walking joints in *file order*, creating a joint's
  parent first when it has not been created yet.
* moment of inertia, when not given explicitly, is ``shape0.computeInertia(mass)`` of
  the body's FIRST shape, evaluated in the shape's own frame (the shape's local
  transform is ignored) and taken about the COM (``<inertia><offset>``).
* ``<mobile>false</mobile>`` skeletons are static: their collision shapes become
  world-fixed shapes.
* the robot skeleton is the LAST skeleton (``dart_env.py:62``).

No GL, no pydart2, no lxml: plain ``xml.etree``.
"""
from __future__ import annotations

import json
import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

JOINT_WELD, JOINT_REVOLUTE, JOINT_PRISMATIC = 0, 1, 2
SHAPE_BOX, SHAPE_CAPSULE, SHAPE_SPHERE, SHAPE_ELLIPSOID, SHAPE_CYLINDER = 0, 1, 2, 3, 4
_JOINT_NAMES = {"weld": JOINT_WELD, "revolute": JOINT_REVOLUTE, "prismatic": JOINT_PRISMATIC}

DEFAULT_FRICTION_COEFF = 1.0  # DART_DEFAULT_FRICTION_COEFF


class SkelError(ValueError):
    """Raised for malformed or unsupported `.skel` content."""


# ----------------------------------------------------------------------------- math
def euler_xyz_to_matrix(rx: float, ry: float, rz: float) -> np.ndarray:
    """DART ``math::eulerXYZToMatrix``: R = Rx * Ry * Rz."""
    cx, sx = math.cos(rx), math.sin(rx)
    cy, sy = math.cos(ry), math.sin(ry)
    cz, sz = math.cos(rz), math.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=np.float64)
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=np.float64)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dtype=np.float64)
    return Rx @ Ry @ Rz


def make_transform(vals) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = euler_xyz_to_matrix(vals[3], vals[4], vals[5])
    T[:3, 3] = vals[:3]
    return T


def inv_transform(T: np.ndarray) -> np.ndarray:
    Ti = np.eye(4)
    Ti[:3, :3] = T[:3, :3].T
    Ti[:3, 3] = -T[:3, :3].T @ T[:3, 3]
    return Ti


def shape_inertia(stype: int, size, mass: float) -> np.ndarray:
    """Moment of inertia about the shape's own centre, in the shape's own frame."""
    I = np.zeros((3, 3))
    if stype == SHAPE_BOX:
        x, y, z = size[:3]
        I[0, 0] = mass / 12.0 * (y * y + z * z)
        I[1, 1] = mass / 12.0 * (x * x + z * z)
        I[2, 2] = mass / 12.0 * (x * x + y * y)
    elif stype == SHAPE_ELLIPSOID:  # size = diameters
        a, b, c = 0.5 * size[0], 0.5 * size[1], 0.5 * size[2]
        I[0, 0] = mass / 5.0 * (b * b + c * c)
        I[1, 1] = mass / 5.0 * (a * a + c * c)
        I[2, 2] = mass / 5.0 * (a * a + b * b)
    elif stype == SHAPE_SPHERE:
        r = size[0]
        I[0, 0] = I[1, 1] = I[2, 2] = 0.4 * mass * r * r
    elif stype == SHAPE_CYLINDER:
        r, h = size[0], size[1]
        I[0, 0] = I[1, 1] = mass * (3.0 * r * r + h * h) / 12.0
        I[2, 2] = 0.5 * mass * r * r
    elif stype == SHAPE_CAPSULE:  # axis = local z, cylinder height h, radius r
        r, h = size[0], size[1]
        r2 = r * r
        vol_c = math.pi * r2 * h
        vol_s = math.pi * r2 * r * 4.0 / 3.0
        rho = mass / (vol_c + vol_s)
        mc, ms = rho * vol_c, rho * vol_s
        ixx = mc * (h * h / 12.0 + r2 / 4.0) + ms * (0.4 * r2 + 0.25 * h * h + 0.375 * h * r)
        izz = mc * r2 / 2.0 + ms * 0.4 * r2
        I[0, 0] = I[1, 1] = ixx
        I[2, 2] = izz
    else:
        raise SkelError("unsupported shape type %r" % stype)
    return I


# ----------------------------------------------------------------------------- data
@dataclass
class Shape:
    type: int
    size: np.ndarray          # box: xyz; capsule/cylinder: (radius, height, 0); sphere: (r,0,0)
    T: np.ndarray             # 4x4, body-local (robot) or world (static)
    body: int = -1            # robot body index, -1 for world-fixed


@dataclass
class Body:
    name: str
    parent: int               # -1 = world
    joint_name: str
    joint_type: int
    dof: int                  # index into q, -1 for weld
    T_parent_joint: np.ndarray
    T_child_joint: np.ndarray
    axis: np.ndarray          # in the joint frame, normalised
    q_lo: float = -math.inf
    q_hi: float = math.inf
    has_limit: bool = False
    limit_enforced: bool = False
    damping: float = 0.0
    coulomb: float = 0.0
    spring_k: float = 0.0
    spring_rest: float = 0.0
    q_init: float = 0.0
    dq_init: float = 0.0
    mass: float = 0.0
    com: np.ndarray = field(default_factory=lambda: np.zeros(3))
    inertia: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))  # about COM
    friction_coeff: float = DEFAULT_FRICTION_COEFF
    T_world0: np.ndarray = field(default_factory=lambda: np.eye(4))


@dataclass
class Model:
    name: str
    dt: float
    gravity: np.ndarray
    bodies: List[Body]
    shapes: List[Shape]        # robot collision shapes (body-local)
    ground: List[Shape]        # world-fixed collision shapes
    source: str = ""
    n_static_skeletons: int = 1   # immobile skeletons before the robot (world.skeletons indexing)

    @property
    def n_bodies(self) -> int:
        return len(self.bodies)

    @property
    def n_dofs(self) -> int:
        return sum(1 for b in self.bodies if b.dof >= 0)

    def dof_bodies(self) -> List[int]:
        out = [-1] * self.n_dofs
        for i, b in enumerate(self.bodies):
            if b.dof >= 0:
                out[b.dof] = i
        return out

    def q_lower(self) -> np.ndarray:
        return np.array([self.bodies[i].q_lo for i in self.dof_bodies()])

    def q_upper(self) -> np.ndarray:
        return np.array([self.bodies[i].q_hi for i in self.dof_bodies()])

    def q_init(self) -> np.ndarray:
        return np.array([self.bodies[i].q_init for i in self.dof_bodies()])

    def dq_init(self) -> np.ndarray:
        return np.array([self.bodies[i].dq_init for i in self.dof_bodies()])

    def enforce_limits(self) -> None:
        """`dart_env.py:64-67`: enable enforcement on every dof that has a limit."""
        for b in self.bodies:
            if b.dof >= 0 and b.has_limit:
                b.limit_enforced = True

    # --- (de)serialisation: engine-agnostic compiled model ---------------------------
    def to_dict(self) -> dict:
        def arr(a):
            return np.asarray(a, dtype=np.float64).tolist()

        def inf(x):
            return "inf" if x == math.inf else "-inf" if x == -math.inf else x

        return {
            "name": self.name, "dt": self.dt, "gravity": arr(self.gravity), "source": self.source,
            "n_static_skeletons": self.n_static_skeletons,
            "bodies": [{
                "name": b.name, "parent": b.parent, "joint_name": b.joint_name,
                "joint_type": b.joint_type, "dof": b.dof,
                "T_parent_joint": arr(b.T_parent_joint), "T_child_joint": arr(b.T_child_joint),
                "axis": arr(b.axis), "q_lo": inf(b.q_lo), "q_hi": inf(b.q_hi),
                "has_limit": b.has_limit, "limit_enforced": b.limit_enforced,
                "damping": b.damping, "coulomb": b.coulomb, "spring_k": b.spring_k,
                "spring_rest": b.spring_rest, "q_init": b.q_init, "dq_init": b.dq_init,
                "mass": b.mass, "com": arr(b.com), "inertia": arr(b.inertia),
                "friction_coeff": b.friction_coeff, "T_world0": arr(b.T_world0),
            } for b in self.bodies],
            "shapes": [{"type": s.type, "size": arr(s.size), "T": arr(s.T), "body": s.body}
                       for s in self.shapes],
            "ground": [{"type": s.type, "size": arr(s.size), "T": arr(s.T), "body": -1}
                       for s in self.ground],
        }

    @staticmethod
    def from_dict(d: dict) -> "Model":
        def f(x):
            return math.inf if x == "inf" else -math.inf if x == "-inf" else float(x)

        bodies = []
        for b in d["bodies"]:
            bodies.append(Body(
                name=b["name"], parent=b["parent"], joint_name=b["joint_name"],
                joint_type=b["joint_type"], dof=b["dof"],
                T_parent_joint=np.array(b["T_parent_joint"]), T_child_joint=np.array(b["T_child_joint"]),
                axis=np.array(b["axis"]), q_lo=f(b["q_lo"]), q_hi=f(b["q_hi"]),
                has_limit=b["has_limit"], limit_enforced=b["limit_enforced"],
                damping=b["damping"], coulomb=b["coulomb"], spring_k=b["spring_k"],
                spring_rest=b["spring_rest"], q_init=b["q_init"], dq_init=b["dq_init"],
                mass=b["mass"], com=np.array(b["com"]), inertia=np.array(b["inertia"]),
                friction_coeff=b["friction_coeff"], T_world0=np.array(b["T_world0"])))
        mk = lambda s: Shape(type=s["type"], size=np.array(s["size"]), T=np.array(s["T"]), body=s["body"])
        return Model(name=d["name"], dt=d["dt"], gravity=np.array(d["gravity"]), bodies=bodies,
                     shapes=[mk(s) for s in d["shapes"]], ground=[mk(s) for s in d["ground"]],
                     source=d.get("source", ""), n_static_skeletons=d.get("n_static_skeletons", 1))

    def save_json(self, path: str) -> None:
        with open(path, "w") as fh:
            json.dump(self.to_dict(), fh, indent=1)

    @staticmethod
    def load_json(path: str) -> "Model":
        with open(path) as fh:
            return Model.from_dict(json.load(fh))


# ----------------------------------------------------------------------------- parser
def _floats(text: Optional[str], n: Optional[int] = None) -> np.ndarray:
    if text is None:
        raise SkelError("missing numeric text")
    v = np.array([float(t) for t in text.split()], dtype=np.float64)
    if n is not None and v.size != n:
        raise SkelError("expected %d numbers, got %r" % (n, text))
    return v


def _child_transform(elem) -> np.ndarray:
    t = elem.find("transformation")
    return make_transform(_floats(t.text, 6)) if t is not None else np.eye(4)


def _read_geometry(shape_elem):
    geo = shape_elem.find("geometry")
    if geo is None or len(geo) == 0:
        raise SkelError("shape without <geometry>")
    g = geo[0]
    tag = g.tag
    if tag == "box":
        return SHAPE_BOX, _floats(g.find("size").text, 3)
    if tag == "ellipsoid":
        return SHAPE_ELLIPSOID, _floats(g.find("size").text, 3)
    if tag == "sphere":
        return SHAPE_SPHERE, np.array([float(g.find("radius").text), 0.0, 0.0])
    if tag in ("capsule", "cylinder"):
        st = SHAPE_CAPSULE if tag == "capsule" else SHAPE_CYLINDER
        return st, np.array([float(g.find("radius").text), float(g.find("height").text), 0.0])
    raise SkelError("unsupported geometry <%s> (mesh / plane / multi_sphere are out of scope)" % tag)


def _read_body(body_elem, skel_frame):
    name = body_elem.get("name")
    Tw = skel_frame @ _child_transform(body_elem)
    mass, com, moment = 0.0, np.zeros(3), None
    inert = body_elem.find("inertia")
    if inert is not None:
        m = inert.find("mass")
        if m is not None:
            mass = float(m.text)
        off = inert.find("offset")
        if off is not None:
            com = _floats(off.text, 3)
        moi = inert.find("moment_of_inertia")
        if moi is not None:
            g = lambda k: float(moi.find(k).text) if moi.find(k) is not None else 0.0
            ixx, iyy, izz, ixy, ixz, iyz = g("ixx"), g("iyy"), g("izz"), g("ixy"), g("ixz"), g("iyz")
            moment = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])
    vis, col = [], []
    for child in body_elem:
        if child.tag == "visualization_shape":
            try:
                st, size = _read_geometry(child)
            except SkelError:
                continue  # meshes only matter for rendering
            vis.append(Shape(st, size, _child_transform(child)))
        elif child.tag == "collision_shape":
            st, size = _read_geometry(child)
            col.append(Shape(st, size, _child_transform(child)))
    if moment is None:
        first = vis[0] if vis else (col[0] if col else None)  # DART: first shape node of the body
        moment = shape_inertia(first.type, first.size, mass) if first is not None else np.zeros((3, 3))
    return dict(name=name, Tw=Tw, mass=mass, com=com, moment=moment, col=col)


def _read_joint(j):
    jt = j.get("type")
    d = dict(name=j.get("name"), type=jt, parent=j.find("parent").text.strip(),
             child=j.find("child").text.strip(), T_child_joint=_child_transform(j),
             axis=np.array([0.0, 0.0, 1.0]), q_lo=-math.inf, q_hi=math.inf, has_limit=False,
             damping=0.0, coulomb=0.0, spring_k=0.0, spring_rest=0.0, q_init=0.0, dq_init=0.0)
    ax = j.find("axis")
    if ax is not None:
        xyz = ax.find("xyz")
        if xyz is not None:
            v = _floats(xyz.text, 3)
            d["axis"] = v / np.linalg.norm(v)
        lim = ax.find("limit")
        if lim is not None:
            lo, hi = lim.find("lower"), lim.find("upper")
            if lo is not None:
                d["q_lo"] = float(lo.text)
            if hi is not None:
                d["q_hi"] = float(hi.text)
            d["has_limit"] = True
        dyn = ax.find("dynamics")
        if dyn is not None:
            for key, tag in (("damping", "damping"), ("coulomb", "friction"),
                             ("spring_k", "spring_stiffness"), ("spring_rest", "spring_rest_position")):
                e = dyn.find(tag)
                if e is not None:
                    d[key] = float(e.text)
    for key, tag in (("q_init", "init_pos"), ("dq_init", "init_vel")):
        e = j.find(tag)
        if e is not None and e.text and e.text.strip():
            d[key] = float(e.text.split()[0])
    return d


def parse_skel(path: str, dt: Optional[float] = None) -> Model:
    """Compile a `.skel` file; `dt` overrides `<time_step>` like `pydart.World(dt, path)`."""
    if not os.path.exists(path):
        raise IOError("File %s does not exist" % path)
    root = ET.parse(path).getroot()
    world = root.find("world")
    if world is None:
        raise SkelError("no <world> element")
    phys = world.find("physics")
    time_step, gravity = 0.001, np.array([0.0, -9.81, 0.0])
    if phys is not None:
        if phys.find("time_step") is not None:
            time_step = float(phys.find("time_step").text)
        if phys.find("gravity") is not None:
            gravity = _floats(phys.find("gravity").text, 3)
    if dt is not None:
        time_step = float(dt)

    skeletons = world.findall("skeleton")
    if not skeletons:
        raise SkelError("no <skeleton>")
    ground: List[Shape] = []
    robot_elem = skeletons[-1]
    for sk in skeletons[:-1]:
        mobile = sk.find("mobile")
        is_mobile = True if mobile is None else mobile.text.strip().lower() not in ("false", "0")
        if is_mobile:
            raise SkelError("only one mobile skeleton (the last) is supported; %r is mobile" % sk.get("name"))
        frame = _child_transform(sk)
        for b in sk.findall("body"):
            info = _read_body(b, frame)
            for s in info["col"]:
                ground.append(Shape(s.type, s.size, info["Tw"] @ s.T, -1))

    frame = _child_transform(robot_elem)
    body_info = {}
    for b in robot_elem.findall("body"):
        info = _read_body(b, frame)
        body_info[info["name"]] = info
    joints = [_read_joint(j) for j in robot_elem.findall("joint")]
    by_child = {j["child"]: j for j in joints}

    bodies: List[Body] = []
    index = {}
    shapes: List[Shape] = []
    ndof = 0

    def create(j):
        nonlocal ndof
        if j["child"] in index:
            return
        if j["parent"] != "world" and j["parent"] not in index:
            if j["parent"] not in by_child:
                raise SkelError("body %r has no parent joint" % j["parent"])
            create(by_child[j["parent"]])
        if j["type"] not in _JOINT_NAMES:
            raise SkelError("joint type %r is out of scope (only weld/revolute/prismatic)" % j["type"])
        jt = _JOINT_NAMES[j["type"]]
        info = body_info[j["child"]]
        parent = -1 if j["parent"] == "world" else index[j["parent"]]
        parent_world = np.eye(4) if parent < 0 else bodies[parent].T_world0
        T_pj = inv_transform(parent_world) @ info["Tw"] @ j["T_child_joint"]
        dof = -1
        if jt != JOINT_WELD:
            dof = ndof
            ndof += 1
        b = Body(name=info["name"], parent=parent, joint_name=j["name"], joint_type=jt, dof=dof,
                 T_parent_joint=T_pj, T_child_joint=j["T_child_joint"], axis=j["axis"],
                 q_lo=j["q_lo"], q_hi=j["q_hi"], has_limit=j["has_limit"] and jt != JOINT_WELD,
                 damping=j["damping"], coulomb=j["coulomb"], spring_k=j["spring_k"],
                 spring_rest=j["spring_rest"], q_init=j["q_init"], dq_init=j["dq_init"],
                 mass=info["mass"], com=info["com"], inertia=info["moment"], T_world0=info["Tw"])
        index[info["name"]] = len(bodies)
        bodies.append(b)
        for s in info["col"]:
            shapes.append(Shape(s.type, s.size, s.T, index[info["name"]]))

    for j in joints:
        create(j)
    return Model(name=robot_elem.get("name", "robot"), dt=time_step, gravity=gravity, bodies=bodies,
                 shapes=shapes, ground=ground, source=os.path.basename(path), n_static_skeletons=len(skeletons) - 1)


# ----------------------------------------------------------------------------- asset lookup
_BUNDLED = os.path.join(os.path.dirname(__file__), "assets")


def find_asset(model_path: str) -> str:
    """Resolve a model path like `dart_env.py:44-52`: absolute paths as-is, otherwise relative to
    $DART_ENV_ASSETS (a directory of `.skel` files, e.g. the reference's gym/envs/dart/assets) when it
    is set, else to the bundled assets.  Nothing else is searched: behaviour does not depend on whether
    a reference checkout happens to be on the machine."""
    if model_path.startswith("/"):
        return model_path
    base = os.environ.get("DART_ENV_ASSETS")
    if base and os.path.exists(os.path.join(base, model_path)):
        return os.path.join(base, model_path)
    return os.path.join(_BUNDLED, model_path)


def load_model(model_path: str, dt: Optional[float] = None) -> Model:
    """Load `<name>.skel` from the assets dir, or the bundled pre-compiled `<name>.model.json`
    (generated by tools/export_models.py) when the `.skel` itself is not on this machine."""
    full = find_asset(model_path)
    if os.path.exists(full) and full.endswith(".skel"):
        return parse_skel(full, dt)
    if full.endswith(".skel"):
        js = os.path.join(_BUNDLED, os.path.basename(full)[:-5] + ".model.json")
        if os.path.exists(js):
            m = Model.load_json(js)
            if dt is not None:
                m.dt = float(dt)
            return m
        raise IOError("File %s does not exist" % full)
    raise SkelError("only .skel models are supported (URDF/SDF loading is not implemented): %s" % full)
