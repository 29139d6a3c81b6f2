"""Minimal MJCF reader: just enough of the MuJoCo compiler to recover kinematics.

The reference parses ``assets/shelf_pnp.xml`` through ``mujoco.MjModel.from_xml_path``
(/root/reference/panda_mujoco_gym/envs/panda_env.py:108).  MuJoCo is not installable in
this image, so this module reads the same file format and exposes the subset of
``MjModel`` fields the IK / reward hot path consumes:

    body_parentid, body_pos, body_quat, body_jntadr, body_jntnum,
    jnt_type, jnt_axis, jnt_pos, jnt_qposadr, jnt_dofadr, jnt_bodyid, jnt_range, qpos0,
    site_bodyid, site_pos, site_quat, nq, nv, opt.timestep, .site(name).id ...

Handled MJCF features (everything the reference's two files use for kinematics):
``<include file=...>``, ``<compiler angle/eulerseq/autolimits>``, nested ``<default class>``
with inheritance, ``childclass`` on bodies, ``class`` on joints/sites, ``<freejoint/>``,
``pos`` / ``quat`` / ``euler`` / ``axisangle`` orientations, quaternion normalisation.
Geoms, meshes, actuators, equality constraints and inertials are ignored: they do not
move the kinematic tree.
"""

from __future__ import annotations

import math
import os
import xml.etree.ElementTree as ET
from types import SimpleNamespace
from typing import Dict, List, Optional

import numpy as np

# mjtJoint enum values (mujoco/mjmodel.h)
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
_JNT_CODES = {"free": JNT_FREE, "ball": JNT_BALL, "slide": JNT_SLIDE, "hinge": JNT_HINGE}
_JNT_NQ = {JNT_FREE: 7, JNT_BALL: 4, JNT_SLIDE: 1, JNT_HINGE: 1}
_JNT_NV = {JNT_FREE: 6, JNT_BALL: 3, JNT_SLIDE: 1, JNT_HINGE: 1}


class MjcfError(ValueError):
    """Raised for MJCF constructs this reader cannot interpret."""


def _floats(text: str, n: Optional[int] = None) -> np.ndarray:
    vals = np.array([float(t) for t in text.split()], dtype=np.float64)
    if n is not None and vals.size != n:
        raise MjcfError(f"expected {n} numbers, got {vals.size}: {text!r}")
    return vals


def _quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array(
        [
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
        ]
    )


def _axisangle_quat(axis: np.ndarray, angle: float) -> np.ndarray:
    n = np.linalg.norm(axis)
    if n < 1e-14:
        return np.array([1.0, 0.0, 0.0, 0.0])
    h = 0.5 * angle
    return np.concatenate([[math.cos(h)], math.sin(h) * axis / n])


def _expand_includes(elem: ET.Element, base_dir: str, depth: int = 0) -> None:
    """Replace every <include file=.../> by the children of the included file's root."""
    if depth > 16:
        raise MjcfError("include nesting too deep")
    i = 0
    while i < len(elem):
        child = elem[i]
        if child.tag == "include":
            path = os.path.join(base_dir, child.attrib["file"])
            sub_root = ET.parse(path).getroot()
            _expand_includes(sub_root, os.path.dirname(path), depth + 1)
            elem.remove(child)
            for k, sub in enumerate(list(sub_root)):
                elem.insert(i + k, sub)
            i += len(sub_root)
        else:
            _expand_includes(child, base_dir, depth + 1)
            i += 1


class _Defaults:
    """Default-class table: class name -> {tag -> attrib dict}, with parent inheritance."""

    def __init__(self) -> None:
        self.classes: Dict[str, Dict[str, Dict[str, str]]] = {"main": {}}

    def ingest(self, node: ET.Element, parent: Optional[str]) -> None:
        name = node.attrib.get("class", "main" if parent is None else None)
        if name is None:
            raise MjcfError("nested <default> needs a class name")
        table = {t: dict(a) for t, a in self.classes.get(parent or "main", {}).items()} if parent else {}
        table.update({t: dict(a) for t, a in self.classes.get(name, {}).items()})
        for child in node:
            if child.tag == "default":
                continue
            merged = dict(table.get(child.tag, {}))
            merged.update(child.attrib)
            table[child.tag] = merged
        self.classes[name] = table
        for child in node:
            if child.tag == "default":
                self.ingest(child, name)

    def resolve(self, tag: str, elem: ET.Element, childclass: Optional[str]) -> Dict[str, str]:
        cls = elem.attrib.get("class", childclass or "main")
        if cls not in self.classes:
            raise MjcfError(f"unknown default class {cls!r}")
        out = dict(self.classes[cls].get(tag, {}))
        out.update({k: v for k, v in elem.attrib.items() if k != "class"})
        return out


class _Named:
    """``model.site(name)`` / ``model.body(name)`` / ``model.joint(name)`` accessor result."""

    def __init__(self, idx: int, name: str) -> None:
        self.id = idx
        self.name = name


class KinematicModel:
    """Kinematics-only stand-in for ``mujoco.MjModel`` (same field names and shapes)."""

    def __init__(self) -> None:
        self.body_names: List[str] = []
        self.joint_names: List[str] = []
        self.site_names: List[str] = []
        self.opt = SimpleNamespace(timestep=0.002)
        self.source_path: Optional[str] = None

    # --- construction -----------------------------------------------------------------
    @classmethod
    def from_xml_path(cls, path: str) -> "KinematicModel":
        path = os.path.abspath(path)
        root = ET.parse(path).getroot()
        if root.tag != "mujoco":
            raise MjcfError("root element must be <mujoco>")
        _expand_includes(root, os.path.dirname(path))
        self = cls()
        self.source_path = path
        self._compile(root)
        return self

    def _orientation(self, attrib: Dict[str, str]) -> np.ndarray:
        if "quat" in attrib:
            q = _floats(attrib["quat"], 4)
        elif "euler" in attrib:
            e = _floats(attrib["euler"], 3) * self._angle_scale
            q = np.array([1.0, 0.0, 0.0, 0.0])
            for ch, ang in zip(self._eulerseq, e):
                ax = np.eye(3)["xyz".index(ch.lower())]
                step = _axisangle_quat(ax, ang)
                # lower-case = intrinsic (post-multiply), upper-case = extrinsic (pre-multiply)
                q = _quat_mul(q, step) if ch.islower() else _quat_mul(step, q)
        elif "axisangle" in attrib:
            a = _floats(attrib["axisangle"], 4)
            q = _axisangle_quat(a[:3], a[3] * self._angle_scale)
        elif any(k in attrib for k in ("xyaxes", "zaxis")):
            raise MjcfError("xyaxes/zaxis orientations are not supported by this reader")
        else:
            q = np.array([1.0, 0.0, 0.0, 0.0])
        n = np.linalg.norm(q)
        if n < 1e-14:
            raise MjcfError("zero quaternion")
        return q / n

    def _compile(self, root: ET.Element) -> None:
        compiler: Dict[str, str] = {}
        for c in root.iter("compiler"):
            compiler.update(c.attrib)
        self._angle_scale = 1.0 if compiler.get("angle", "degree") == "radian" else math.pi / 180.0
        self._eulerseq = compiler.get("eulerseq", "xyz")
        for o in root.iter("option"):
            if "timestep" in o.attrib:
                self.opt.timestep = float(o.attrib["timestep"])

        defaults = _Defaults()
        for d in root.findall("default"):
            defaults.ingest(d, None)

        body_parent, body_pos, body_quat, body_jntadr, body_jntnum = [0], [np.zeros(3)], [np.array([1.0, 0, 0, 0])], [-1], [0]
        self.body_names.append("world")
        jnt = dict(type=[], axis=[], pos=[], qposadr=[], dofadr=[], bodyid=[], range=[], limited=[])
        qpos0: List[float] = []
        site = dict(bodyid=[], pos=[], quat=[])
        nv = 0

        def add_joint(attrib: Dict[str, str], bid: int, free: bool, name: str) -> None:
            nonlocal nv
            jt = JNT_FREE if free else _JNT_CODES[attrib.get("type", "hinge")]
            axis = _floats(attrib.get("axis", "0 0 1"), 3)
            if jt in (JNT_HINGE, JNT_SLIDE):
                axis = axis / np.linalg.norm(axis)
            else:
                axis = np.array([0.0, 0.0, 1.0])
            rng = _floats(attrib["range"], 2) if "range" in attrib else np.zeros(2)
            ref = float(attrib.get("ref", "0"))
            if jt == JNT_HINGE:
                rng = rng * self._angle_scale
                ref *= self._angle_scale
            jnt["type"].append(jt)
            jnt["axis"].append(axis)
            jnt["pos"].append(_floats(attrib.get("pos", "0 0 0"), 3) if jt != JNT_FREE else np.zeros(3))
            jnt["qposadr"].append(len(qpos0))
            jnt["dofadr"].append(nv)
            jnt["bodyid"].append(bid)
            jnt["range"].append(rng)
            jnt["limited"].append("range" in attrib and attrib.get("limited", "auto") != "false")
            self.joint_names.append(name)
            if jt == JNT_FREE:
                qpos0.extend(list(body_pos[bid]) + list(body_quat[bid]))
            elif jt == JNT_BALL:
                qpos0.extend([1.0, 0.0, 0.0, 0.0])
            else:
                qpos0.append(ref)
            nv += _JNT_NV[jt]

        def add_site(elem: ET.Element, bid: int, childclass: Optional[str]) -> None:
            a = defaults.resolve("site", elem, childclass)
            site["bodyid"].append(bid)
            site["pos"].append(_floats(a.get("pos", "0 0 0"), 3))
            site["quat"].append(self._orientation(a))
            self.site_names.append(a.get("name", f"site{len(self.site_names)}"))

        def walk(elem: ET.Element, bid: int, childclass: Optional[str]) -> None:
            # MuJoCo numbers bodies depth-first in document order; joints/sites likewise.
            for child in elem:
                if child.tag == "site":
                    add_site(child, bid, childclass)
                elif child.tag == "body":
                    cc = child.attrib.get("childclass", childclass)
                    new_id = len(body_parent)
                    body_parent.append(bid)
                    body_pos.append(_floats(child.attrib.get("pos", "0 0 0"), 3))
                    body_quat.append(self._orientation(child.attrib))
                    body_jntadr.append(-1)
                    body_jntnum.append(0)
                    self.body_names.append(child.attrib.get("name", f"body{new_id}"))
                    for j in child:
                        if j.tag in ("joint", "freejoint"):
                            a = defaults.resolve("joint", j, cc) if j.tag == "joint" else dict(j.attrib)
                            if body_jntnum[new_id] == 0:
                                body_jntadr[new_id] = len(jnt["type"])
                            body_jntnum[new_id] += 1
                            is_free = j.tag == "freejoint" or a.get("type") == "free"
                            add_joint(a, new_id, is_free, a.get("name", f"joint{len(self.joint_names)}"))
                    walk(child, new_id, cc)

        for wb in root.findall("worldbody"):
            walk(wb, 0, None)

        self.nbody = len(body_parent)
        self.body_parentid = np.array(body_parent, dtype=np.int32)
        self.body_pos = np.array(body_pos, dtype=np.float64).reshape(-1, 3)
        self.body_quat = np.array(body_quat, dtype=np.float64).reshape(-1, 4)
        self.body_jntadr = np.array(body_jntadr, dtype=np.int32)
        self.body_jntnum = np.array(body_jntnum, dtype=np.int32)
        self.njnt = len(jnt["type"])
        self.jnt_type = np.array(jnt["type"], dtype=np.int32)
        self.jnt_axis = np.array(jnt["axis"], dtype=np.float64).reshape(-1, 3)
        self.jnt_pos = np.array(jnt["pos"], dtype=np.float64).reshape(-1, 3)
        self.jnt_qposadr = np.array(jnt["qposadr"], dtype=np.int32)
        self.jnt_dofadr = np.array(jnt["dofadr"], dtype=np.int32)
        self.jnt_bodyid = np.array(jnt["bodyid"], dtype=np.int32)
        self.jnt_range = np.array(jnt["range"], dtype=np.float64).reshape(-1, 2)
        self.jnt_limited = np.array(jnt["limited"], dtype=bool)
        self.qpos0 = np.array(qpos0, dtype=np.float64)
        self.nq = int(self.qpos0.size)
        self.nv = int(nv)
        self.nsite = len(site["bodyid"])
        self.site_bodyid = np.array(site["bodyid"], dtype=np.int32)
        self.site_pos = np.array(site["pos"], dtype=np.float64).reshape(-1, 3)
        self.site_quat = np.array(site["quat"], dtype=np.float64).reshape(-1, 4)

    # --- MjModel-style named access ---------------------------------------------------
    def _lookup(self, names: List[str], name: str, kind: str) -> _Named:
        try:
            return _Named(names.index(name), name)
        except ValueError:
            raise KeyError(f"Invalid name {name!r} for {kind}. Valid names: {names}") from None

    def site(self, name: str) -> _Named:
        return self._lookup(self.site_names, name, "site")

    def body(self, name: str) -> _Named:
        return self._lookup(self.body_names, name, "body")

    def joint(self, name: str) -> _Named:
        return self._lookup(self.joint_names, name, "joint")


class KinematicData:
    """Stand-in for ``mujoco.MjData``: the fields the IK controller reads and writes."""

    def __init__(self, model: KinematicModel) -> None:
        self.qpos = model.qpos0.copy()
        self.qvel = np.zeros(model.nv)
        self.site_xpos = np.zeros((model.nsite, 3))
        self.site_xmat = np.tile(np.eye(3).reshape(9), (model.nsite, 1))
        self.time = 0.0
