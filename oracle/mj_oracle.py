"""ORACLE (test infrastructure, never imported by the product package).

FP64 NumPy restatement of the MuJoCo routines the reference's hot path calls:

    mujoco.MjModel.from_xml_path  /root/reference/panda_mujoco_gym/envs/panda_env.py:108
    mujoco.mj_kinematics          /root/reference/panda_mujoco_gym/skills/ik_solver.py:58
    mujoco.mj_forward             ik_solver.py:52,83 (only its position stage matters to IK)
    mujoco.mj_jacSite             ik_solver.py:72
    mujoco.mju_mat2Quat           panda_env.py:337-342

MuJoCo itself (pinned ``mujoco==2.3.3``, /root/reference/requirements.txt:1) is a third-party
dependency that is absent from /root/reference and cannot be installed in this image, so its
published algorithm is restated here (engine_core_smooth.c ``mj_kinematics``,
engine_support.c ``mj_jac``, engine_util_spatial.c ``mju_mat2Quat``).

Pinning status: the only real-MuJoCo number that exists inside the reference for this path is
``home_wpt = [1.23843967, 0.0, 0.49740014]`` (scripts/execute_pnp.py:38, test/reward_test.py:47),
the EE-site position at the neutral pose; tests/test_oracle.py checks it, plus central
finite differences for the Jacobian.  Beyond that: PARITY UNPINNED against real MuJoCo.

This file deliberately has its own MJCF walk (quaternion based, whole body tree, world poses
for every body and site) so that it is independent of the product's chain extraction in
mujoco_panda_pnp_b200/tree.py.
"""

from __future__ import annotations

import math
import os
import xml.etree.ElementTree as ET

import numpy as np

mjJNT_FREE, mjJNT_BALL, mjJNT_SLIDE, mjJNT_HINGE = 0, 1, 2, 3


# ---------------------------------------------------------------------------------------
# quaternion helpers (mju_* semantics, wxyz)
# ---------------------------------------------------------------------------------------
def mju_mulQuat(a, b):
    return np.array(
        [
            a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
            a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
            a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
            a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0],
        ]
    )


def mju_quat2Mat(q):
    q00, q01, q02, q03 = q[0] * q[0], q[0] * q[1], q[0] * q[2], q[0] * q[3]
    q11, q12, q13 = q[1] * q[1], q[1] * q[2], q[1] * q[3]
    q22, q23, q33 = q[2] * q[2], q[2] * q[3], q[3] * q[3]
    return np.array(
        [
            [q00 + q11 - q22 - q33, 2 * (q12 - q03), 2 * (q13 + q02)],
            [2 * (q12 + q03), q00 - q11 + q22 - q33, 2 * (q23 - q01)],
            [2 * (q13 - q02), 2 * (q23 + q01), q00 - q11 - q22 + q33],
        ]
    )


def mju_rotVecQuat(v, q):
    return mju_quat2Mat(q) @ np.asarray(v, dtype=np.float64)


def mju_axisAngle2Quat(axis, angle):
    s = math.sin(0.5 * angle)
    return np.array([math.cos(0.5 * angle), axis[0] * s, axis[1] * s, axis[2] * s])


def mju_normalize4(q):
    n = math.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])
    if n < 1e-15:
        return np.array([1.0, 0.0, 0.0, 0.0])
    return q / n


def mju_mat2Quat(quat_out, mat):
    """engine_util_spatial.c mju_mat2Quat: branch on the trace / largest diagonal, normalise.

    ``mat`` is the 9-vector (row-major) the reference passes (panda_env.py:339-341).
    Writes wxyz into ``quat_out`` like the C function.
    """
    m = np.asarray(mat, dtype=np.float64).reshape(9)
    q = np.zeros(4)
    if m[0] + m[4] + m[8] > 0:
        q[0] = 0.5 * math.sqrt(1 + m[0] + m[4] + m[8])
        q[1] = 0.25 * (m[7] - m[5]) / q[0]
        q[2] = 0.25 * (m[2] - m[6]) / q[0]
        q[3] = 0.25 * (m[3] - m[1]) / q[0]
    elif m[0] > m[4] and m[0] > m[8]:
        q[1] = 0.5 * math.sqrt(1 + m[0] - m[4] - m[8])
        q[0] = 0.25 * (m[7] - m[5]) / q[1]
        q[2] = 0.25 * (m[1] + m[3]) / q[1]
        q[3] = 0.25 * (m[2] + m[6]) / q[1]
    elif m[4] > m[8]:
        q[2] = 0.5 * math.sqrt(1 - m[0] + m[4] - m[8])
        q[0] = 0.25 * (m[2] - m[6]) / q[2]
        q[1] = 0.25 * (m[1] + m[3]) / q[2]
        q[3] = 0.25 * (m[5] + m[7]) / q[2]
    else:
        q[3] = 0.5 * math.sqrt(1 - m[0] - m[4] + m[8])
        q[0] = 0.25 * (m[3] - m[1]) / q[3]
        q[1] = 0.25 * (m[2] + m[6]) / q[3]
        q[2] = 0.25 * (m[5] + m[7]) / q[3]
    quat_out[:] = mju_normalize4(q)


# ---------------------------------------------------------------------------------------
# model / data
# ---------------------------------------------------------------------------------------
class _Handle:
    def __init__(self, i, name):
        self.id, self.name = i, name


class MjModel:
    """The slice of mujoco.MjModel the hot path touches, compiled from MJCF."""

    @staticmethod
    def from_xml_path(path):
        return _compile_mjcf(path)

    def _named(self, names, name):
        if name not in names:
            raise KeyError(name)
        return _Handle(names.index(name), name)

    def site(self, name):
        return self._named(self.names_site, name)

    def body(self, name):
        return self._named(self.names_body, name)

    def joint(self, name):
        return self._named(self.names_joint, name)


class MjData:
    def __init__(self, model):
        self.qpos = model.qpos0.copy()
        self.qvel = np.zeros(model.nv)
        self.xpos = np.zeros((model.nbody, 3))
        self.xquat = np.tile(np.array([1.0, 0, 0, 0]), (model.nbody, 1))
        self.xmat = np.tile(np.eye(3).reshape(9), (model.nbody, 1))
        self.xanchor = np.zeros((model.njnt, 3))
        self.xaxis = np.zeros((model.njnt, 3))
        self.site_xpos = np.zeros((model.nsite, 3))
        self.site_xmat = np.tile(np.eye(3).reshape(9), (model.nsite, 1))
        self.time = 0.0


def _load_xml(path):
    root = ET.parse(path).getroot()
    here = os.path.dirname(os.path.abspath(path))

    def splice(node):
        kids = []
        for k in list(node):
            if k.tag == "include":
                kids.extend(list(_load_xml(os.path.join(here, k.get("file")))))
            else:
                splice(k)
                kids.append(k)
        for k in list(node):
            node.remove(k)
        node.extend(kids)

    splice(root)
    return root


def _compile_mjcf(path):
    root = _load_xml(path)
    comp = {}
    for c in root.findall("compiler"):
        comp.update(c.attrib)
    ang = 1.0 if comp.get("angle", "degree") == "radian" else math.pi / 180.0

    # default classes -> flattened joint/site attribute dictionaries
    cls_attr = {"main": {"joint": {}, "site": {}}}

    def read_defaults(node, parent):
        name = node.get("class") or "main"
        mine = {k: dict(v) for k, v in cls_attr[parent].items()} if parent else {"joint": {}, "site": {}}
        if name in cls_attr:
            for k, v in cls_attr[name].items():
                mine[k].update(v)
        for k in node:
            if k.tag in ("joint", "site"):
                mine[k.tag].update(k.attrib)
        cls_attr[name] = mine
        for k in node:
            if k.tag == "default":
                read_defaults(k, name)

    for d in root.findall("default"):
        read_defaults(d, None)

    def attrs(tag, el, childclass):
        c = el.get("class") or childclass or "main"
        out = dict(cls_attr[c][tag])
        out.update(el.attrib)
        return out

    def vec(s, n):
        v = [float(x) for x in s.split()]
        assert len(v) == n, s
        return np.array(v)

    def quat_of(a):
        if "quat" in a:
            return mju_normalize4(vec(a["quat"], 4))
        assert not any(k in a for k in ("euler", "axisangle", "xyaxes", "zaxis")), "oracle reader: quat only"
        return np.array([1.0, 0.0, 0.0, 0.0])

    m = MjModel()
    m.names_body, m.names_joint, m.names_site = ["world"], [], []
    parent, bpos, bquat, jadr, jnum = [0], [np.zeros(3)], [np.array([1.0, 0, 0, 0])], [-1], [0]
    jtype, jaxis, jpos, jqadr, jdadr, jbody, jrange = [], [], [], [], [], [], []
    qpos0, sbody, spos, squat = [], [], [], []
    nv = [0]

    def visit(node, bid, childclass):
        for k in node:
            if k.tag == "site":
                a = attrs("site", k, childclass)
                sbody.append(bid)
                spos.append(vec(a.get("pos", "0 0 0"), 3))
                squat.append(quat_of(a))
                m.names_site.append(a.get("name", ""))
            elif k.tag == "body":
                cc = k.get("childclass") or childclass
                me = len(parent)
                parent.append(bid)
                bpos.append(vec(k.get("pos", "0 0 0"), 3))
                bquat.append(quat_of(k.attrib))
                jadr.append(-1)
                jnum.append(0)
                m.names_body.append(k.get("name", ""))
                for j in k:
                    if j.tag not in ("joint", "freejoint"):
                        continue
                    a = attrs("joint", j, cc) if j.tag == "joint" else dict(j.attrib)
                    t = "free" if j.tag == "freejoint" else a.get("type", "hinge")
                    code = {"free": 0, "ball": 1, "slide": 2, "hinge": 3}[t]
                    if jnum[me] == 0:
                        jadr[me] = len(jtype)
                    jnum[me] += 1
                    jtype.append(code)
                    ax = vec(a.get("axis", "0 0 1"), 3)
                    jaxis.append(ax / np.linalg.norm(ax) if code >= 2 else np.array([0.0, 0, 1]))
                    jpos.append(vec(a.get("pos", "0 0 0"), 3) if code != 0 else np.zeros(3))
                    jqadr.append(len(qpos0))
                    jdadr.append(nv[0])
                    jbody.append(me)
                    r = vec(a["range"], 2) if "range" in a else np.zeros(2)
                    jrange.append(r * ang if code == 3 else r)
                    m.names_joint.append(a.get("name", ""))
                    if code == 0:
                        qpos0.extend(list(bpos[me]) + list(bquat[me]))
                        nv[0] += 6
                    elif code == 1:
                        qpos0.extend([1.0, 0, 0, 0])
                        nv[0] += 3
                    else:
                        qpos0.append(float(a.get("ref", "0")) * (ang if code == 3 else 1.0))
                        nv[0] += 1
                visit(k, me, cc)

    for wb in root.findall("worldbody"):
        visit(wb, 0, None)

    m.nbody, m.njnt, m.nsite = len(parent), len(jtype), len(sbody)
    m.body_parentid = np.array(parent)
    m.body_pos = np.array(bpos).reshape(-1, 3)
    m.body_quat = np.array(bquat).reshape(-1, 4)
    m.body_jntadr, m.body_jntnum = np.array(jadr), np.array(jnum)
    m.jnt_type = np.array(jtype)
    m.jnt_axis = np.array(jaxis).reshape(-1, 3)
    m.jnt_pos = np.array(jpos).reshape(-1, 3)
    m.jnt_qposadr, m.jnt_dofadr, m.jnt_bodyid = np.array(jqadr), np.array(jdadr), np.array(jbody)
    m.jnt_range = np.array(jrange).reshape(-1, 2)
    m.qpos0 = np.array(qpos0)
    m.nq, m.nv = len(qpos0), nv[0]
    m.site_bodyid = np.array(sbody)
    m.site_pos = np.array(spos).reshape(-1, 3)
    m.site_quat = np.array(squat).reshape(-1, 4)
    return m


# ---------------------------------------------------------------------------------------
# engine routines
# ---------------------------------------------------------------------------------------
def mj_kinematics(m, d):
    """engine_core_smooth.c mj_kinematics, restated."""
    d.xpos[0] = 0.0
    d.xquat[0] = [1.0, 0, 0, 0]
    d.xmat[0] = np.eye(3).reshape(9)
    for i in range(1, m.nbody):
        pid = m.body_parentid[i]
        jn, ja = m.body_jntnum[i], m.body_jntadr[i]
        if jn == 1 and m.jnt_type[ja] == mjJNT_FREE:
            qa = m.jnt_qposadr[ja]
            xpos = d.qpos[qa : qa + 3].copy()
            xquat = mju_normalize4(d.qpos[qa + 3 : qa + 7].copy())
            d.xanchor[ja] = xpos
            d.xaxis[ja] = m.jnt_axis[ja]
        else:
            pmat = d.xmat[pid].reshape(3, 3)
            xpos = d.xpos[pid] + pmat @ m.body_pos[i]
            xquat = mju_mulQuat(d.xquat[pid], m.body_quat[i])
            for j in range(ja, ja + jn):
                qa, jt = m.jnt_qposadr[j], m.jnt_type[j]
                xanchor = mju_rotVecQuat(m.jnt_pos[j], xquat) + xpos
                xaxis = mju_rotVecQuat(m.jnt_axis[j], xquat)
                if jt == mjJNT_SLIDE:
                    xpos = xpos + xaxis * (d.qpos[qa] - m.qpos0[qa])
                elif jt == mjJNT_HINGE:
                    qloc = mju_axisAngle2Quat(m.jnt_axis[j], d.qpos[qa] - m.qpos0[qa])
                    xquat = mju_mulQuat(xquat, qloc)
                    xpos = xanchor - mju_rotVecQuat(m.jnt_pos[j], xquat)
                elif jt == mjJNT_BALL:
                    qloc = mju_normalize4(d.qpos[qa : qa + 4].copy())
                    xquat = mju_mulQuat(xquat, qloc)
                    xpos = xanchor - mju_rotVecQuat(m.jnt_pos[j], xquat)
                else:
                    raise ValueError("free joint must be alone on its body")
                d.xanchor[j] = xanchor
                d.xaxis[j] = xaxis
        xquat = mju_normalize4(xquat)
        d.xpos[i] = xpos
        d.xquat[i] = xquat
        d.xmat[i] = mju_quat2Mat(xquat).reshape(9)
    for s in range(m.nsite):
        b = m.site_bodyid[s]
        bm = d.xmat[b].reshape(3, 3)
        d.site_xpos[s] = d.xpos[b] + bm @ m.site_pos[s]
        d.site_xmat[s] = (bm @ mju_quat2Mat(m.site_quat[s])).reshape(9)


def mj_forward(m, d):
    """Only the position stage of mj_forward influences the IK path (ik_solver.py:52,83):
    collision, constraint and acceleration stages never feed back into site_xpos / jacobians."""
    mj_kinematics(m, d)


def mj_jac(m, d, jacp, jacr, point, body):
    """engine_support.c mj_jac in closed form: for every dof on the path body -> world,
    hinge: jacr = axis, jacp = axis x (point - anchor); slide: jacp = axis;
    free: translational identity + rotational columns about the body's axes;
    ball: the three body axes.  Algebraically identical to MuJoCo's cdof/subtree_com route."""
    if jacp is not None:
        jacp[:] = 0.0
    if jacr is not None:
        jacr[:] = 0.0
    b = body
    while b != 0:
        ja, jn = m.body_jntadr[b], m.body_jntnum[b]
        for j in range(ja, ja + jn):
            da, jt = m.jnt_dofadr[j], m.jnt_type[j]
            if jt == mjJNT_HINGE:
                ax = d.xaxis[j]
                if jacp is not None:
                    jacp[:, da] = np.cross(ax, point - d.xanchor[j])
                if jacr is not None:
                    jacr[:, da] = ax
            elif jt == mjJNT_SLIDE:
                if jacp is not None:
                    jacp[:, da] = d.xaxis[j]
            elif jt == mjJNT_FREE:
                R = d.xmat[b].reshape(3, 3)
                for k in range(3):
                    if jacp is not None:
                        jacp[k, da + k] = 1.0
                        jacp[:, da + 3 + k] = np.cross(R[:, k], point - d.xpos[b])
                    if jacr is not None:
                        jacr[:, da + 3 + k] = R[:, k]
            else:  # ball: rotation axes are the body axes at the anchor
                R = d.xmat[b].reshape(3, 3)
                for k in range(3):
                    if jacp is not None:
                        jacp[:, da + k] = np.cross(R[:, k], point - d.xanchor[j])
                    if jacr is not None:
                        jacr[:, da + k] = R[:, k]
        b = m.body_parentid[b]


def mj_jacSite(m, d, jacp, jacr, site):
    mj_jac(m, d, jacp, jacr, d.site_xpos[site].copy(), m.site_bodyid[site])
