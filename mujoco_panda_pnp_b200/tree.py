"""Kinematic tree extraction: MjModel-like object -> canonical 7-link chain for the GPU.

The reference never materialises the tree; it lets MuJoCo walk ``MjModel`` inside
``mj_kinematics`` / ``mj_jacSite`` (/root/reference/panda_mujoco_gym/skills/ik_solver.py:58,72).
Here the chain world -> ``ee_center_site`` is extracted once, reduced to a canonical form and
uploaded to ``__constant__`` memory (``pnp_set_tree``, include/pnp_b200.h).

Canonical form (all FP64 on the host):

    frame_0 = identity
    for joint i in 0..6:
        A_i     = frame_i * Fixed(pos_i, rot_i)       # joint frame: origin = hinge anchor,
                                                      #              z column = hinge axis
        frame_{i+1} = A_i * Rz(qpos_i - qref_i)
    site = frame_7 * Fixed(ee_pos, ee_rot)

Everything MuJoCo allows on the chain is folded into the ``Fixed`` transforms on the host:
joint-less bodies (link0, hand, ee_center_body), a non-zero ``jnt_pos`` (hinge anchor offset)
and an arbitrary ``jnt_axis`` (conjugated by the rotation that maps z onto the axis).  The
kernels therefore only ever rotate about local z, and the geometric Jacobian column of joint
i is ``z(A_i) x (p_site - origin(A_i))`` which is what ``mj_jacSite`` returns for a hinge
(SURVEY.md appendix B).
"""

from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field
from typing import Any, Optional

import numpy as np

from .mjcf import JNT_HINGE, KinematicModel

N_ARM = 7  # the reference hard-codes qpos[:7] as the arm (ik_solver.py:31-33,51)

DEFAULT_ASSET = os.path.join(os.path.dirname(__file__), "assets", "panda_shelf_kinematic.xml")


def quat_to_mat(q: np.ndarray) -> np.ndarray:
    """wxyz unit quaternion -> 3x3 rotation matrix (mju_quat2Mat layout)."""
    w, x, y, z = q
    return np.array(
        [
            [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
        ]
    )


def _z_to_axis_rotation(axis: np.ndarray) -> np.ndarray:
    """Rotation Q with Q @ ez == axis (identity when axis is already ez)."""
    ez = np.array([0.0, 0.0, 1.0])
    a = axis / np.linalg.norm(axis)
    if np.allclose(a, ez, atol=0, rtol=0):
        return np.eye(3)
    if np.allclose(a, -ez):
        return np.diag([1.0, -1.0, -1.0])
    v = np.cross(ez, a)
    c = float(ez @ a)
    vx = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    return np.eye(3) + vx + vx @ vx / (1.0 + c)


class PnpTreeStruct(ctypes.Structure):
    """Mirror of ``struct PnpTree`` in include/pnp_b200.h (host side, FP64)."""

    _fields_ = [
        ("njoint", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("link_pos", ctypes.c_double * (N_ARM * 3)),
        ("link_rot", ctypes.c_double * (N_ARM * 9)),
        ("ee_pos", ctypes.c_double * 3),
        ("ee_rot", ctypes.c_double * 9),
        ("lower", ctypes.c_double * N_ARM),
        ("upper", ctypes.c_double * N_ARM),
        ("qref", ctypes.c_double * N_ARM),
    ]


@dataclass
class KinematicTree:
    """Canonical chain (see module docstring)."""

    link_pos: np.ndarray  # (7,3)
    link_rot: np.ndarray  # (7,3,3)
    ee_pos: np.ndarray  # (3,)
    ee_rot: np.ndarray  # (3,3)
    lower: np.ndarray  # (7,)
    upper: np.ndarray  # (7,)
    qref: np.ndarray  # (7,)
    site_name: str = "ee_center_site"
    site_id: int = -1
    body_chain: list = field(default_factory=list)

    # --- loaders ---------------------------------------------------------------------
    @classmethod
    def from_mjcf(cls, path: Optional[str] = None, site_name: str = "ee_center_site") -> "KinematicTree":
        """Loader (ii) of SURVEY.md D6: self-contained MJCF reader (no mujoco needed)."""
        return cls.from_mjmodel(KinematicModel.from_xml_path(path or DEFAULT_ASSET), site_name)

    @classmethod
    def from_mjmodel(cls, model: Any, site_name: str = "ee_center_site") -> "KinematicTree":
        """Loader (i): works on a live ``mujoco.MjModel`` and on ``KinematicModel`` alike.

        Reads exactly the fields listed in SURVEY.md section 8b.
        """
        site_id = int(model.site(site_name).id)
        chain = []
        b = int(model.site_bodyid[site_id])
        while b != 0:
            chain.append(b)
            b = int(model.body_parentid[b])
        chain.reverse()

        acc_p, acc_r = np.zeros(3), np.eye(3)
        link_pos, link_rot, lower, upper, qref = [], [], [], [], []
        for b in chain:
            bp = np.asarray(model.body_pos[b], dtype=np.float64)
            br = quat_to_mat(np.asarray(model.body_quat[b], dtype=np.float64))
            acc_p, acc_r = acc_p + acc_r @ bp, acc_r @ br
            adr, num = int(model.body_jntadr[b]), int(model.body_jntnum[b])
            for j in range(adr, adr + num):
                if int(model.jnt_type[j]) != JNT_HINGE:
                    raise ValueError(f"joint {j} on the chain to {site_name!r} is not a hinge")
                k = len(link_pos)
                if int(model.jnt_qposadr[j]) != k:
                    raise ValueError(
                        f"arm joints must be qpos[0:{N_ARM}] in chain order "
                        f"(joint {j} has qposadr {int(model.jnt_qposadr[j])}, expected {k})"
                    )
                anchor = np.asarray(model.jnt_pos[j], dtype=np.float64)
                q_align = _z_to_axis_rotation(np.asarray(model.jnt_axis[j], dtype=np.float64))
                link_pos.append(acc_p + acc_r @ anchor)
                link_rot.append(acc_r @ q_align)
                # after the hinge: undo the axis alignment and the anchor shift
                acc_p, acc_r = -(q_align.T @ anchor), q_align.T.copy()
                lower.append(float(model.jnt_range[j][0]))
                upper.append(float(model.jnt_range[j][1]))
                qref.append(float(model.qpos0[int(model.jnt_qposadr[j])]))
        if len(link_pos) != N_ARM:
            raise ValueError(f"expected {N_ARM} hinge joints on the chain, found {len(link_pos)}")
        sp = np.asarray(model.site_pos[site_id], dtype=np.float64)
        sr = quat_to_mat(np.asarray(model.site_quat[site_id], dtype=np.float64))
        return cls(
            link_pos=np.array(link_pos),
            link_rot=np.array(link_rot),
            ee_pos=acc_p + acc_r @ sp,
            ee_rot=acc_r @ sr,
            lower=np.array(lower),
            upper=np.array(upper),
            qref=np.array(qref),
            site_name=site_name,
            site_id=site_id,
            body_chain=chain,
        )

    # --- packing ---------------------------------------------------------------------
    def to_struct(self) -> PnpTreeStruct:
        s = PnpTreeStruct()
        s.njoint = N_ARM
        s.reserved = 0
        s.link_pos[:] = self.link_pos.reshape(-1).tolist()
        s.link_rot[:] = self.link_rot.reshape(-1).tolist()
        s.ee_pos[:] = self.ee_pos.tolist()
        s.ee_rot[:] = self.ee_rot.reshape(-1).tolist()
        s.lower[:] = self.lower.tolist()
        s.upper[:] = self.upper.tolist()
        s.qref[:] = self.qref.tolist()
        return s

    def snapped(self, tol: float = 1e-12) -> "KinematicTree":
        """Copy with rotation entries within ``tol`` of {0, +-1} snapped exactly.

        MJCF quats such as ``1 1 0 0`` normalise to 1/sqrt(2) whose square is not exactly
        0.5, leaving ~1e-16 residue in the +-90 degree link twists.  The specialised kernel
        instantiation is generated from the snapped tree so that the compiler can drop the
        zero terms; the deviation (<= 3e-16) is far below the FP32 working precision.
        """

        def snap(a: np.ndarray) -> np.ndarray:
            a = a.copy()
            a[np.abs(a) < tol] = 0.0
            a[np.abs(a - 1.0) < tol] = 1.0
            a[np.abs(a + 1.0) < tol] = -1.0
            return a

        return KinematicTree(
            link_pos=snap(self.link_pos),
            link_rot=snap(self.link_rot),
            ee_pos=snap(self.ee_pos),
            ee_rot=snap(self.ee_rot),
            lower=self.lower.copy(),
            upper=self.upper.copy(),
            qref=self.qref.copy(),
            site_name=self.site_name,
            site_id=self.site_id,
            body_chain=list(self.body_chain),
        )
