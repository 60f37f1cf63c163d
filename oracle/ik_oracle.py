"""TEST INFRASTRUCTURE — numpy restatement of qpos_from_site_pose (env/inverse_kinematics.py:18-135) and
nullspace_method (:295-303) on top of the kinematic chain of the dynamics description (mopa_dyn_desc).
mju_mat2Quat / mju_negQuat / mju_mulQuat / mju_quat2Vel (dm_control mjlib, absent here) are restated from the
MuJoCo documentation."""
from __future__ import annotations

import numpy as np


def _q2m(q):
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def _qmul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def _mat2quat(R):
    m = R.ravel()
    q = np.zeros(4)
    if m[0] + m[4] + m[8] > 0:
        q[0] = 0.5 * np.sqrt(1 + m[0] + m[4] + m[8])
        q[1], q[2], q[3] = 0.25 * (m[7] - m[5]) / q[0], 0.25 * (m[2] - m[6]) / q[0], 0.25 * (m[3] - m[1]) / q[0]
    elif m[0] > m[4] and m[0] > m[8]:
        q[1] = 0.5 * np.sqrt(1 + m[0] - m[4] - m[8])
        q[0], q[2], q[3] = 0.25 * (m[7] - m[5]) / q[1], 0.25 * (m[1] + m[3]) / q[1], 0.25 * (m[2] + m[6]) / q[1]
    elif m[4] > m[8]:
        q[2] = 0.5 * np.sqrt(1 - m[0] + m[4] - m[8])
        q[0], q[1], q[3] = 0.25 * (m[2] - m[6]) / q[2], 0.25 * (m[1] + m[3]) / q[2], 0.25 * (m[5] + m[7]) / q[2]
    else:
        q[3] = 0.5 * np.sqrt(1 - m[0] - m[4] + m[8])
        q[0], q[1], q[2] = 0.25 * (m[3] - m[1]) / q[3], 0.25 * (m[2] + m[6]) / q[3], 0.25 * (m[5] + m[7]) / q[3]
    return q / np.linalg.norm(q)


def _quat2vel(q):
    ax = q[1:].copy()
    s = np.linalg.norm(ax)
    if s > 0:
        ax /= s
    speed = 2 * np.arctan2(s, q[0])
    if speed > np.pi:
        speed -= 2 * np.pi
    return ax * speed


class IKOracle:
    def __init__(self, dynmodel, body, site_local, joint_dofs):
        self.A, self.body, self.site, self.jdof = dynmodel._arr, int(body), np.asarray(site_local, np.float64), [int(j) for j in joint_dofs]
        chain, b = [], self.body
        while b >= 0:
            chain.append(b)
            b = int(self.A["b_parent"][b])
        self.chain = chain[::-1]

    def site_pose(self, q):
        """site position, rotation and the (axis, anchor, type) of every movable joint on the chain."""
        A = self.A
        pos, quat, joints = None, None, {}
        for c, i in enumerate(self.chain):
            Pp, Pq = (A["b_rootpos"][i], A["b_rootquat"][i]) if c == 0 else (pos, quat)
            jt = int(A["b_jtype"][i])
            if jt == 0:
                a = int(A["b_qadr"][i])
                pos, quat = q[a:a + 3].copy(), q[a + 3:a + 7] / np.linalg.norm(q[a + 3:a + 7])
                continue
            pos = Pp + _q2m(Pq) @ A["b_pos"][i]
            quat = _qmul(Pq, A["b_quat"][i])
            if jt in (2, 3):
                anchor = pos + _q2m(quat) @ A["b_jpos"][i]
                if jt == 3:
                    ang = q[int(A["b_qadr"][i])] - A["b_qpos0"][i]
                    quat = _qmul(quat, np.concatenate([[np.cos(0.5 * ang)], np.sin(0.5 * ang) * A["b_jaxis"][i]]))
                    R = _q2m(quat)
                    pos = anchor - R @ A["b_jpos"][i]
                    ax = R @ A["b_jaxis"][i]
                else:
                    ax = _q2m(quat) @ A["b_jaxis"][i]
                    pos = pos + ax * (q[int(A["b_qadr"][i])] - A["b_qpos0"][i])
                joints[int(A["b_dadr"][i])] = (ax, anchor, jt)
        R = _q2m(quat)
        return pos + R @ self.site, R, joints

    def solve(self, qpos, target_pos, target_quat=None, max_steps=100, rot_weight=1.0, tol=1e-14, max_update_norm=2.0,
              progress_thresh=20.0, regularization_strength=3e-2):
        A = self.A
        q = np.array(qpos, np.float64)
        steps, success, err_norm = 0, False, 0.0
        for steps in range(max_steps):
            sp, R, joints = self.site_pose(q)
            err = target_pos - sp
            err_norm = np.linalg.norm(err)
            if target_quat is not None:
                sq = _mat2quat(R)
                erot = _quat2vel(_qmul(np.asarray(target_quat, np.float64), np.array([sq[0], -sq[1], -sq[2], -sq[3]])))
                err_norm += np.linalg.norm(erot) * rot_weight
                err = np.concatenate([err, erot])
            if err_norm < tol:
                success = True
                break
            J = np.zeros((len(err), len(self.jdof)))
            for c, d in enumerate(self.jdof):
                if d in joints:
                    ax, anchor, jt = joints[d]
                    J[:3, c] = np.cross(ax, sp - anchor) if jt == 3 else ax
                    if target_quat is not None and jt == 3:
                        J[3:, c] = ax
            H = J.T @ J + np.eye(len(self.jdof)) * regularization_strength    # nullspace_method, regularised branch
            upd = np.linalg.solve(H, J.T @ err)
            un = np.linalg.norm(upd)
            if err_norm / un > progress_thresh:
                break
            if un > max_update_norm:
                upd = upd * (max_update_norm / un)
            for c, d in enumerate(self.jdof):
                q[int(A["d_qadr"][d])] += upd[c]
        return q, err_norm, steps, success
