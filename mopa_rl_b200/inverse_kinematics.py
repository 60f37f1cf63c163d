"""Batched inverse kinematics on the GPU behind the reference's ``qpos_from_site_pose`` semantics
(env/inverse_kinematics.py:18-135; used by the IK-target presets through
MoPARolloutRunner._cart2dispalcement, rl/mopa_rollouts.py:683-728).

``qpos_from_site_pose_batch(venv, site, qpos, target_pos, target_quat=None, max_steps=100, tol=1e-2)`` solves one
problem per row with the kernel ``mopa_ik_batch`` and returns an ``IKResult`` of tensors (qpos, err_norm, steps,
success).  The movable joints are the arm joints (``env.robot_joints``); everything else keeps its value.
"""
from __future__ import annotations

import collections
import ctypes as C

import numpy as np

from .capi import check, lib

IKResult = collections.namedtuple("IKResult", ["qpos", "err_norm", "steps", "success"])


def site_frame(model, dyn, site):
    """(simulated-body index, local position) of a named site."""
    sid = model.site_name2id(site)
    sim_body = {b: i for i, b in enumerate(dyn.bodies)}
    return sim_body[int(model.site_bodyid[sid])], np.ascontiguousarray(model.site_pos[sid], dtype=np.float64)


def qpos_from_site_pose_batch(venv, site, qpos, target_pos, target_quat=None, max_steps=100, tol=1e-2):
    torch = venv.torch
    L = lib()
    L.mopa_ik_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    body, local = site_frame(venv.model, venv.dyn, site)
    dofs = np.ascontiguousarray([int(venv.task.arm_dof[k]) for k in range(7)], dtype=np.int32)
    q = qpos.to(device=venv.dev, dtype=torch.float64).contiguous()
    n = q.shape[0]
    tp = target_pos.to(device=venv.dev, dtype=torch.float64).contiguous()
    tq = target_quat.to(device=venv.dev, dtype=torch.float64).contiguous() if target_quat is not None else None
    out = torch.empty_like(q)
    err = torch.empty(n, dtype=torch.float64, device=venv.dev)
    steps = torch.empty(n, dtype=torch.int32, device=venv.dev)
    ok = torch.empty(n, dtype=torch.uint8, device=venv.dev)
    check(L.mopa_ik_batch(venv.h, q.data_ptr(), tp.data_ptr(), tq.data_ptr() if tq is not None else None, int(body), local.ctypes.data,
                          dofs.ctypes.data, 7, n, int(max_steps), float(tol), out.data_ptr(), err.data_ptr(), steps.data_ptr(), ok.data_ptr(),
                          C.c_void_p(torch.cuda.current_stream(venv.dev).cuda_stream)))
    torch.cuda.current_stream(venv.dev).synchronize()   # the input staging tensors must outlive the launch
    return IKResult(qpos=out, err_norm=err, steps=steps, success=ok.bool())
