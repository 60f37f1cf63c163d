"""GPU <-> oracle parity of the batched inverse kinematics (env/inverse_kinematics.py:18-135), through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("with_quat", [False, True], ids=["position", "pose"])
def test_ik_batch_matches_oracle(push_model, with_quat):
    import torch

    from helpers import PUSH_INIT_QPOS
    from mopa_rl_b200.envs import VecSawyerPushObstacle
    from mopa_rl_b200.inverse_kinematics import qpos_from_site_pose_batch, site_frame
    from oracle.ik_oracle import IKOracle, _mat2quat

    n = 64
    venv = VecSawyerPushObstacle(n, seed=5)
    body, local = site_frame(push_model, venv.dyn, "grip_site")
    dofs = [int(venv.task.arm_dof[k]) for k in range(7)]
    orc = IKOracle(venv.dyn, body, local, dofs)
    rng = np.random.default_rng(3)
    q0 = np.tile(push_model.qpos0, (n, 1))
    q0[:, :7] = PUSH_INIT_QPOS + rng.normal(0, 0.02, (n, 7))
    # targets: the site pose of a nearby configuration (reachable), every fourth one far away (progress criterion / step cap)
    tp, tq = np.zeros((n, 3)), np.zeros((n, 4))
    for i in range(n):
        qt = q0[i].copy()
        qt[:7] += rng.uniform(-0.4, 0.4, 7)
        sp, R, _ = orc.site_pose(qt)
        tp[i], tq[i] = sp, _mat2quat(R)
        if i % 4 == 3:
            tp[i] += rng.uniform(-1.5, 1.5, 3)
    res = qpos_from_site_pose_batch(venv, "grip_site", torch.as_tensor(q0), torch.as_tensor(tp), torch.as_tensor(tq) if with_quat else None,
                                    max_steps=100, tol=1e-2)
    gq, gerr, gsteps, gok = res.qpos.cpu().numpy(), res.err_norm.cpu().numpy(), res.steps.cpu().numpy(), res.success.cpu().numpy()
    n_ok = 0
    for i in range(n):
        q, err, steps, ok = orc.solve(q0[i], tp[i], tq[i] if with_quat else None, max_steps=100, tol=1e-2)
        assert steps == gsteps[i] and ok == bool(gok[i]), (i, steps, gsteps[i], ok, gok[i])
        # converged problems agree to rounding; the 100-step wander towards an unreachable target amplifies rounding differences
        tq_ = 1e-9 if ok else 1e-4
        assert np.abs(q - gq[i]).max() < tq_ and abs(err - gerr[i]) < tq_, (i, ok, np.abs(q - gq[i]).max())
        assert np.array_equal(q[7:], q0[i][7:])          # only the arm joints move
        n_ok += ok
    assert 0 < n_ok < n                                   # both outcomes are exercised
    print("IK (%s): %d / %d converged, mean steps %.1f" % ("pose" if with_quat else "position", n_ok, n, gsteps.mean()))
