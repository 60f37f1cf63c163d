"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares; ctypes
struct layouts match the C headers; the product never imports the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "mopa_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mopa_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mopa_rl_b200 import capi

    L = capi.lib()
    names = _declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), "libmopa_b200.so does not export %s" % n
    assert L.mopa_last_error() is not None


def test_no_gpu_means_a_loud_error_not_a_fallback(push_model):
    from mopa_rl_b200 import capi

    if capi.lib().mopa_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(capi.MopaError) as e:
        capi.NativePlanner(push_model, list(range(7, push_model.nq)), [], -0.002, 0.1)
    assert "CUDA" in str(e.value) or "cuda" in str(e.value)
    # model errors are reported as such (the reference throws from makeCompoundStateSpace, mujoco_ompl_interface.cpp:268-272)
    with pytest.raises(capi.MopaError) as e:
        capi.NativePlanner(push_model, [7, 8], [], -0.002, 0.1)
    assert "passive" in str(e.value)


def test_struct_layouts_match_the_headers(tmp_path):
    """Compile a tiny C program against include/*.h and compare sizeof/offsetof with the ctypes mirrors."""
    from mopa_rl_b200.dynmodel import DynDesc
    from mopa_rl_b200.envs import EnvBuffers, SawyerTask
    from mopa_rl_b200.model import ModelDesc
    from mopa_rl_b200.rollout import _RolloutConfig

    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "mopa_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu\n", sizeof(mopa_model_desc), sizeof(mopa_dyn_desc), sizeof(mopa_sawyer_task), sizeof(mopa_env_buffers));
  printf("%zu %zu %zu %zu\n", offsetof(mopa_model_desc, site_quat), offsetof(mopa_dyn_desc, p_g2), offsetof(mopa_sawyer_task, success_reward), offsetof(mopa_env_buffers, ncon));
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(mopa_rollout_config), offsetof(mopa_rollout_config, qpos0), offsetof(mopa_rollout_config, seed_reuse),
         offsetof(mopa_rollout_config, ac_space_normal), offsetof(mopa_sawyer_task, geom_cube), offsetof(mopa_sawyer_task, bin_z),
         offsetof(mopa_model_desc, mesh_vert));
  printf("%zu %zu %zu %zu\n", offsetof(mopa_sawyer_task, unstable_penalty), offsetof(mopa_env_buffers, grasp), offsetof(mopa_env_buffers, unstable),
         offsetof(mopa_rollout_config, debug_block_mod));
  return 0; }'''
    src = tmp_path / "layout.c"
    src.write_text(prog)
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)], text=True).split()
    sizes, offs, more, r2 = list(map(int, out[:4])), list(map(int, out[4:8])), list(map(int, out[8:15])), list(map(int, out[15:]))
    assert sizes == [C.sizeof(ModelDesc), C.sizeof(DynDesc), C.sizeof(SawyerTask), C.sizeof(EnvBuffers)]
    assert offs == [ModelDesc.site_quat.offset, DynDesc.p_g2.offset, SawyerTask.success_reward.offset, EnvBuffers.ncon.offset]
    # the rollout configuration and the fields added for the lift task / mesh collider (a mismatch here would only show on a GPU)
    assert more == [C.sizeof(_RolloutConfig), _RolloutConfig.qpos0.offset, _RolloutConfig.seed_reuse.offset, _RolloutConfig.ac_space_normal.offset,
                    SawyerTask.geom_cube.offset, SawyerTask.bin_z.offset, ModelDesc.mesh_vert.offset]
    # round 2: instability guard, persistent grasp flags, fallback planners
    assert r2 == [SawyerTask.unstable_penalty.offset, EnvBuffers.grasp.offset, EnvBuffers.unstable.offset, _RolloutConfig.debug_block_mod.offset]
    # the library reports the sizes it was compiled with; the binding refuses a stale build (capi._check_abi)
    from mopa_rl_b200 import capi

    got = (C.c_int32 * 5)()
    assert capi.lib().mopa_abi_sizes(got) == 0
    assert list(got) == sizes + [C.sizeof(_RolloutConfig)]


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mopa_rl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("the oracle", "").replace("CPU oracle", "").replace("with the oracle", "") or \
                    not re.search(r"(^|\n)\s*(from|import)\s+oracle|#include\s+[\"<].*oracle", txt), f


def test_rng_and_reset_distribution(push_model):
    from mopa_rl_b200 import rng
    from mopa_rl_b200.envs import PUSH_INIT_QPOS, push_reset_state

    u = rng.uniform01(1, np.arange(1000, dtype=np.uint64)[:, None], np.uint64(3), np.arange(8, dtype=np.uint64)[None, :])
    assert u.shape == (1000, 8) and 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.02
    assert np.array_equal(u[5], rng.uniform01(1, np.uint64(5), np.uint64(3), np.arange(8, dtype=np.uint64)))   # batch independent
    z = rng.normal(2, np.arange(20000, dtype=np.uint64), np.uint64(0), np.uint64(0))
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1) < 0.03
    q, v = push_reset_state(push_model, 9, np.arange(2000), np.zeros(2000, dtype=np.int64))
    assert q.shape == (2000, 36) and (v == 0).all()
    d = q[:, :7] - PUSH_INIT_QPOS
    assert abs(d.std() - 0.02) < 0.002 and abs(d.mean()) < 0.002               # arm = init_qpos + N(0, 0.02^2)
    t = q[:, 34:36]
    assert t.min() >= -0.01 and t.max() <= 0.01 and t.std() > 0.004              # target += U(-0.01, 0.01)
    assert np.array_equal(q[:, 7:34], np.tile(push_model.qpos0[7:34], (2000, 1)))
    q2, _ = push_reset_state(push_model, 9, [7], [1])
    assert not np.array_equal(q2[0], q[7])                                       # next episode, new draw
