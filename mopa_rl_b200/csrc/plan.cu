// RRT-Connect over the active joints (state space + planner), sm_100a.
#include <stdexcept>

#include "planner_state.h"

namespace mopa {

void build_space(const mopa_model_desc *d, const int32_t *passive, int n_passive, double range, double resolution,
                 uint64_t seed, Space &out) {
    out = Space();
    out.nq = d->nq;
    out.range = (float)range;
    out.resolution = (float)resolution;
    out.seed = seed;
    for (int j = 0; j < d->njnt; j++) {
        int adr = d->jnt_qposadr[j];
        bool is_passive = false;
        for (int k = 0; k < n_passive; k++)
            if (passive[k] == adr) is_passive = true;
        if (is_passive) continue;
        int t = d->jnt_type[j];
        if (t == MOPA_JNT_FREE || t == MOPA_JNT_BALL)
            throw std::runtime_error("free/ball joints cannot be active planner joints (pass their qpos indices as passive)");
        out.active_qadr.push_back(adr);
        if (t == MOPA_JNT_HINGE && !d->jnt_limited[j]) {
            out.is_so2.push_back(1);
            out.lo.push_back(-3.14159265358979323846f);
            out.hi.push_back(3.14159265358979323846f);
        } else {
            out.is_so2.push_back(0);
            out.lo.push_back((float)d->jnt_range[2 * j]);
            out.hi.push_back((float)d->jnt_range[2 * j + 1]);
        }
    }
    out.n_active = (int)out.active_qadr.size();
    // the reference throws when the joint dimensions do not add up (mujoco_ompl_interface.cpp:268-272)
    if (out.n_active != d->nq - n_passive) throw std::runtime_error("Total joint dimensions are not equal to nq - size(passive_joints)");
}

void free_plan_buffers(mopa_planner *p) { (void)p; }

}  // namespace mopa
