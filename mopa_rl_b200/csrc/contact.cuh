// Contact generation for the physics step (placeholder until the contact routines land).
#pragma once
#include "dyn.cuh"

namespace mopa {
DYN_HD inline int contact_rows(const DynDev &m, const DynData &D, const Sv6 *S, CRow *rows, int maxrows) {
    (void)m; (void)D; (void)S; (void)rows; (void)maxrows;
    return 0;
}
}  // namespace mopa
