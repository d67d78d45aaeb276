// SOFA-side glue: MechanicalObject<B200Vec3Types>::vOp / vMultiOp / vDot / resetForce / accumulateForce, each forwarding to one C-ABI entry point (the set SofaCUDA
// specialises for the same scenes: applications/plugins/SofaCUDA/Component/src/SofaCUDA/component/statecontainer/CudaMechanicalObject.h:133-142).
#include <sofa/component/statecontainer/MechanicalObject.inl>
#include <sofa/core/ObjectFactory.h>

#include "B200Types.h"

namespace sofa::component::statecontainer {
using sofa::b200::B200Vec3Types;

namespace {
/// device pointer of the state vector `id` of `mo` (null id -> null pointer), for reading or for writing
template <class MO> void* b200_vec(MO* mo, core::ConstVecId id, bool write) {
    if (id.isNull()) return nullptr;
    if (id.type == core::V_COORD) {
        auto* d = mo->write(core::VecCoordId(id.index));
        if (write) { void* p = sofa::b200::devWrite(*d->beginEdit()); d->endEdit(); return p; }
        return const_cast<void*>(sofa::b200::devRead(d->getValue()));
    }
    auto* d = mo->write(core::VecDerivId(id.index));
    if (write) { void* p = sofa::b200::devWrite(*d->beginEdit()); d->endEdit(); return p; }
    return const_cast<void*>(sofa::b200::devRead(d->getValue()));
}
}  // namespace

#define B200_MO(TReal)                                                                                                               \
    template <> void MechanicalObject<B200Vec3Types<TReal>>::vOp(const core::ExecParams*, core::VecId r, core::ConstVecId a,        \
                                                                 core::ConstVecId b, SReal k) {                                      \
        /* null ids become null pointers; the C ABI reproduces the dispatch of MechanicalObject.inl:2075-2203 */                     \
        if (r.isNull()) { msg_error() << "Invalid vOp operation: the result vector is null"; return; }                               \
        void* pr = b200_vec(this, r, true);                                                                                          \
        const void* pa = a == core::ConstVecId(r) ? pr : b200_vec(this, a, false);                                                   \
        const void* pb = b == core::ConstVecId(r) ? pr : b200_vec(this, b, false);                                                   \
        if (sofab200_mo_vop(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, this->getSize(), pr, pa, pb, k) != SOFAB200_OK) \
            msg_error() << sofab200_last_error();                                                                                    \
    }                                                                                                                                \
    template <> SReal MechanicalObject<B200Vec3Types<TReal>>::vDot(const core::ExecParams*, core::ConstVecId a, core::ConstVecId b) { \
        double r = 0.0;                                                                                                              \
        if (sofab200_mo_vdot(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, this->getSize(), b200_vec(this, a, false),  \
                             b200_vec(this, b, false), &r) != SOFAB200_OK)                                                           \
            msg_error() << sofab200_last_error();                                                                                    \
        return r;                                                                                                                    \
    }                                                                                                                                \
    template <> void MechanicalObject<B200Vec3Types<TReal>>::vMultiOp(const core::ExecParams* params, const VMultiOp& ops) {        \
        /* the integration of EulerImplicitSolver (EulerImplicitSolver.cpp:259-284): v += a*f ; x += v*h -> ONE kernel; anything else   \
           takes the reference's own fallback, a sequence of vOp (BaseMechanicalState.cpp:42-79) */                                  \
        if (ops.size() == 2 && ops[0].second.size() == 2 && ops[1].second.size() == 2) {                                             \
            const core::VecId v = ops[0].first.getId(this), x = ops[1].first.getId(this);                                            \
            const core::ConstVecId v0 = ops[0].second[0].first.getId(this), a = ops[0].second[1].first.getId(this);                  \
            const core::ConstVecId x0 = ops[1].second[0].first.getId(this), v1 = ops[1].second[1].first.getId(this);                 \
            if (v0 == core::ConstVecId(v) && x0 == core::ConstVecId(x) && v1 == core::ConstVecId(v) && ops[0].second[0].second == 1.0 && \
                ops[1].second[0].second == 1.0 && v.type == core::V_DERIV && x.type == core::V_COORD && a.type == core::V_DERIV) {   \
                if (sofab200_mo_vmultiop_integrate(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, this->getSize(), b200_vec(this, v, true), \
                                                   b200_vec(this, x, true), b200_vec(this, a, false), ops[0].second[1].second,      \
                                                   ops[1].second[1].second) != SOFAB200_OK)                                          \
                    msg_error() << sofab200_last_error();                                                                            \
                return;                                                                                                              \
            }                                                                                                                        \
        }                                                                                                                            \
        core::behavior::BaseMechanicalState::vMultiOp(params, ops);                                                                  \
    }                                                                                                                                \
    template <> void MechanicalObject<B200Vec3Types<TReal>>::resetForce(const core::ExecParams*, core::VecDerivId f) {              \
        /* resetDataTypeVec of the force vector (MechanicalObject.inl:2497-2505) = vOp(f) with null operands */                      \
        if (sofab200_mo_vop(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, this->getSize(), b200_vec(this, core::ConstVecId(f), true), \
                            nullptr, nullptr, 0.0) != SOFAB200_OK)                                                                   \
            msg_error() << sofab200_last_error();                                                                                    \
    }                                                                                                                                \
    template <> void MechanicalObject<B200Vec3Types<TReal>>::accumulateForce(const core::ExecParams*, core::VecDerivId f) {         \
        /* MechanicalObject.inl:1356-1375: f += externalForce for the rows that differ from Deriv(), only when the Data is not empty */ \
        const VecDeriv& ext = this->read(core::vec_id::read_access::externalForce)->getValue();                                      \
        if (ext.empty()) return;                                                                                                     \
        if (sofab200_mo_accumulate_force(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, ext.size(), b200_vec(this, core::ConstVecId(f), true), \
                                         ext.deviceRead()) != SOFAB200_OK)                                                           \
            msg_error() << sofab200_last_error();                                                                                    \
    }
B200_MO(float)
B200_MO(double)
template class MechanicalObject<sofa::b200::B200Vec3fTypes>;
template class MechanicalObject<sofa::b200::B200Vec3dTypes>;
}  // namespace sofa::component::statecontainer

namespace sofa::b200 {
void registerMechanicalObject(sofa::core::ObjectFactory* factory) {
    using namespace sofa::component::statecontainer;
    factory->registerObjects(sofa::core::ObjectRegistrationData("MechanicalObject whose state vectors live in B200 HBM (sofa_b200)")
                                 .add<MechanicalObject<B200Vec3fTypes>>()
                                 .add<MechanicalObject<B200Vec3dTypes>>());
}
}  // namespace sofa::b200
