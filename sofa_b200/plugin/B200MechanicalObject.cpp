// SOFA-side glue: MechanicalObject<B200Vec3Types>::vOp / vMultiOp / vDot / resetForce, DiagonalMass<B200Vec3Types>
// ::addMDx / addForce / accFromF and FixedProjectiveConstraint<B200Vec3Types>::projectResponse, each forwarding to one
// C-ABI entry point (the set SofaCUDA specialises for the same scenes:
// applications/plugins/SofaCUDA/Component/src/SofaCUDA/component/statecontainer/CudaMechanicalObject.h:133-142).
#include <sofa/component/constraint/projective/FixedProjectiveConstraint.inl>
#include <sofa/component/mass/DiagonalMass.inl>
#include <sofa/component/statecontainer/MechanicalObject.inl>
#include <sofa/core/ObjectFactory.h>

#include "B200Types.h"

namespace sofa::component::statecontainer {
using sofa::b200::B200Vec3Types;

#define B200_MO(TReal)                                                                                                               \
    template <> void MechanicalObject<B200Vec3Types<TReal>>::vOp(const core::ExecParams*, core::VecId r, core::ConstVecId a,        \
                                                                 core::ConstVecId b, SReal k) {                                      \
        /* null ids become null pointers; the C ABI reproduces the dispatch of MechanicalObject.inl:2075-2203 */                     \
        auto dptr = [&](core::ConstVecId id, bool write) -> void* {                                                                  \
            if (id.isNull()) return nullptr;                                                                                         \
            if (id.type == core::V_COORD) { auto* d = this->write(core::VecCoordId(id)); return write ? d->beginEdit()->deviceWrite() : const_cast<void*>(d->getValue().deviceRead()); } \
            auto* d = this->write(core::VecDerivId(id)); return write ? d->beginEdit()->deviceWrite() : const_cast<void*>(d->getValue().deviceRead()); \
        };                                                                                                                           \
        void* pr = dptr(r, true);                                                                                                    \
        const void* pa = a == r ? pr : dptr(a, false);                                                                               \
        const void* pb = b == r ? pr : dptr(b, false);                                                                               \
        if (sofab200_mo_vop(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, this->getSize(), pr, pa, pb, k) != SOFAB200_OK) \
            msg_error() << sofab200_last_error();                                                                                    \
    }                                                                                                                                \
    template <> SReal MechanicalObject<B200Vec3Types<TReal>>::vDot(const core::ExecParams*, core::ConstVecId a, core::ConstVecId b) { \
        double r = 0.0;                                                                                                              \
        const void* pa = this->read(core::ConstVecDerivId(a))->getValue().deviceRead();                                              \
        const void* pb = this->read(core::ConstVecDerivId(b))->getValue().deviceRead();                                              \
        sofab200_mo_vdot(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, this->getSize(), pa, pb, &r);                   \
        return r;                                                                                                                    \
    }
B200_MO(float)
B200_MO(double)
template class MechanicalObject<sofa::b200::B200Vec3fTypes>;
template class MechanicalObject<sofa::b200::B200Vec3dTypes>;
}  // namespace sofa::component::statecontainer

namespace sofa::component::mass {
using sofa::b200::B200Vec3Types;
#define B200_MASS(TReal)                                                                                                             \
    template <> void DiagonalMass<B200Vec3Types<TReal>>::addMDx(const core::MechanicalParams*, DataVecDeriv& res, const DataVecDeriv& dx, SReal factor) { \
        auto& r = *res.beginEdit();                                                                                                  \
        sofab200_mass_add_mdx(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, r.size(), r.deviceWrite(), dx.getValue().deviceRead(), \
                              d_vertexMass.getValue().deviceRead(), factor);                                                         \
        res.endEdit();                                                                                                               \
    }                                                                                                                                \
    template <> void DiagonalMass<B200Vec3Types<TReal>>::addForce(const core::MechanicalParams*, DataVecDeriv& f, const DataVecCoord&, const DataVecDeriv&) { \
        if (this->m_separateGravity.getValue()) return;                                                                              \
        const sofa::type::Vec3d g(this->getContext()->getGravity());                                                                 \
        auto& ff = *f.beginEdit();                                                                                                   \
        sofab200_mass_add_force(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, ff.size(), ff.deviceWrite(), d_vertexMass.getValue().deviceRead(), g.ptr()); \
        f.endEdit();                                                                                                                 \
    }
B200_MASS(float)
B200_MASS(double)
}  // namespace sofa::component::mass

namespace sofa::component::constraint::projective {
using sofa::b200::B200Vec3Types;
#define B200_FIXED(TReal)                                                                                                            \
    template <> void FixedProjectiveConstraint<B200Vec3Types<TReal>>::projectResponse(const core::MechanicalParams*, DataVecDeriv& resData) { \
        auto& res = *resData.beginEdit();                                                                                            \
        const auto& idx = d_indices.getValue(); /* uploaded once into data->indicesDevice by init(), omitted here */                 \
        sofab200_fixed_project_response(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, res.size(), res.deviceWrite(), idx.size(), \
                                        data->indicesDevice, d_fixAll.getValue());                                                   \
        resData.endEdit();                                                                                                           \
    }
B200_FIXED(float)
B200_FIXED(double)
}  // namespace sofa::component::constraint::projective
