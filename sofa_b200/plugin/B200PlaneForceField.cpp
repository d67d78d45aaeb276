// SOFA-side glue: PlaneForceField<B200Vec3Types>::addForce / addDForce -> sofab200_plane_add_force / sofab200_plane_add_dforce (it is in every SofaCUDA
// FEM benchmark scene).  The contact list m_contacts of the reference (PlaneForceField.h:65) becomes a per-node flag vector on the device, kept in the
// PlaneForceFieldInternalData member the class reserves (PlaneForceField.h:37-40,67).
#include <sofa/component/mechanicalload/PlaneForceField.inl>
#include <sofa/core/ObjectFactory.h>

#include "B200Types.h"

namespace sofa::component::mechanicalload {
using sofa::b200::B200Vec3Types;

template <class TReal> class PlaneForceFieldInternalData<B200Vec3Types<TReal>> {
public:
    sofa::b200::B200Vector<unsigned char> contacts;     // 1 = the node was in contact at the last addForce
};

namespace {
template <class FF> sofab200_plane_desc b200_plane(const FF* ff) {
    sofab200_plane_desc p{};
    const auto n = ff->d_planeNormal.getValue();
    p.normal[0] = n[0]; p.normal[1] = n[1]; p.normal[2] = n[2];
    p.d = ff->d_planeD.getValue(); p.stiffness = ff->d_stiffness.getValue(); p.damping = ff->d_damping.getValue();
    p.max_force = ff->d_maxForce.getValue(); p.bilateral = ff->d_bilateral.getValue() ? 1 : 0;
    return p;
}
}  // namespace

#define B200_PLANE(TReal)                                                                                                            \
    template <> void PlaneForceField<B200Vec3Types<TReal>>::addForce(const core::MechanicalParams*, DataVecDeriv& f, const DataVecCoord& x, const DataVecDeriv& v) { \
        auto& ff = *f.beginEdit();      /* PlaneForceField.inl:139-205 (d_localRange is not supported on the device) */              \
        const auto& xx = x.getValue();                                                                                               \
        ff.resize(xx.size());                                                                                                        \
        m_data.contacts.resize(xx.size());                                                                                           \
        const sofab200_plane_desc p = b200_plane(this);                                                                              \
        if (sofab200_plane_add_force(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, xx.size(), sofa::b200::devWrite(ff), sofa::b200::devRead(xx), \
                                     sofa::b200::devRead(v.getValue()), &p, static_cast<unsigned char*>(sofa::b200::devWrite(m_data.contacts))) != SOFAB200_OK) \
            msg_error() << sofab200_last_error();                                                                                    \
        f.endEdit();                                                                                                                 \
    }                                                                                                                                \
    template <> void PlaneForceField<B200Vec3Types<TReal>>::addDForce(const core::MechanicalParams* mparams, DataVecDeriv& df, const DataVecDeriv& dx) { \
        auto& dff = *df.beginEdit();    /* PlaneForceField.inl:208-226 */                                                            \
        const auto& dxx = dx.getValue();                                                                                             \
        dff.resize(dxx.size());                                                                                                      \
        const sofab200_plane_desc p = b200_plane(this);                                                                              \
        const double k = sofa::core::mechanicalparams::kFactorIncludingRayleighDamping(mparams, this->rayleighStiffness.getValue()); \
        if (m_data.contacts.size() == dxx.size() &&                                                                                  \
            sofab200_plane_add_dforce(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, dxx.size(), sofa::b200::devWrite(dff), sofa::b200::devRead(dxx), &p, \
                                      static_cast<const unsigned char*>(sofa::b200::devRead(m_data.contacts)), k) != SOFAB200_OK)    \
            msg_error() << sofab200_last_error();                                                                                    \
        df.endEdit();                                                                                                                \
    }
B200_PLANE(float)
B200_PLANE(double)
template class PlaneForceField<sofa::b200::B200Vec3fTypes>;
template class PlaneForceField<sofa::b200::B200Vec3dTypes>;
}  // namespace sofa::component::mechanicalload

namespace sofa::b200 {
void registerPlaneForceField(sofa::core::ObjectFactory* factory) {
    using namespace sofa::component::mechanicalload;
    factory->registerObjects(sofa::core::ObjectRegistrationData("PlaneForceField on a B200 GPU (sofa_b200)").add<PlaneForceField<B200Vec3fTypes>>().add<PlaneForceField<B200Vec3dTypes>>());
}
}  // namespace sofa::b200
