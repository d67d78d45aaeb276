// SOFA-side glue: DiagonalMass<B200Vec3Types>::addMDx / addForce / accFromF and UniformMass<B200Vec3Types>::addMDx / addForce, each forwarding to
// one C-ABI entry point.  The reference keeps vertexMass in a host-side topology container (PointData<vector<MassType>>, DiagonalMass.h:46-47,81);
// the device copy the kernels read lives in a file-local table keyed by the component and is refreshed whenever the Data's counter moves.
#include <mutex>
#include <unordered_map>

#include <sofa/component/mass/DiagonalMass.inl>
#include <sofa/component/mass/UniformMass.inl>
#include <sofa/core/ObjectFactory.h>

#include "B200Types.h"

namespace sofa::component::mass {
using sofa::b200::B200Vec3Types;

namespace {
template <class TReal> struct DeviceMass { sofa::b200::B200Vector<TReal> m; int counter = -1; };
std::mutex g_mutex;
template <class TReal> std::unordered_map<const void*, DeviceMass<TReal>>& table() { static std::unordered_map<const void*, DeviceMass<TReal>> t; return t; }
/// device pointer of the component's vertexMass (uploaded when the Data changed since the last call)
template <class TReal, class Mass> const void* deviceVertexMass(const Mass* self) {
    std::lock_guard<std::mutex> l(g_mutex);
    DeviceMass<TReal>& d = table<TReal>()[self];
    if (d.counter != self->d_vertexMass.getCounter()) {
        const auto& m = self->d_vertexMass.getValue();
        d.m.resize(m.size());
        std::copy(m.begin(), m.end(), d.m.hostWrite());
        d.counter = self->d_vertexMass.getCounter();
    }
    return sofa::b200::devRead(d.m);
}
}  // namespace

#define B200_DIAGONAL_MASS(TReal)                                                                                                    \
    template <> void DiagonalMass<B200Vec3Types<TReal>>::addMDx(const core::MechanicalParams*, DataVecDeriv& res, const DataVecDeriv& dx, SReal factor) { \
        auto& r = *res.beginEdit();     /* DiagonalMass.inl:535-575 */                                                               \
        if (sofab200_mass_add_mdx(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, r.size(), sofa::b200::devWrite(r), sofa::b200::devRead(dx.getValue()), \
                                  deviceVertexMass<TReal>(this), factor) != SOFAB200_OK)                                             \
            msg_error() << sofab200_last_error();                                                                                    \
        res.endEdit();                                                                                                               \
    }                                                                                                                                \
    template <> void DiagonalMass<B200Vec3Types<TReal>>::addForce(const core::MechanicalParams*, DataVecDeriv& f, const DataVecCoord&, const DataVecDeriv&) { \
        if (this->m_separateGravity.getValue()) return;      /* DiagonalMass.inl:1392-1413 */                                        \
        const sofa::type::Vec3d g(this->getContext()->getGravity());                                                                 \
        auto& ff = *f.beginEdit();                                                                                                   \
        if (sofab200_mass_add_force(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, ff.size(), sofa::b200::devWrite(ff), \
                                    deviceVertexMass<TReal>(this), g.ptr()) != SOFAB200_OK)                                          \
            msg_error() << sofab200_last_error();                                                                                    \
        f.endEdit();                                                                                                                 \
    }                                                                                                                                \
    template <> void DiagonalMass<B200Vec3Types<TReal>>::accFromF(const core::MechanicalParams*, DataVecDeriv& a, const DataVecDeriv& f) { \
        auto& aa = *a.beginEdit();      /* DiagonalMass.inl:577-590 */                                                               \
        if (sofab200_mass_acc_from_f(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, aa.size(), sofa::b200::devWrite(aa), \
                                     sofa::b200::devRead(f.getValue()), deviceVertexMass<TReal>(this)) != SOFAB200_OK)               \
            msg_error() << sofab200_last_error();                                                                                    \
        a.endEdit();                                                                                                                 \
    }                                                                                                                                \
    template <> void UniformMass<B200Vec3Types<TReal>>::addMDx(const core::MechanicalParams*, DataVecDeriv& res, const DataVecDeriv& dx, SReal factor) { \
        if (!this->isComponentStateValid()) return;          /* UniformMass.inl:403-420 (d_localRange is not supported on the device) */ \
        auto& r = *res.beginEdit();                                                                                                  \
        if (sofab200_uniform_mass_add_mdx(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, r.size(), sofa::b200::devWrite(r), \
                                          sofa::b200::devRead(dx.getValue()), double(d_vertexMass.getValue()), factor) != SOFAB200_OK) \
            msg_error() << sofab200_last_error();                                                                                    \
        res.endEdit();                                                                                                               \
    }                                                                                                                                \
    template <> void UniformMass<B200Vec3Types<TReal>>::addForce(const core::MechanicalParams*, DataVecDeriv& f, const DataVecCoord&, const DataVecDeriv&) { \
        if (this->m_separateGravity.getValue()) return;      /* UniformMass.inl:469-496 */                                           \
        const sofa::type::Vec3d g(this->getContext()->getGravity());                                                                 \
        auto& ff = *f.beginEdit();                                                                                                   \
        if (sofab200_uniform_mass_add_force(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, ff.size(), sofa::b200::devWrite(ff), \
                                            double(d_vertexMass.getValue()), g.ptr()) != SOFAB200_OK)                                \
            msg_error() << sofab200_last_error();                                                                                    \
        f.endEdit();                                                                                                                 \
    }
B200_DIAGONAL_MASS(float)
B200_DIAGONAL_MASS(double)
template class DiagonalMass<sofa::b200::B200Vec3fTypes>;
template class DiagonalMass<sofa::b200::B200Vec3dTypes>;
template class UniformMass<sofa::b200::B200Vec3fTypes>;
template class UniformMass<sofa::b200::B200Vec3dTypes>;
}  // namespace sofa::component::mass

namespace sofa::b200 {
void registerDiagonalMass(sofa::core::ObjectFactory* factory) {
    using namespace sofa::component::mass;
    factory->registerObjects(sofa::core::ObjectRegistrationData("DiagonalMass on a B200 GPU (sofa_b200)").add<DiagonalMass<B200Vec3fTypes>>().add<DiagonalMass<B200Vec3dTypes>>());
}
void registerUniformMass(sofa::core::ObjectFactory* factory) {
    using namespace sofa::component::mass;
    factory->registerObjects(sofa::core::ObjectRegistrationData("UniformMass on a B200 GPU (sofa_b200)").add<UniformMass<B200Vec3fTypes>>().add<UniformMass<B200Vec3dTypes>>());
}
}  // namespace sofa::b200
