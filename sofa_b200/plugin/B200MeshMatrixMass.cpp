// SOFA-side glue: MeshMatrixMass<B200Vec3fTypes / B200Vec3dTypes>.
// init() / massInitialization() stay the reference's own (d_vertexMass, d_edgeMass, m_massLumpingCoeff as the CPU class computes them,
// MeshMatrixMass.inl:547-665,1400-1500); the device applies the matrix.  Device state hangs off the MeshMatrixMassInternalData member
// the class reserves (MeshMatrixMass.h:46,128).  Not compiled in this repository (no SOFA tree here): see INTEGRATION.md.
#include <sofa/component/mass/MeshMatrixMass.inl>
#include <sofa/core/ObjectFactory.h>

#include "B200Types.h"

namespace sofa::component::mass {
using sofa::b200::B200Vec3Types;

template <class TReal, class MassType, class GeometricalTypes> class MeshMatrixMassInternalData<B200Vec3Types<TReal>, MassType, GeometricalTypes> {
public:
    typedef MeshMatrixMass<B200Vec3Types<TReal>, GeometricalTypes> Main;
    sofab200_meshmass* mm = nullptr;
    ~MeshMatrixMassInternalData() { if (mm) sofab200_meshmass_destroy(mm); }
    // called lazily by the virtuals below (and again after a topological change): hand the host arrays over.  This class is a friend of
    // MeshMatrixMass (MeshMatrixMass.h:129), which is what gives it m_massLumpingCoeff.
    static int upload(Main* m) {
        auto& self = m->data;
        if (self.mm) { sofab200_meshmass_destroy(self.mm); self.mm = nullptr; }
        const auto& vm = m->d_vertexMass.getValue(); const auto& em = m->d_edgeMass.getValue();
        const auto& edges = m->l_topology->getEdges();   // topology order = the order addMDx adds the edge terms in
        return sofab200_meshmass_create(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, vm.size(), vm.data(), edges.size(),
                                        reinterpret_cast<const uint32_t*>(edges.data()), em.data(), m->isLumped() ? 1 : 0,
                                        double(m->m_massLumpingCoeff), &self.mm);
    }
};

#define B200_MESHMASS(TReal)                                                                                                         \
    template <> void MeshMatrixMass<B200Vec3Types<TReal>>::addMDx(const core::MechanicalParams*, DataVecDeriv& vres, const DataVecDeriv& vdx, SReal factor) { \
        if (!data.mm && decltype(data)::upload(this) != SOFAB200_OK) { msg_error() << sofab200_last_error(); return; }                          \
        auto& res = *vres.beginEdit();                                                                                               \
        if (sofab200_meshmass_add_mdx(data.mm, res.deviceWrite(), vdx.getValue().deviceRead(), factor) != SOFAB200_OK) msg_error() << sofab200_last_error(); \
        vres.endEdit();                                                                                                              \
    }                                                                                                                                \
    template <> void MeshMatrixMass<B200Vec3Types<TReal>>::addForce(const core::MechanicalParams*, DataVecDeriv& vf, const DataVecCoord&, const DataVecDeriv&) { \
        if (this->m_separateGravity.getValue()) return;                                                                              \
        if (!data.mm && decltype(data)::upload(this) != SOFAB200_OK) { msg_error() << sofab200_last_error(); return; }                          \
        const sofa::type::Vec3d g(this->getContext()->getGravity());                                                                 \
        auto& f = *vf.beginEdit();                                                                                                   \
        if (sofab200_meshmass_add_force(data.mm, f.deviceWrite(), g.ptr()) != SOFAB200_OK) msg_error() << sofab200_last_error();     \
        vf.endEdit();                                                                                                                \
    }                                                                                                                                \
    template <> void MeshMatrixMass<B200Vec3Types<TReal>>::accFromF(const core::MechanicalParams*, DataVecDeriv& a, const DataVecDeriv& f) { \
        if (!isLumped()) { msg_error() << "the method 'accFromF' can't be used with MeshMatrixMass as this SPARSE mass matrix can't be inversed easily."; return; } \
        if (!data.mm && decltype(data)::upload(this) != SOFAB200_OK) { msg_error() << sofab200_last_error(); return; }                          \
        auto& acc = *a.beginEdit();                                                                                                  \
        if (sofab200_meshmass_acc_from_f(data.mm, acc.deviceWrite(), f.getValue().deviceRead()) != SOFAB200_OK) msg_error() << sofab200_last_error(); \
        a.endEdit();                                                                                                                 \
    }
B200_MESHMASS(float)
B200_MESHMASS(double)

template class MeshMatrixMass<sofa::b200::B200Vec3fTypes>;
template class MeshMatrixMass<sofa::b200::B200Vec3dTypes>;
}  // namespace sofa::component::mass

namespace sofa::b200 {
void registerMeshMatrixMass(sofa::core::ObjectFactory* factory) {
    using namespace sofa::component::mass;
    factory->registerObjects(sofa::core::ObjectRegistrationData("MeshMatrixMass on a B200 GPU (sofa_b200)")
                                 .add<MeshMatrixMass<B200Vec3fTypes>>()
                                 .add<MeshMatrixMass<B200Vec3dTypes>>());
}
}  // namespace sofa::b200
