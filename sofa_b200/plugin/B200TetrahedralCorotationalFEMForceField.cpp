// SOFA-side glue: TetrahedralCorotationalFEMForceField<B200Vec3fTypes / B200Vec3dTypes> (what examples/Demos/liver.scn uses).
// The class has no *InternalData member, so the device handle lives in a side table keyed by the component.  Its arithmetic on this path is
// statement for statement TetrahedronFEMForceField's (TetrahedralCorotationalFEMForceField.inl:356-1175), hence the same C entry points with
// sofab200_tetfem_desc::tetrahedral_corotational = 1.  Not compiled in this repository (no SOFA tree here): see INTEGRATION.md.
#include <sofa/component/solidmechanics/fem/elastic/TetrahedralCorotationalFEMForceField.inl>
#include <sofa/core/ObjectFactory.h>

#include <mutex>
#include <unordered_map>

#include "B200Types.h"

namespace sofa::component::solidmechanics::fem::elastic {
using sofa::b200::B200Vec3Types;

namespace {
std::mutex g_mutex;
std::unordered_map<const void*, sofab200_tetfem*> g_handles;   // entries are dropped by the component's destructor
sofab200_tetfem*& handle(const void* self) { std::lock_guard<std::mutex> l(g_mutex); return g_handles[self]; }
}  // namespace

#define B200_TETCOROT(TReal)                                                                                                         \
    template <> void TetrahedralCorotationalFEMForceField<B200Vec3Types<TReal>>::reinit() {                                          \
        /* replaces reinit() .inl:122-160: the per-element precomputation runs inside sofab200_tetfem_create */                      \
        { const std::string& m = d_method.getValue(); setMethod(m == "small" ? SMALL : (m == "polar" ? POLAR : LARGE)); }   /* as init() does, .inl:135-147 */ \
        const auto& rest = this->mstate->read(core::vec_id::read_access::restPosition)->getValue();                                  \
        const auto& tetras = this->l_topology->getTetrahedra();                                                                      \
        std::vector<double> young(this->d_youngModulus.getValue().begin(), this->d_youngModulus.getValue().end());                  \
        std::vector<double> poisson(this->d_poissonRatio.getValue().begin(), this->d_poissonRatio.getValue().end());                \
        std::vector<double> lsf(d_localStiffnessFactor.getValue().begin(), d_localStiffnessFactor.getValue().end());                \
        sofab200_tetfem_desc desc{};                                                                                                 \
        desc.tetrahedral_corotational = 1;                                                                                           \
        desc.method = method == SMALL ? SOFAB200_TET_SMALL : (method == LARGE ? SOFAB200_TET_LARGE : SOFAB200_TET_POLAR);             \
        desc.n_young = young.size(); desc.young = young.data();                                                                      \
        desc.n_poisson = poisson.size(); desc.poisson = poisson.data();                                                              \
        desc.n_local_stiffness = lsf.size(); desc.local_stiffness = lsf.data();                                                     \
        desc.update_stiffness_matrix = d_updateStiffnessMatrix.getValue() ? 1 : 0;                                                   \
        sofab200_tetfem*& ff = handle(this);                                                                                         \
        if (ff) { sofab200_tetfem_destroy(ff); ff = nullptr; }                                                                       \
        if (sofab200_tetfem_create(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, rest.size(), rest.hostRead(), tetras.size(), \
                                   reinterpret_cast<const uint32_t*>(tetras.data()), &desc, &ff) != SOFAB200_OK) {                  \
            msg_error() << "sofa_b200: " << sofab200_last_error();                                                                   \
            this->d_componentState.setValue(core::objectmodel::ComponentState::Invalid);                                            \
        }                                                                                                                            \
    }                                                                                                                                \
    template <> void TetrahedralCorotationalFEMForceField<B200Vec3Types<TReal>>::addForce(const core::MechanicalParams*, DataVecDeriv& d_f, \
                                                                                          const DataVecCoord& d_x, const DataVecDeriv&) { \
        VecDeriv& f = *d_f.beginEdit();                                                                                              \
        const VecCoord& x = d_x.getValue();                                                                                          \
        f.resize(x.size());                                                                                                          \
        if (sofab200_tetfem_add_force(handle(this), f.deviceWrite(), x.deviceRead()) != SOFAB200_OK) msg_error() << sofab200_last_error(); \
        d_f.endEdit();                                                                                                               \
    }                                                                                                                                \
    template <> void TetrahedralCorotationalFEMForceField<B200Vec3Types<TReal>>::addDForce(const core::MechanicalParams* mparams,    \
                                                                                           DataVecDeriv& d_df, const DataVecDeriv& d_dx) { \
        VecDeriv& df = *d_df.beginEdit();                                                                                            \
        const VecDeriv& dx = d_dx.getValue();                                                                                        \
        df.resize(dx.size());                                                                                                        \
        const double k = sofa::core::mechanicalparams::kFactorIncludingRayleighDamping(mparams, this->rayleighStiffness.getValue()); /* .inl:208 */ \
        if (sofab200_tetfem_add_dforce(handle(this), df.deviceWrite(), dx.deviceRead(), k) != SOFAB200_OK) msg_error() << sofab200_last_error(); \
        d_df.endEdit();                                                                                                              \
    }                                                                                                                                \
    /* (the class declares no cleanup(); its destructor is the one member that runs when the component goes away) */              \
    template <> TetrahedralCorotationalFEMForceField<B200Vec3Types<TReal>>::~TetrahedralCorotationalFEMForceField() {               \
        std::lock_guard<std::mutex> l(g_mutex);                                                                                      \
        auto it = g_handles.find(this);                                                                                              \
        if (it != g_handles.end()) { if (it->second) sofab200_tetfem_destroy(it->second); g_handles.erase(it); }                     \
    }
B200_TETCOROT(float)
B200_TETCOROT(double)

template class TetrahedralCorotationalFEMForceField<sofa::b200::B200Vec3fTypes>;
template class TetrahedralCorotationalFEMForceField<sofa::b200::B200Vec3dTypes>;
}  // namespace sofa::component::solidmechanics::fem::elastic

namespace sofa::b200 {
void registerTetrahedralCorotationalFEMForceField(sofa::core::ObjectFactory* factory) {
    using namespace sofa::component::solidmechanics::fem::elastic;
    factory->registerObjects(sofa::core::ObjectRegistrationData("TetrahedralCorotationalFEMForceField on a B200 GPU (sofa_b200)")
                                 .add<TetrahedralCorotationalFEMForceField<B200Vec3fTypes>>()
                                 .add<TetrahedralCorotationalFEMForceField<B200Vec3dTypes>>());
}
}  // namespace sofa::b200
