// SOFA-side glue: HexahedronFEMForceField<B200Vec3fTypes / B200Vec3dTypes>.
// Same pattern as B200TetrahedronFEMForceField.cpp: the class, its Data fields (youngModulus, poissonRatio, method,
// rayleighStiffness) and init() are the reference's own template; reinit / addForce / addDForce forward to ONE entry point of
// include/sofa_b200.h each.  Device state hangs off the HexahedronFEMForceFieldInternalData member the reference class
// reserves (HexahedronFEMForceField.h:41-54,188-189).  Syntax-checked against the reference headers by tools/plugin_syntax_check.sh.
#include <sofa/component/solidmechanics/fem/elastic/HexahedronFEMForceField.inl>
#include <sofa/core/ObjectFactory.h>

#include "B200Handles.h"

namespace sofa::component::solidmechanics::fem::elastic {
using sofa::b200::B200Vec3Types;

template <class TReal> class HexahedronFEMForceFieldInternalData<B200Vec3Types<TReal>> {
public:
    sofab200_hexfem* ff = nullptr;
    void initPtrData(HexahedronFEMForceField<B200Vec3Types<TReal>>*) {}
    static sofab200_hexfem* handle(HexahedronFEMForceField<B200Vec3Types<TReal>>* m) { return m->data ? m->data->ff : nullptr; }   // (called by the class's constructor, HexahedronFEMForceField.inl:65)
    ~HexahedronFEMForceFieldInternalData() { if (ff) sofab200_hexfem_destroy(ff); }
};

#define B200_HEXFEM(TReal)                                                                                                           \
    template <> void HexahedronFEMForceField<B200Vec3Types<TReal>>::reinit() {                                                      \
        /* replaces reinit() .inl:155-192: material stiffness, rest rotations, rotated rest shapes and the 24x24 element           \
           stiffness matrices (computeElementStiffness .inl:306-536) are computed inside sofab200_hexfem_create */                   \
        if (this->d_componentState.getValue() == core::objectmodel::ComponentState::Invalid) return;                                \
        { const std::string& m = d_method.getValue(); setMethod(m == "large" ? LARGE : (m == "polar" ? POLAR : SMALL)); } /* as init() does; LARGE=0, POLAR=1, SMALL=2 in both enums */ \
        const auto& rest = this->mstate->read(core::vec_id::read_access::restPosition)->getValue();                                 \
        const auto& hexas = this->l_topology->getHexahedra();                                                                        \
        std::vector<double> young(this->d_youngModulus.getValue().begin(), this->d_youngModulus.getValue().end());                 \
        std::vector<double> poisson(this->d_poissonRatio.getValue().begin(), this->d_poissonRatio.getValue().end());               \
        sofab200_hexfem_desc desc{};                                                                                                 \
        desc.method = int(method);                                                                                                   \
        desc.n_young = young.size(); desc.young = young.data();                                                                      \
        desc.n_poisson = poisson.size(); desc.poisson = poisson.data();                                                              \
        if (data->ff) { sofab200_hexfem_destroy(data->ff); data->ff = nullptr; }                                                     \
        const int rc = sofab200_hexfem_create(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, rest.size(), rest.hostRead(), \
                                              hexas.size(), reinterpret_cast<const uint32_t*>(hexas.data()), &desc, &data->ff);      \
        if (rc != SOFAB200_OK) {                                                                                                     \
            msg_error() << "sofa_b200: " << sofab200_last_error();                                                                   \
            this->d_componentState.setValue(core::objectmodel::ComponentState::Invalid);                                            \
        }                                                                                                                            \
    }                                                                                                                                \
    template <> void HexahedronFEMForceField<B200Vec3Types<TReal>>::addForce(const core::MechanicalParams*, DataVecDeriv& d_f,      \
                                                                             const DataVecCoord& d_x, const DataVecDeriv&) {        \
        VecDeriv& f = *d_f.beginEdit();                                                                                              \
        const VecCoord& x = d_x.getValue();                                                                                          \
        f.resize(x.size());                                                                                                          \
        if (sofab200_hexfem_add_force(data->ff, f.deviceWrite(), x.deviceRead()) != SOFAB200_OK) msg_error() << sofab200_last_error(); \
        d_f.endEdit();                                                                                                               \
    }                                                                                                                                \
    template <> void HexahedronFEMForceField<B200Vec3Types<TReal>>::addDForce(const core::MechanicalParams* mparams,                \
                                                                              DataVecDeriv& d_df, const DataVecDeriv& d_dx) {       \
        VecDeriv& df = *d_df.beginEdit();                                                                                            \
        const VecDeriv& dx = d_dx.getValue();                                                                                        \
        df.resize(dx.size());                                                                                                        \
        const double k = sofa::core::mechanicalparams::kFactorIncludingRayleighDamping(mparams, this->rayleighStiffness.getValue()); \
        if (sofab200_hexfem_add_dforce(data->ff, df.deviceWrite(), dx.deviceRead(), k) != SOFAB200_OK) msg_error() << sofab200_last_error(); \
        d_df.endEdit();                                                                                                              \
    }
B200_HEXFEM(float)
B200_HEXFEM(double)

template class HexahedronFEMForceField<sofa::b200::B200Vec3fTypes>;
template class HexahedronFEMForceField<sofa::b200::B200Vec3dTypes>;
}  // namespace sofa::component::solidmechanics::fem::elastic

namespace sofa::b200 {
sofab200_hexfem* hexfemHandle(sofa::component::solidmechanics::fem::elastic::HexahedronFEMForceField<B200Vec3fTypes>* ff) {
    return sofa::component::solidmechanics::fem::elastic::HexahedronFEMForceFieldInternalData<B200Vec3fTypes>::handle(ff);
}
sofab200_hexfem* hexfemHandle(sofa::component::solidmechanics::fem::elastic::HexahedronFEMForceField<B200Vec3dTypes>* ff) {
    return sofa::component::solidmechanics::fem::elastic::HexahedronFEMForceFieldInternalData<B200Vec3dTypes>::handle(ff);
}
void registerHexahedronFEMForceField(sofa::core::ObjectFactory* factory) {
    using namespace sofa::component::solidmechanics::fem::elastic;
    factory->registerObjects(sofa::core::ObjectRegistrationData("HexahedronFEMForceField on a B200 GPU (sofa_b200)")
                                 .add<HexahedronFEMForceField<B200Vec3fTypes>>()
                                 .add<HexahedronFEMForceField<B200Vec3dTypes>>());
}
}  // namespace sofa::b200
