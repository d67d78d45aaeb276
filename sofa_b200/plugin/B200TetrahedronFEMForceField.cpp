// SOFA-side glue: TetrahedronFEMForceField<B200Vec3fTypes / B200Vec3dTypes>.
// The class, its Data fields (youngModulus, poissonRatio, method, localStiffnessFactor, rayleighStiffness, ...) and its
// init() are the reference's own template; only the virtuals on the hot path are specialised, each forwarding to ONE
// entry point of include/sofa_b200.h.  Per-class device state hangs off the *InternalData member the reference class
// reserves for exactly this (TetrahedronFEMForceField.h:57-70,162-163).
#include <sofa/component/solidmechanics/fem/elastic/TetrahedronFEMForceField.inl>
#include <sofa/core/ObjectFactory.h>

#include "B200Handles.h"

namespace sofa::component::solidmechanics::fem::elastic {
using sofa::b200::B200Vec3Types;

template <class TReal> class TetrahedronFEMForceFieldInternalData<B200Vec3Types<TReal>> {
public:
    typedef TetrahedronFEMForceField<B200Vec3Types<TReal>> Main;
    sofab200_tetfem* ff = nullptr;
    sofa::b200::B200Vector<TReal> vmElem, vmNode;   // device-side results of computeVonMisesStress (the Data of the class are host vectors)
    void initPtrData(Main*) {}
    static sofab200_tetfem* handle(Main* m) { return m->data.ff; }
    ~TetrahedronFEMForceFieldInternalData() { if (ff) sofab200_tetfem_destroy(ff); }
};

#define B200_TETFEM(TReal)                                                                                                          \
    template <> void TetrahedronFEMForceField<B200Vec3Types<TReal>>::reinit() {                                                    \
        /* replaces reinit() .inl:1390-1505: the per-element precomputation runs inside sofab200_tetfem_create */                  \
        if (this->d_componentState.getValue() == core::objectmodel::ComponentState::Invalid) return;                                \
        if (!this->l_topology->getTetrahedra().empty()) _indexedElements = &this->l_topology->getTetrahedra();                      \
        setMethod(d_method.getValue());                                                                                              \
        const auto& rest = this->mstate->read(core::vec_id::read_access::restPosition)->getValue();                                 \
        std::vector<double> young(this->d_youngModulus.getValue().begin(), this->d_youngModulus.getValue().end());                 \
        std::vector<double> poisson(this->d_poissonRatio.getValue().begin(), this->d_poissonRatio.getValue().end());               \
        std::vector<double> lsf(d_localStiffnessFactor.getValue().begin(), d_localStiffnessFactor.getValue().end());               \
        sofab200_tetfem_desc desc{};                                                                                                 \
        desc.method = int(method); /* SMALL=0, LARGE=1, POLAR=2, SVD=3 in both enums */                                              \
        desc.n_young = young.size(); desc.young = young.data();                                                                      \
        desc.n_poisson = poisson.size(); desc.poisson = poisson.data();                                                              \
        desc.n_local_stiffness = lsf.size(); desc.local_stiffness = lsf.data();                                                     \
        desc.plastic_max_threshold = double(d_plasticMaxThreshold.getValue());     /* plasticity branch of computeForce, .inl:357-371 */ \
        desc.plastic_yield_threshold = double(d_plasticYieldThreshold.getValue());                                                  \
        desc.plastic_creep = double(d_plasticCreep.getValue());                                                                      \
        desc.update_stiffness_matrix = d_updateStiffnessMatrix.getValue() ? 1 : 0;   /* large / polar / svd (see include/sofa_b200.h) */                   \
        desc.compute_von_mises = isComputeVonMisesStressMethodSet() ? int(d_computeVonMisesStress.getValue()) : 0;                   \
        if (data.ff) { sofab200_tetfem_destroy(data.ff); data.ff = nullptr; }                                                        \
        const int rc = sofab200_tetfem_create(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, rest.size(), rest.hostRead(), \
                                              _indexedElements->size(), reinterpret_cast<const uint32_t*>(_indexedElements->data()), \
                                              &desc, &data.ff);                                                                      \
        if (rc != SOFAB200_OK) {                                                                                                     \
            msg_error() << "sofa_b200: " << sofab200_last_error();                                                                   \
            this->d_componentState.setValue(core::objectmodel::ComponentState::Invalid);                                            \
        }                                                                                                                            \
    }                                                                                                                                \
    template <> void TetrahedronFEMForceField<B200Vec3Types<TReal>>::addForce(const core::MechanicalParams*, DataVecDeriv& d_f,     \
                                                                              const DataVecCoord& d_x, const DataVecDeriv&) {       \
        VecDeriv& f = *d_f.beginEdit();                                                                                              \
        const VecCoord& x = d_x.getValue();                                                                                          \
        f.resize(x.size());                                                                                                          \
        if (sofab200_tetfem_add_force(data.ff, f.deviceWrite(), x.deviceRead()) != SOFAB200_OK) msg_error() << sofab200_last_error(); \
        d_f.endEdit();                                                                                                               \
    }                                                                                                                                \
    template <> void TetrahedronFEMForceField<B200Vec3Types<TReal>>::addDForce(const core::MechanicalParams* mparams,               \
                                                                               DataVecDeriv& d_df, const DataVecDeriv& d_dx) {      \
        VecDeriv& df = *d_df.beginEdit();                                                                                            \
        const VecDeriv& dx = d_dx.getValue();                                                                                        \
        df.resize(dx.size());                                                                                                        \
        /* the CPU class's factor (.inl:1615), not SofaCUDA's bare kFactor (CudaTetrahedronFEMForceField.inl:643) */                 \
        const double k = sofa::core::mechanicalparams::kFactorIncludingRayleighDamping(mparams, this->rayleighStiffness.getValue()); \
        if (sofab200_tetfem_add_dforce(data.ff, df.deviceWrite(), dx.deviceRead(), k) != SOFAB200_OK) msg_error() << sofab200_last_error(); \
        d_df.endEdit();                                                                                                              \
    }                                                                                                                                \
    template <> void TetrahedronFEMForceField<B200Vec3Types<TReal>>::reset() { /* .inl:1380-1388 */                                  \
        if (data.ff && sofab200_tetfem_reset(data.ff) != SOFAB200_OK) msg_error() << sofab200_last_error();                          \
    }                                                                                                                                \
    template <> void TetrahedronFEMForceField<B200Vec3Types<TReal>>::computeVonMisesStress() { /* .inl:2196-2372, values only */   \
        if (!data.ff || !isComputeVonMisesStressMethodSet()) return;                                                                 \
        const VecCoord& x = this->mstate->read(core::vec_id::read_access::position)->getValue();                                     \
        data.vmElem.resize(_indexedElements->size()); data.vmNode.resize(x.size());                                                  \
        if (sofab200_tetfem_compute_von_mises(data.ff, x.deviceRead(), data.vmElem.deviceWrite(), data.vmNode.deviceWrite()) != SOFAB200_OK) \
            msg_error() << sofab200_last_error();                                                                                    \
        auto& vME = *d_vonMisesPerElement.beginEdit(); auto& vMN = *d_vonMisesPerNode.beginEdit();                                   \
        vME.assign(data.vmElem.hostRead(), data.vmElem.hostRead() + data.vmElem.size());   /* D2H on demand (vector_device) */        \
        vMN.assign(data.vmNode.hostRead(), data.vmNode.hostRead() + data.vmNode.size());                                             \
        d_vonMisesPerElement.endEdit(); d_vonMisesPerNode.endEdit();                                                                 \
        updateVonMisesStress = false;                                                                                                \
    }                                                                                                                                \
    template <> void TetrahedronFEMForceField<B200Vec3Types<TReal>>::getRotations(VecReal& vecR) { /* .inl:2033-2042 */             \
        vecR.resize(9 * this->mstate->getSize());                                                                                    \
        if (sofab200_tetfem_get_rotations(data.ff, vecR.deviceWrite()) != SOFAB200_OK) msg_error() << sofab200_last_error();         \
    }
B200_TETFEM(float)
B200_TETFEM(double)

template class TetrahedronFEMForceField<sofa::b200::B200Vec3fTypes>;
template class TetrahedronFEMForceField<sofa::b200::B200Vec3dTypes>;
}  // namespace sofa::component::solidmechanics::fem::elastic

namespace sofa::b200 {
using namespace sofa::component::solidmechanics::fem::elastic;
sofab200_tetfem* tetfemHandle(TetrahedronFEMForceField<B200Vec3fTypes>* ff) { return TetrahedronFEMForceFieldInternalData<B200Vec3fTypes>::handle(ff); }
sofab200_tetfem* tetfemHandle(TetrahedronFEMForceField<B200Vec3dTypes>* ff) { return TetrahedronFEMForceFieldInternalData<B200Vec3dTypes>::handle(ff); }
void registerTetrahedronFEMForceField(sofa::core::ObjectFactory* factory) {
    using namespace sofa::component::solidmechanics::fem::elastic;
    factory->registerObjects(sofa::core::ObjectRegistrationData("TetrahedronFEMForceField on a B200 GPU (sofa_b200)")
                                 .add<TetrahedronFEMForceField<B200Vec3fTypes>>()
                                 .add<TetrahedronFEMForceField<B200Vec3dTypes>>());
}
}  // namespace sofa::b200
