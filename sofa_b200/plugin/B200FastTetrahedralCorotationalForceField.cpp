// SOFA-side glue: FastTetrahedralCorotationalForceField<B200Vec3fTypes / B200Vec3dTypes>.
// The class reserves FastTetrahedralCorotationalForceFieldData<DataTypes> "for additional storage within template specializations"
// (FastTetrahedralCorotationalForceField.h:43-50, member m_data :178-179): the device handle lives there.  init() is replaced as a whole -- the
// per-tetrahedron precomputation (.inl:38-150), the edge orientations (:249-270) and the per-edge matrices of addDForce (:414-450) are kept on
// the device by sofab200_tetfem_create with sofab200_tetfem_desc::fast_corotational = 1; the host-side TopologyData (d_tetrahedronInfo,
// d_edgeInfo, d_pointInfo) stay empty, so topological changes and addKToMatrix / buildStiffnessMatrix are not served by this specialisation.
// Not compiled in this repository (no SOFA tree here): see INTEGRATION.md.
#include <sofa/component/solidmechanics/fem/elastic/FastTetrahedralCorotationalForceField.h>

#include "B200Types.h"

namespace sofa::component::solidmechanics::fem::elastic {
using sofa::b200::B200Vec3Types;

template <class TReal> class FastTetrahedralCorotationalForceFieldData<B200Vec3Types<TReal>> {
public:
    typedef FastTetrahedralCorotationalForceField<B200Vec3Types<TReal>> Main;
    sofab200_tetfem* ff = nullptr;
    void reinit(Main*) {}
    ~FastTetrahedralCorotationalForceFieldData() { if (ff) sofab200_tetfem_destroy(ff); }
};
}  // namespace sofa::component::solidmechanics::fem::elastic

#include <sofa/component/solidmechanics/fem/elastic/FastTetrahedralCorotationalForceField.inl>
#include <sofa/core/ObjectFactory.h>

namespace sofa::component::solidmechanics::fem::elastic {

#define B200_FASTTET(TReal)                                                                                                          \
    template <> void FastTetrahedralCorotationalForceField<B200Vec3Types<TReal>>::init() {                                           \
        this->Inherited::init();                                                                                                     \
        if (this->d_componentState.getValue() == sofa::core::objectmodel::ComponentState::Invalid) return;                           \
        if (this->l_topology->getNbTetrahedra() == 0) msg_error() << "No tetrahedra found in linked Topology.";                      \
        const std::string& method = d_method.getValue();   /* .inl:191-203 */                                                        \
        sofab200_tetfem_desc desc{};                                                                                                 \
        desc.fast_corotational = 1;                                                                                                  \
        if (method == "polar") { m_decompositionMethod = POLAR_DECOMPOSITION; desc.method = SOFAB200_TET_POLAR; }                    \
        else if (method == "qr" || method == "large") { m_decompositionMethod = QR_DECOMPOSITION; desc.method = SOFAB200_TET_LARGE; } \
        else if (method == "polar2") { m_decompositionMethod = POLAR_DECOMPOSITION_MODIFIED; desc.method = SOFAB200_TET_POLAR2; }    \
        else if (method == "none" || method == "linear" || method == "small") { m_decompositionMethod = LINEAR_ELASTIC; desc.method = SOFAB200_TET_SMALL; } \
        else { msg_error() << "cannot recognize method " << method << ". Must be either qr, polar, polar2 or none"; desc.method = SOFAB200_TET_LARGE; } \
        const auto& rest = this->mstate->read(core::vec_id::read_access::restPosition)->getValue();                                  \
        const auto& tetras = this->l_topology->getTetrahedra();                                                                      \
        const auto& edges = this->l_topology->getEdges();   /* the container's own numbering, whatever created it */               \
        std::vector<double> young(this->d_youngModulus.getValue().begin(), this->d_youngModulus.getValue().end());                  \
        std::vector<double> poisson(this->d_poissonRatio.getValue().begin(), this->d_poissonRatio.getValue().end());                \
        desc.n_young = young.size(); desc.young = young.data();                                                                      \
        desc.n_poisson = poisson.size(); desc.poisson = poisson.data();                                                              \
        desc.n_edges = edges.size(); desc.edges = reinterpret_cast<const uint32_t*>(edges.data());                                   \
        if (m_data.ff) { sofab200_tetfem_destroy(m_data.ff); m_data.ff = nullptr; }                                                  \
        if (sofab200_tetfem_create(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, rest.size(), rest.hostRead(), tetras.size(), \
                                   reinterpret_cast<const uint32_t*>(tetras.data()), &desc, &m_data.ff) != SOFAB200_OK) {           \
            msg_error() << "sofa_b200: " << sofab200_last_error();                                                                   \
            this->d_componentState.setValue(core::objectmodel::ComponentState::Invalid);                                            \
        }                                                                                                                            \
    }                                                                                                                                \
    template <> void FastTetrahedralCorotationalForceField<B200Vec3Types<TReal>>::addForce(const core::MechanicalParams*, DataVecDeriv& d_f, \
                                                                                         const DataVecCoord& d_x, const DataVecDeriv&) { \
        VecDeriv& f = *d_f.beginEdit();                                                                                              \
        const VecCoord& x = d_x.getValue();                                                                                          \
        f.resize(x.size());                                                                                                          \
        if (sofab200_tetfem_add_force(m_data.ff, f.deviceWrite(), x.deviceRead()) != SOFAB200_OK) msg_error() << sofab200_last_error(); \
        updateMatrix = true;   /* .inl:396 (the library re-assembles the edge matrices at its next addDForce) */                    \
        d_f.endEdit();                                                                                                               \
    }                                                                                                                                \
    template <> void FastTetrahedralCorotationalForceField<B200Vec3Types<TReal>>::addDForce(const core::MechanicalParams* mparams,   \
                                                                                          DataVecDeriv& d_df, const DataVecDeriv& d_dx) { \
        VecDeriv& df = *d_df.beginEdit();                                                                                            \
        const VecDeriv& dx = d_dx.getValue();                                                                                        \
        df.resize(dx.size());                                                                                                        \
        const double k = sofa::core::mechanicalparams::kFactorIncludingRayleighDamping(mparams, this->rayleighStiffness.getValue()); /* .inl:408 */ \
        if (sofab200_tetfem_add_dforce(m_data.ff, df.deviceWrite(), dx.deviceRead(), k) != SOFAB200_OK) msg_error() << sofab200_last_error(); \
        updateMatrix = false;                                                                                                        \
        d_df.endEdit();                                                                                                              \
    }
B200_FASTTET(float)
B200_FASTTET(double)

template class FastTetrahedralCorotationalForceField<sofa::b200::B200Vec3fTypes>;
template class FastTetrahedralCorotationalForceField<sofa::b200::B200Vec3dTypes>;
}  // namespace sofa::component::solidmechanics::fem::elastic

namespace sofa::b200 {
void registerFastTetrahedralCorotationalForceField(sofa::core::ObjectFactory* factory) {
    using namespace sofa::component::solidmechanics::fem::elastic;
    typedef FastTetrahedralCorotationalForceField<B200Vec3fTypes> FastF;
    typedef FastTetrahedralCorotationalForceField<B200Vec3dTypes> FastD;
    factory->registerObjects(sofa::core::ObjectRegistrationData("FastTetrahedralCorotationalForceField on a B200 GPU (sofa_b200)").add<FastF>().add<FastD>());
}
}  // namespace sofa::b200
