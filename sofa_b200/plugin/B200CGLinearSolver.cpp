// SOFA-side glue: CGLinearSolver<B200GraphScattered> (see B200CGLinearSolver.h).
#include "B200CGLinearSolver.h"

#include <sofa/component/constraint/projective/FixedProjectiveConstraint.h>
#include <sofa/component/mass/DiagonalMass.h>
#include <sofa/component/mass/UniformMass.h>
#include <sofa/component/mechanicalload/PlaneForceField.h>
#include <sofa/component/statecontainer/MechanicalObject.h>
#include <sofa/core/ObjectFactory.h>
#include <sofa/core/behavior/BaseForceField.h>
#include <sofa/core/behavior/BaseMass.h>
#include <sofa/core/behavior/BaseProjectiveConstraintSet.h>
#include <sofa/helper/AdvancedTimer.h>

#include "B200Handles.h"

namespace sofa::b200 {
using namespace sofa::component;

B200CGLinearSolver::~B200CGLinearSolver() { if (m_node) sofab200_node_destroy(m_node); }

template <class DataTypes> bool B200CGLinearSolver::build() {
    typedef typename DataTypes::Real Real;
    auto* ctx = this->getContext();
    auto* mo = ctx->template get<statecontainer::MechanicalObject<DataTypes>>();
    if (!mo) return false;
    // every force field / mass / projective constraint of the node must be one this glue knows how to put into the device node
    type::vector<core::behavior::BaseForceField*> allFF;
    ctx->template get<core::behavior::BaseForceField>(&allFF, core::objectmodel::BaseContext::Local);
    type::vector<core::behavior::BaseProjectiveConstraintSet*> allPC;
    ctx->template get<core::behavior::BaseProjectiveConstraintSet>(&allPC, core::objectmodel::BaseContext::Local);
    auto* tet = ctx->template get<solidmechanics::fem::elastic::TetrahedronFEMForceField<DataTypes>>();
    auto* hex = ctx->template get<solidmechanics::fem::elastic::HexahedronFEMForceField<DataTypes>>();
    auto* dmass = ctx->template get<mass::DiagonalMass<DataTypes>>();
    auto* umass = ctx->template get<mass::UniformMass<DataTypes>>();
    auto* plane = ctx->template get<mechanicalload::PlaneForceField<DataTypes>>();
    auto* fixed = ctx->template get<constraint::projective::FixedProjectiveConstraint<DataTypes>>();
    if ((tet != nullptr) == (hex != nullptr)) return false;          // exactly one FEM force field
    const size_t known = 1 + (dmass ? 1 : 0) + (umass ? 1 : 0) + (plane ? 1 : 0);
    if (allFF.size() != known || (dmass && umass)) return false;     // (a Mass is a ForceField)
    if (allPC.size() != (fixed ? 1u : 0u)) return false;

    sofab200_node_desc desc{};
    if (tet) desc.tetfem = tetfemHandle(tet); else desc.hexfem = hexfemHandle(hex);
    if (!desc.tetfem && !desc.hexfem) return false;
    std::vector<Real> vm;
    if (dmass) { const auto& m = dmass->d_vertexMass.getValue(); vm.assign(m.begin(), m.end()); desc.vertex_mass_host = vm.data(); }
    if (umass) { desc.uniform_mass = 1; desc.uniform_vertex_mass = double(umass->d_vertexMass.getValue()); }
    std::vector<uint32_t> idx;
    if (fixed) { const auto& f = fixed->d_indices.getValue(); idx.assign(f.begin(), f.end()); desc.n_fixed = idx.size(); desc.fixed_host = idx.data(); desc.fix_all = fixed->d_fixAll.getValue() ? 1 : 0; }
    sofab200_plane_desc pd{};
    if (plane) {
        const auto n = plane->d_planeNormal.getValue();
        pd.normal[0] = n[0]; pd.normal[1] = n[1]; pd.normal[2] = n[2]; pd.d = plane->d_planeD.getValue(); pd.stiffness = plane->d_stiffness.getValue();
        pd.damping = plane->d_damping.getValue(); pd.max_force = plane->d_maxForce.getValue(); pd.bilateral = plane->d_bilateral.getValue() ? 1 : 0;
        desc.plane = &pd; desc.plane_rayleigh_stiffness = plane->rayleighStiffness.getValue();
    }
    // order of the mass and the force field in the node (the reference's visitors add their terms in scene order)
    desc.mass_first = 1;
    if (dmass || umass) {
        core::behavior::BaseForceField* massFF = dmass ? static_cast<core::behavior::BaseForceField*>(dmass) : static_cast<core::behavior::BaseForceField*>(umass);
        core::behavior::BaseForceField* femFF = tet ? static_cast<core::behavior::BaseForceField*>(tet) : static_cast<core::behavior::BaseForceField*>(hex);
        for (auto* f : allFF) { if (f == massFF) { desc.mass_first = 1; break; } if (f == femFF) { desc.mass_first = 0; break; } }
    }
    m_ffRayleighStiffness = tet ? tet->rayleighStiffness.getValue() : hex->rayleighStiffness.getValue();
    m_massRayleighMass = dmass ? dmass->rayleighMass.getValue() : (umass ? umass->rayleighMass.getValue() : 0.0);
    if (sofab200_node_create(threadContext(), DataTypes::abiReal, mo->getSize(), &desc, &m_node) != SOFAB200_OK) {
        msg_warning() << "sofa_b200: " << sofab200_last_error() << " -- falling back to the host CG loop";
        m_node = nullptr;
        return false;
    }
    m_state = mo;
    m_double = sizeof(Real) == 8;
    return true;
}

void B200CGLinearSolver::bwdInit() {
    if (m_node) { sofab200_node_destroy(m_node); m_node = nullptr; }
    if (!build<B200Vec3fTypes>() && !build<B200Vec3dTypes>())
        msg_info() << "the node holds components the device-resident solver does not know: CGLinearSolver runs its host loop over the per-operation kernels";
}

void* B200CGLinearSolver::devicePtr(GraphScatteredVector& v, bool write) {
    const core::VecDerivId id = v.id().getId(m_state);
    if (m_double) {
        auto* d = static_cast<statecontainer::MechanicalObject<B200Vec3dTypes>*>(m_state)->write(id);
        if (write) { void* p = devWrite(*d->beginEdit()); d->endEdit(); return p; }
        return const_cast<void*>(devRead(d->getValue()));
    }
    auto* d = static_cast<statecontainer::MechanicalObject<B200Vec3fTypes>*>(m_state)->write(id);
    if (write) { void* p = devWrite(*d->beginEdit()); d->endEdit(); return p; }
    return const_cast<void*>(devRead(d->getValue()));
}

void B200CGLinearSolver::solve(GraphScatteredMatrix& A, GraphScatteredVector& x, GraphScatteredVector& b) {
    if (!m_node) { Inherit::solve(A, x, b); return; }
    sofab200_solver_params p{};
    const auto g = this->getContext()->getGravity();
    p.gravity[0] = g[0]; p.gravity[1] = g[1]; p.gravity[2] = g[2];
    p.dt = this->getContext()->getDt();
    p.iterations = d_maxIter.getValue(); p.tolerance = d_tolerance.getValue(); p.threshold = d_smallDenominatorThreshold.getValue();
    p.warm_start = d_warmStart.getValue() ? 1 : 0;
    p.ff_rayleigh_stiffness = m_ffRayleighStiffness; p.mass_rayleigh_mass = m_massRayleighMass;
    sofab200_node_set_params(m_node, &p);
    int nbIter = 0;
    if (sofab200_node_cg_solve(m_node, devicePtr(x, true), devicePtr(b, false), A.mparams.mFactor(), A.mparams.bFactor(), A.mparams.kFactor(), &nbIter) != SOFAB200_OK) {
        msg_error() << sofab200_last_error();
        return;
    }
    sofa::helper::AdvancedTimer::valSet("CG iterations", nbIter);   // same timer value as CGLinearSolver.inl:301-306
    publishGraph();
}

void B200CGLinearSolver::publishGraph() {
    constexpr size_t cap = 1100;
    std::vector<double> err(cap), den(cap);
    size_t nErr = 0, nDen = 0;
    int it = 0, end = 0;
    if (sofab200_node_last_solve(m_node, &it, &end, err.data(), &nErr, den.data(), &nDen, cap) != SOFAB200_OK) return;
    auto& graph = *d_graph.beginEdit();
    auto& ge = graph["Error"]; ge.assign(err.begin(), err.begin() + std::min(nErr, cap));
    auto& gd = graph["Denominator"]; gd.assign(den.begin(), den.begin() + std::min(nDen, cap));
    d_graph.endEdit();
}

void registerCGLinearSolver(sofa::core::ObjectFactory* factory) {
    factory->registerObjects(sofa::core::ObjectRegistrationData("Conjugate gradient whose whole loop runs on a B200 GPU (sofa_b200)").add<B200CGLinearSolver>());
}
}  // namespace sofa::b200
