// SOFA-side glue: a linear-solver component that keeps the whole CG solve on the device.
// SOFA's CGLinearSolver is templated on the matrix/vector type (GraphScattered), not on DataTypes
// (Sofa/Component/LinearSolver/Iterative/src/sofa/component/linearsolver/iterative/CGLinearSolver.cpp:68-80), so changing
// the MechanicalObject's template alone still runs the reference's host loop, with two blocking vDot per iteration.  This
// component registers under the same class name with template "B200GraphScattered": when the solver node's
// MechanicalObject, mass, force field and fixed constraint are all B200-typed it builds one sofab200_node and replaces
// solve() by sofab200_node_cg_solve (same Data: iterations, tolerance, threshold, warmStart, graph).  Otherwise it falls
// back to the inherited host loop over the per-op C-ABI calls.
#pragma once
#include <sofa/component/linearsolver/iterative/CGLinearSolver.h>
#include <sofa/component/linearsolver/iterative/GraphScatteredTypes.h>

#include "B200Types.h"

namespace sofa::b200 {
using sofa::component::linearsolver::GraphScatteredMatrix;
using sofa::component::linearsolver::GraphScatteredVector;

class B200CGLinearSolver : public sofa::component::linearsolver::iterative::CGLinearSolver<GraphScatteredMatrix, GraphScatteredVector> {
public:
    SOFA_CLASS(B200CGLinearSolver, SOFA_TEMPLATE2(sofa::component::linearsolver::iterative::CGLinearSolver, GraphScatteredMatrix, GraphScatteredVector));
    static std::string GetCustomClassName() { return "CGLinearSolver"; }
    static std::string GetCustomTemplateName() { return "B200GraphScattered"; }
    void bwdInit() override;   // discovers the B200-typed components of the node and calls sofab200_node_create
    void solve(GraphScatteredMatrix& A, GraphScatteredVector& x, GraphScatteredVector& b) override {
        if (!m_node) return Inherit1::solve(A, x, b);
        sofab200_solver_params p = paramsFromData();           // iterations / tolerance / threshold / warmStart + mparams factors
        sofab200_node_set_params(m_node, &p);
        int nbIter = 0;
        sofab200_node_cg_solve(m_node, devicePtr(x), devicePtr(b), A.parent->mparams.mFactor(), A.parent->mparams.bFactor(),
                               A.parent->mparams.kFactor(), &nbIter);
        sofa::helper::AdvancedTimer::valSet("CG iterations", nbIter);   // same timer value as CGLinearSolver.inl:301-306
        publishGraph();                                          // sofab200_node_last_solve -> d_graph ("Error", "Denominator")
    }
private:
    sofab200_node* m_node = nullptr;
    sofab200_solver_params paramsFromData() const;
    void* devicePtr(GraphScatteredVector& v);
    void publishGraph();
};
}  // namespace sofa::b200
