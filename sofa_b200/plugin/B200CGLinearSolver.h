// SOFA-side glue: a linear-solver component that keeps the whole CG solve on the device.
// SOFA's CGLinearSolver is templated on the matrix/vector type (GraphScattered), not on DataTypes
// (Sofa/Component/LinearSolver/Iterative/src/sofa/component/linearsolver/iterative/CGLinearSolver.cpp:68-80), so changing the MechanicalObject's
// template alone still runs the reference's host loop, with two blocking vDot per iteration.  This component registers under the same class name
// with template "B200GraphScattered": when the solver node's MechanicalObject, mass, force field(s) and fixed constraint are all B200-typed it builds
// one sofab200_node (bwdInit) and replaces solve() by sofab200_node_cg_solve (same Data: iterations, tolerance, threshold, warmStart, graph).
// Any component it does not know makes it fall back to the inherited host loop over the per-operation C-ABI calls, which is always correct.
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
    typedef sofa::component::linearsolver::iterative::CGLinearSolver<GraphScatteredMatrix, GraphScatteredVector> Inherit;
    static std::string GetCustomClassName() { return "CGLinearSolver"; }
    static std::string GetCustomTemplateName() { return "B200GraphScattered"; }
    ~B200CGLinearSolver() override;
    void bwdInit() override;   // discovers the B200-typed components of the node and calls sofab200_node_create
    void solve(GraphScatteredMatrix& A, GraphScatteredVector& x, GraphScatteredVector& b) override;

private:
    sofab200_node* m_node = nullptr;
    sofa::core::behavior::BaseMechanicalState* m_state = nullptr;
    bool m_double = false;
    double m_ffRayleighStiffness = 0.0, m_massRayleighMass = 0.0;
    template <class DataTypes> bool build();                 // true when every component of the node is known and B200-typed
    void* devicePtr(GraphScatteredVector& v, bool write);
    void publishGraph();                                      // sofab200_node_last_solve -> d_graph ("Error", "Denominator")
};
}  // namespace sofa::b200
