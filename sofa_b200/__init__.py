"""sofa_b200 -- B200-native implicit-dynamics FEM hot path of SOFA behind a C ABI.

Host-side mirror (Python over ctypes; PyTorch only provides device memory and streams) of the reference
components on the path: MechanicalObject vector ops, TetrahedronFEMForceField / HexahedronFEMForceField, TetrahedralCorotationalFEMForceField,
DiagonalMass, FixedProjectiveConstraint, and the EulerImplicitSolver + CGLinearSolver node.
The computing is done by sofa_b200/lib/libsofa_b200.so (hand-written sm_100a CUDA); nothing here computes.
"""
from ._lib import F32, F64, Sofab200Error, load  # noqa: F401
from .components import (Communicator, Context, DiagonalMass, MeshMatrixMass, UniformMass, PlaneForceField, FixedProjectiveConstraint, HexahedronFEMForceField,  # noqa: F401
                         MechanicalObject, SolverNode, TetrahedralCorotationalFEMForceField, TetrahedronFEMForceField,
                         FastTetrahedralCorotationalForceField)
from . import topology  # noqa: F401

__all__ = ["Context", "Communicator", "MechanicalObject", "TetrahedronFEMForceField", "TetrahedralCorotationalFEMForceField", "FastTetrahedralCorotationalForceField", "HexahedronFEMForceField", "DiagonalMass", "MeshMatrixMass", "UniformMass", "PlaneForceField",
           "FixedProjectiveConstraint", "SolverNode", "topology", "load", "Sofab200Error", "F32", "F64"]
