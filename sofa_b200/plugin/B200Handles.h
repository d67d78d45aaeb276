// The C-ABI handles the B200 force-field components own, for the component that assembles them into one device-resident solver node
// (B200CGLinearSolver).  Defined next to each component's InternalData (the only friend of the component class).
#pragma once
#include <sofa/component/solidmechanics/fem/elastic/HexahedronFEMForceField.h>
#include <sofa/component/solidmechanics/fem/elastic/TetrahedronFEMForceField.h>

#include "B200Types.h"

namespace sofa::b200 {
sofab200_tetfem* tetfemHandle(sofa::component::solidmechanics::fem::elastic::TetrahedronFEMForceField<B200Vec3fTypes>* ff);
sofab200_tetfem* tetfemHandle(sofa::component::solidmechanics::fem::elastic::TetrahedronFEMForceField<B200Vec3dTypes>* ff);
sofab200_hexfem* hexfemHandle(sofa::component::solidmechanics::fem::elastic::HexahedronFEMForceField<B200Vec3fTypes>* ff);
sofab200_hexfem* hexfemHandle(sofa::component::solidmechanics::fem::elastic::HexahedronFEMForceField<B200Vec3dTypes>* ff);
}  // namespace sofa::b200
