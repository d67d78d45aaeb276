// SOFA-side glue: the generic components a scene around the hot path instantiates on the state's DataTypes, registered for the B200 types with the
// reference's own (host) implementation -- vector_device keeps the host copy coherent, so they work unchanged: BoxROI,
// TetrahedronSetGeometryAlgorithms (IdentityMapping: B200IdentityMapping.cpp) (the set SofaCUDA registers for the same reason:
// applications/plugins/SofaCUDA/Component/src/SofaCUDA/component/init.cpp:205,218,250).
#include <sofa/component/engine/select/BoxROI.inl>
#include <sofa/component/topology/container/dynamic/TetrahedronSetGeometryAlgorithms.inl>
#include <sofa/core/ObjectFactory.h>

#include "B200Types.h"

namespace sofa::component::engine::select::boxroi {
template class BoxROI<sofa::b200::B200Vec3fTypes>;
template class BoxROI<sofa::b200::B200Vec3dTypes>;
}  // namespace sofa::component::engine::select::boxroi
namespace sofa::component::topology::container::dynamic {
template class TetrahedronSetGeometryAlgorithms<sofa::b200::B200Vec3fTypes>;
template class TetrahedronSetGeometryAlgorithms<sofa::b200::B200Vec3dTypes>;
}  // namespace sofa::component::topology::container::dynamic

namespace sofa::b200 {
void registerEngines(sofa::core::ObjectFactory* factory) {
    typedef sofa::component::engine::select::boxroi::BoxROI<B200Vec3fTypes> BoxROIf;
    typedef sofa::component::engine::select::boxroi::BoxROI<B200Vec3dTypes> BoxROId;
    typedef sofa::component::topology::container::dynamic::TetrahedronSetGeometryAlgorithms<B200Vec3fTypes> TetraAlgoF;
    typedef sofa::component::topology::container::dynamic::TetrahedronSetGeometryAlgorithms<B200Vec3dTypes> TetraAlgoD;
    factory->registerObjects(sofa::core::ObjectRegistrationData("BoxROI on B200-typed positions (host implementation)").add<BoxROIf>().add<BoxROId>());
    factory->registerObjects(sofa::core::ObjectRegistrationData("TetrahedronSetGeometryAlgorithms on B200-typed positions (host implementation)").add<TetraAlgoF>().add<TetraAlgoD>());
}
}  // namespace sofa::b200
