// SOFA-side glue: IdentityMapping from a B200-typed state to a host-typed one (visual / collision models stay on the host), with the reference's
// own implementation (SofaCUDA registers the same pairs: applications/plugins/SofaCUDA/Component/src/SofaCUDA/component/init.cpp:218).
// NEEDS_EIGEN: IdentityMapping.inl pulls in <Eigen/Sparse> through EigenSparseMatrix.h, which this image does not have, so
// tools/plugin_syntax_check.sh skips this one file (and says so); everything it uses from the glue (B200Vec3Types, the accessors) is
// exercised by the other files.
#include <sofa/component/mapping/linear/IdentityMapping.inl>
#include <sofa/core/ObjectFactory.h>

#include "B200Types.h"

namespace sofa::component::mapping::linear {
template class IdentityMapping<sofa::b200::B200Vec3fTypes, sofa::defaulttype::Vec3Types>;
template class IdentityMapping<sofa::b200::B200Vec3dTypes, sofa::defaulttype::Vec3Types>;
}  // namespace sofa::component::mapping::linear

namespace sofa::b200 {
void registerIdentityMapping(sofa::core::ObjectFactory* factory) {
    using namespace sofa::component::mapping::linear;
    factory->registerObjects(sofa::core::ObjectRegistrationData("IdentityMapping from a B200-typed state to a host-typed one")
                                 .add<IdentityMapping<B200Vec3fTypes, sofa::defaulttype::Vec3Types>>()
                                 .add<IdentityMapping<B200Vec3dTypes, sofa::defaulttype::Vec3Types>>());
}
}  // namespace sofa::b200
