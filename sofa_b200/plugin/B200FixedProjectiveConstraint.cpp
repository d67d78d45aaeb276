// SOFA-side glue: FixedProjectiveConstraint<B200Vec3Types>::projectResponse / projectVelocity -> sofab200_fixed_project_response.
// The indices (a host-side TopologySubsetIndices) are mirrored into a device vector kept in the InternalData the reference class reserves for
// this (FixedProjectiveConstraint.h:41-44,88-89), refreshed when the Data's counter moves.
#include <sofa/component/constraint/projective/FixedProjectiveConstraint.inl>
#include <sofa/core/ObjectFactory.h>

#include "B200Types.h"

namespace sofa::component::constraint::projective {
using sofa::b200::B200Vec3Types;

template <class TReal> class FixedProjectiveConstraintInternalData<B200Vec3Types<TReal>> {
public:
    sofa::b200::B200Vector<unsigned int> indicesDevice;
    int counter = -1;
    /// device pointer of the indices of `m` (uploaded when d_indices changed since the last call)
    const uint32_t* indices(const FixedProjectiveConstraint<B200Vec3Types<TReal>>* m) {
        if (counter != m->d_indices.getCounter()) {
            const auto& idx = m->d_indices.getValue();
            indicesDevice.resize(idx.size());
            std::copy(idx.begin(), idx.end(), indicesDevice.hostWrite());
            counter = m->d_indices.getCounter();
        }
        return static_cast<const uint32_t*>(sofa::b200::devRead(indicesDevice));
    }
};

#define B200_FIXED(TReal)                                                                                                            \
    template <> void FixedProjectiveConstraint<B200Vec3Types<TReal>>::projectResponse(const core::MechanicalParams*, DataVecDeriv& resData) { \
        if (!data) data.reset(new FixedProjectiveConstraintInternalData<B200Vec3Types<TReal>>());    /* .inl:183-206 */              \
        auto& res = *resData.beginEdit();                                                                                            \
        if (sofab200_fixed_project_response(sofa::b200::threadContext(), B200Vec3Types<TReal>::abiReal, res.size(), sofa::b200::devWrite(res), \
                                            d_indices.getValue().size(), data->indices(this), d_fixAll.getValue() ? 1 : 0) != SOFAB200_OK) \
            msg_error() << sofab200_last_error();                                                                                    \
        resData.endEdit();                                                                                                           \
    }                                                                                                                                \
    template <> void FixedProjectiveConstraint<B200Vec3Types<TReal>>::projectVelocity(const core::MechanicalParams* mparams, DataVecDeriv& vData) { \
        if (d_projectVelocity.getValue()) projectResponse(mparams, vData);        /* .inl:228-252: the same rows are zeroed */       \
    }
B200_FIXED(float)
B200_FIXED(double)
template class FixedProjectiveConstraint<sofa::b200::B200Vec3fTypes>;
template class FixedProjectiveConstraint<sofa::b200::B200Vec3dTypes>;
}  // namespace sofa::component::constraint::projective

namespace sofa::b200 {
void registerFixedProjectiveConstraint(sofa::core::ObjectFactory* factory) {
    using namespace sofa::component::constraint::projective;
    factory->registerObjects(sofa::core::ObjectRegistrationData("FixedProjectiveConstraint on a B200 GPU (sofa_b200)")
                                 .add<FixedProjectiveConstraint<B200Vec3fTypes>>().add<FixedProjectiveConstraint<B200Vec3dTypes>>());
}
}  // namespace sofa::b200
