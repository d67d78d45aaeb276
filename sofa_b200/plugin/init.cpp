// SOFA plugin entry points (dlopen'ed by PluginManager: Sofa/framework/Helper/src/sofa/helper/system/PluginManager.cpp:88-94;
// registerObjects looked up by ObjectFactory::registerObjectsFromPlugin, Sofa/framework/Core/src/sofa/core/ObjectFactory.cpp:756-798).
#include <sofa/core/ObjectFactory.h>
#include <sofa/helper/system/PluginManager.h>

#include "B200Types.h"

namespace sofa::b200 {
void registerMechanicalObject(sofa::core::ObjectFactory*);
void registerTetrahedronFEMForceField(sofa::core::ObjectFactory*);
void registerHexahedronFEMForceField(sofa::core::ObjectFactory*);
void registerTetrahedralCorotationalFEMForceField(sofa::core::ObjectFactory*);
void registerFastTetrahedralCorotationalForceField(sofa::core::ObjectFactory*);
void registerMeshMatrixMass(sofa::core::ObjectFactory*);
void registerDiagonalMass(sofa::core::ObjectFactory*);
void registerUniformMass(sofa::core::ObjectFactory*);
void registerPlaneForceField(sofa::core::ObjectFactory*);
void registerEngines(sofa::core::ObjectFactory*);
void registerIdentityMapping(sofa::core::ObjectFactory*);
void registerFixedProjectiveConstraint(sofa::core::ObjectFactory*);
void registerCGLinearSolver(sofa::core::ObjectFactory*);
}  // namespace sofa::b200

extern "C" {
SOFA_EXPORT_DYNAMIC_LIBRARY void initExternalModule() {
    static bool first = true;
    if (first) { sofa::helper::system::PluginManager::getInstance().registerPlugin("SofaB200"); first = false; }
}
SOFA_EXPORT_DYNAMIC_LIBRARY const char* getModuleName() { return "SofaB200"; }
SOFA_EXPORT_DYNAMIC_LIBRARY const char* getModuleVersion() { return sofab200_version(); }
SOFA_EXPORT_DYNAMIC_LIBRARY const char* getModuleLicense() { return "LGPL"; }
SOFA_EXPORT_DYNAMIC_LIBRARY const char* getModuleDescription() { return "B200-native corotational FEM + CG hot path (templates B200Vec3f, B200Vec3d)"; }
SOFA_EXPORT_DYNAMIC_LIBRARY void registerObjects(sofa::core::ObjectFactory* factory) {
    sofa::b200::registerMechanicalObject(factory);
    sofa::b200::registerTetrahedronFEMForceField(factory);
    sofa::b200::registerHexahedronFEMForceField(factory);
    sofa::b200::registerTetrahedralCorotationalFEMForceField(factory);
    sofa::b200::registerFastTetrahedralCorotationalForceField(factory);
    sofa::b200::registerMeshMatrixMass(factory);
    sofa::b200::registerDiagonalMass(factory);
    sofa::b200::registerUniformMass(factory);
    sofa::b200::registerPlaneForceField(factory);
    sofa::b200::registerEngines(factory);
    sofa::b200::registerIdentityMapping(factory);
    sofa::b200::registerFixedProjectiveConstraint(factory);
    sofa::b200::registerCGLinearSolver(factory);
}
}
