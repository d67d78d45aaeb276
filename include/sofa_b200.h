/*
 * sofa_b200.h -- C ABI of the B200-native implicit-dynamics FEM hot path for SOFA.
 *
 * This is the drop-in boundary: plain C, opaque handles, raw pointers and sizes, int status
 * returns.  A SOFA plugin (sofa_b200/plugin/) specialises the reference's
 * component templates for a device DataTypes and forwards each virtual to one entry point below;
 * the Python/ctypes host layer used by tests/ and bench.py binds exactly the same symbols.
 *
 * Conventions
 *   - `real`: SOFAB200_F32 = Vec3f build of the reference, SOFAB200_F64 = Vec3d.  State vectors are
 *     AoS Vec3 of that Real (x0 y0 z0 x1 ...), the memory layout of sofa::type::vector<Vec<3,Real>>.
 *   - pointers named *_dev are CUDA device pointers valid on the context's device; *_host are host
 *     pointers.  Indices are uint32_t (sofa::Index).
 *   - every call enqueues on the context's stream and returns without synchronising unless the
 *     comment says "(sync)".  No entry point ever falls back to a CPU computation.
 *   - return value: SOFAB200_OK or a negative error; sofab200_last_error() gives the text
 *     (the plugin maps non-zero to msg_error() + ComponentState::Invalid, the reference's own
 *     failure convention: TetrahedronFEMForceField.inl:1299-1311).
 *
 * Reference interfaces replaced (paths relative to the SOFA tree):
 *   [TFF]  Sofa/Component/SolidMechanics/FEM/Elastic/src/sofa/component/solidmechanics/fem/elastic/TetrahedronFEMForceField.{h,inl}
 *   [HFF]  .../elastic/HexahedronFEMForceField.{h,inl}
 *   [MO]   Sofa/Component/StateContainer/src/sofa/component/statecontainer/MechanicalObject.inl
 *   [DM]   Sofa/Component/Mass/src/sofa/component/mass/DiagonalMass.inl
 *   [FPC]  Sofa/Component/Constraint/Projective/src/sofa/component/constraint/projective/FixedProjectiveConstraint.inl
 *   [CG]   Sofa/Component/LinearSolver/Iterative/src/sofa/component/linearsolver/iterative/CGLinearSolver.inl
 *   [GS]   .../iterative/GraphScatteredTypes.cpp
 *   [EI]   Sofa/Component/ODESolver/Backward/src/sofa/component/odesolver/backward/EulerImplicitSolver.cpp
 *   [CUDA] applications/plugins/SofaCUDA/Component/src/SofaCUDA/component/solidmechanics/fem/elastic/CudaTetrahedronFEMForceField.inl:31-55
 *          (the extern "C" kernel-launcher layer of the incumbent GPU plugin, whose role this header takes)
 */
#ifndef SOFA_B200_H
#define SOFA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOFAB200_OK 0
#define SOFAB200_ERR_INVALID (-1)  /* bad argument / inconsistent sizes                        */
#define SOFAB200_ERR_CUDA (-2)     /* a CUDA runtime call failed (text in sofab200_last_error) */
#define SOFAB200_ERR_NO_DEVICE (-3)/* no CUDA device: this library never computes on the CPU   */
#define SOFAB200_ERR_UNSUPPORTED (-4)

typedef enum { SOFAB200_F32 = 0, SOFAB200_F64 = 1 } sofab200_real;

/* TetrahedronFEMForceField `method` Data ([TFF].h:75-77, setMethod .inl) */
typedef enum { SOFAB200_TET_SMALL = 0, SOFAB200_TET_LARGE = 1, SOFAB200_TET_POLAR = 2, SOFAB200_TET_SVD = 3,
               SOFAB200_TET_POLAR2 = 4 /* FastTetrahedralCorotationalForceField only ("polar2") */ } sofab200_tet_method;
/* HexahedronFEMForceField method numbering ([HFF].h setMethod: 0 large, 1 polar, 2 small) */
typedef enum { SOFAB200_HEX_LARGE = 0, SOFAB200_HEX_POLAR = 1, SOFAB200_HEX_SMALL = 2 } sofab200_hex_method;

typedef struct sofab200_ctx sofab200_ctx;       /* one device + one stream; one per force-field owner thread */
typedef struct sofab200_tetfem sofab200_tetfem; /* TetrahedronFEMForceFieldInternalData<B200Types>            */
typedef struct sofab200_hexfem sofab200_hexfem; /* HexahedronFEMForceFieldInternalData<B200Types>             */
typedef struct sofab200_node sofab200_node;     /* one solver node kept resident on the device                */

/* ------------------------------------------------------------------------------------------------ */
/* library / context                                                                                */
/* ------------------------------------------------------------------------------------------------ */
const char* sofab200_version(void);
/* Text of the last error raised on the calling thread ("" when none). */
const char* sofab200_last_error(void);
/* device: CUDA ordinal.  cuda_stream: a cudaStream_t to enqueue on (pass cudaStreamLegacy / cudaStreamPerThread
 * explicitly for the default streams), or NULL to create an own
 * non-blocking stream.  Replaces mycudaInit() (SofaCUDA/Core/src/sofa/gpu/cuda/mycuda.cu:142-155). */
int sofab200_ctx_create(int device, void* cuda_stream, sofab200_ctx** out);
int sofab200_ctx_destroy(sofab200_ctx* ctx);
int sofab200_ctx_set_stream(sofab200_ctx* ctx, void* cuda_stream);
int sofab200_ctx_synchronize(sofab200_ctx* ctx); /* (sync) */
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t sofab200_ctx_launch_count(const sofab200_ctx* ctx);
/* Per-kernel-class device timing with CUDA events recorded on the context's stream around every launch
 * (replaces SofaCUDA's CUDA_TIMER_SYNC env switch, mycuda.cpp:146).  Classes:
 *   0 element pass of addDForce / A*p   1 boundary gather   2 element pass of addForce   3 CG / vector kernels
 *   4 persistent CG kernel (the whole CGLinearSolver loop: one launch per solve)
 * profile_begin enables recording; profile_end (sync) disables it and returns, per class, the summed
 * milliseconds and the number of launches.  total_ms / count: arrays of SOFAB200_PROFILE_CLASSES. */
#define SOFAB200_PROFILE_CLASSES 5
int sofab200_ctx_profile_begin(sofab200_ctx* ctx);
int sofab200_ctx_profile_end(sofab200_ctx* ctx, double* total_ms, uint64_t* count);
/* In-kernel phase timestamps (diagnostics; no reference counterpart).  After trace_begin every tile CTA of the solver
 * node's element passes and every CTA of the fused CG tail records %globaltimer (ns) at its phase boundaries, 16 words per
 * CTA (word 7 = SM id); later launches overwrite earlier ones.  trace_end (sync) copies the first n words to `out` and
 * turns tracing off.  Words [0, 4096*16): element pass; [4096*16, 2*4096*16): CG loop. */
int sofab200_ctx_trace_begin(sofab200_ctx* ctx);
int sofab200_ctx_trace_end(sofab200_ctx* ctx, uint64_t* out, size_t n);

/* PlaneForceField<B200Vec3Types> (MechanicalLoad/.../PlaneForceField.inl; in every SofaCUDA FEM benchmark scene), per-operation
 * level: addForce :158-205 (penalty + damping below the plane, optional maxForce clamp, records the contacts), addDForce
 * :208-226 (df[p] += n * (fact * (dx[p].n)) on the recorded contacts, fact = Real(-stiffness * k_factor)).  normal / d are the Data
 * values BEFORE setPlane's normalisation (:139-145), which is applied here in Real.  contacts_dev: n bytes, written by add_force. */
typedef struct sofab200_plane_desc { double normal[3]; double d; double stiffness, damping, max_force; int bilateral; } sofab200_plane_desc;
int sofab200_plane_add_force(sofab200_ctx* ctx, sofab200_real real, size_t n, void* f_dev, const void* x_dev, const void* v_dev,
                             const sofab200_plane_desc* plane, unsigned char* contacts_dev);
int sofab200_plane_add_dforce(sofab200_ctx* ctx, sofab200_real real, size_t n, void* df_dev, const void* dx_dev,
                              const sofab200_plane_desc* plane, const unsigned char* contacts_dev, double k_factor);

/* ------------------------------------------------------------------------------------------------ */
/* MechanicalObject<B200Vec3Types> vector operations  -- [MO]                                       */
/* ------------------------------------------------------------------------------------------------ */
/* MechanicalObject::vOp(r, a, b, k) [MO]:2075-2203.  a_dev / b_dev NULL == null VecId:
 *   a,b null: r = 0 | a null, b==r: r *= k | a null: r = b*k | b null: r = a | r==a: r += b*k
 *   | r==b: r = a + r*k | else r = a + b*k   (k == 1 takes the reference's multiplication-free forms). */
int sofab200_mo_vop(sofab200_ctx* ctx, sofab200_real real, size_t n, void* r_dev, const void* a_dev, const void* b_dev, double k);
/* MechanicalObject::vDot [MO]:2333-2356.  Accumulated in double with a fixed-order tree (run-to-run
 * reproducible); the reference's serial Real accumulation order cannot be kept in parallel. (sync) */
int sofab200_mo_vdot(sofab200_ctx* ctx, sofab200_real real, size_t n, const void* a_dev, const void* b_dev, double* result_host);
/* The same dot product restricted to the nodes whose mask byte is non-zero (NULL = all) with the result left in device
 * memory (no synchronisation): the per-rank term of a distributed vDot (owned nodes only) before its allreduce. */
int sofab200_mo_vdot_dev(sofab200_ctx* ctx, sofab200_real real, size_t n, const void* a_dev, const void* b_dev, const unsigned char* node_mask_dev, double* result_dev);
/* MechanicalObject::vMultiOp integration fast path [MO]:2208-2241: v += a*f_v_a ; x += v*f_x_v. */
int sofab200_mo_vmultiop_integrate(sofab200_ctx* ctx, sofab200_real real, size_t n, void* v_dev, void* x_dev, const void* a_dev, double f_v_a, double f_x_v);

/* MechanicalObject::accumulateForce [MO]:1356-1375: f[i] += externalForce[i] for every row of externalForce that differs from Deriv(). */
int sofab200_mo_accumulate_force(sofab200_ctx* ctx, sofab200_real real, size_t n, void* f_dev, const void* ext_dev);

/* ------------------------------------------------------------------------------------------------ */
/* DiagonalMass / FixedProjectiveConstraint pieces of A*p and of the right-hand side                  */
/* ------------------------------------------------------------------------------------------------ */
/* DiagonalMass::addMDx [DM]:535-559: res[i] += (dx[i]*m[i])*factor  (factor==1: res[i] += dx[i]*m[i]). */
int sofab200_mass_add_mdx(sofab200_ctx* ctx, sofab200_real real, size_t n, void* res_dev, const void* dx_dev, const void* vertex_mass_dev, double factor);
/* DiagonalMass::addForce [DM]:1392-1413: f[i] += gravity*m[i]. */
int sofab200_mass_add_force(sofab200_ctx* ctx, sofab200_real real, size_t n, void* f_dev, const void* vertex_mass_dev, const double gravity[3]);
/* DiagonalMass::accFromF [DM]:563-575: a[i] = f[i]/m[i]. */
int sofab200_mass_acc_from_f(sofab200_ctx* ctx, sofab200_real real, size_t n, void* a_dev, const void* f_dev, const void* vertex_mass_dev);
/* UniformMass<DataTypes>::addMDx (Mass/.../UniformMass.inl:403-420): m = vertexMass (*= Real(factor) if factor != 1); res[i] += dx[i]*m
 * and ::addForce (:469-496): mg = gravity*m once; f[i] += mg. */
int sofab200_uniform_mass_add_mdx(sofab200_ctx* ctx, sofab200_real real, size_t n, void* res_dev, const void* dx_dev, double vertex_mass, double factor);
int sofab200_uniform_mass_add_force(sofab200_ctx* ctx, sofab200_real real, size_t n, void* f_dev, double vertex_mass, const double gravity[3]);
/* MeshMatrixMass<DataTypes> (Sofa/Component/Mass/src/sofa/component/mass/MeshMatrixMass.inl) -- [MMM].  The component's init() stays on
 * the host (d_vertexMass, d_edgeMass, m_massLumpingCoeff as the CPU class computes them, l_topology->getEdges() in topology order); the
 * device applies the matrix.  Every node gathers its half-edges in ascending edge index after the vertex term -- the order the
 * reference's sequential scatter adds them in -- so the result is bit-identical and reproducible (no atomics). (create: sync) */
typedef struct sofab200_meshmass sofab200_meshmass;
int sofab200_meshmass_create(sofab200_ctx* ctx, sofab200_real real, size_t n_nodes, const void* vertex_mass_host, size_t n_edges,
                             const uint32_t* edges_host /* n_edges x 2 */, const void* edge_mass_host, int lumping, double mass_lumping_coeff,
                             sofab200_meshmass** out);
int sofab200_meshmass_destroy(sofab200_meshmass* mm);
/* addMDx [MMM]:1987-2048 */
int sofab200_meshmass_add_mdx(sofab200_meshmass* mm, void* res_dev, const void* dx_dev, double factor);
/* addForce [MMM]:2072-2092: f[i] += gravity*vertexMass[i]*m_massLumpingCoeff */
int sofab200_meshmass_add_force(sofab200_meshmass* mm, void* f_dev, const double gravity[3]);
/* accFromF [MMM]:2050-2069: lumped only; SOFAB200_ERR_UNSUPPORTED otherwise (the reference prints an error and returns) */
int sofab200_meshmass_acc_from_f(sofab200_meshmass* mm, void* a_dev, const void* f_dev);
/* FixedProjectiveConstraint::projectResponse / projectVelocity [FPC]:183-206,236-258. */
int sofab200_fixed_project_response(sofab200_ctx* ctx, sofab200_real real, size_t n, void* res_dev, size_t n_indices, const uint32_t* indices_dev, int fix_all);

/* ------------------------------------------------------------------------------------------------ */
/* TetrahedronFEMForceField<B200Vec3Types>  -- [TFF]                                                 */
/* ------------------------------------------------------------------------------------------------ */
typedef struct sofab200_tetfem_desc {
    int method;                 /* sofab200_tet_method  (Data `method`)                                    */
    size_t n_young;             /* Data `youngModulus`: one value, or one per element                      */
    const double* young;
    size_t n_poisson;           /* Data `poissonRatio`: idem                                               */
    const double* poisson;
    size_t n_local_stiffness;   /* Data `localStiffnessFactor` (may be 0)                                  */
    const double* local_stiffness;
    int tile_elems;             /* elements per CTA tile of the device layout; 0 = library default         */
    const unsigned char* shared_nodes; /* optional, n_nodes flags: nodes whose contributions must take the staging path
                                 * (partition-interface nodes of a multi-GPU run, see sofab200_node_set_peer); NULL = none */
    double plastic_max_threshold;   /* Data `plasticMaxThreshold` (2-norm of the strain); <= 0 = no plasticity (the default)   */
    double plastic_yield_threshold; /* Data `plasticYieldThreshold` (reference default 0.0001)                                  */
    double plastic_creep;           /* Data `plasticCreep` (reference default 0.9)                                              */
    int update_stiffness_matrix;    /* Data `updateStiffnessMatrix`: addForce recomputes the strain-displacement terms from the deformed element every
                                     * step ([TFF].inl:1063-1067,1174-1177).  With `large` the reference rewrites 9 single entries of J, all in its
                                     * normal-strain columns ([TFF].inl:908-922): the element then carries a second set of 12 cofactors for the shear
                                     * columns, and the CG loop runs as separate kernels (not the persistent one).  Ignored by `small`. */
    int tetrahedral_corotational;   /* 1: the component is a TetrahedralCorotationalFEMForceField (…/fem/elastic/TetrahedralCorotationalFEMForceField.inl,
                                     * what Demos/liver.scn uses): its init, accumulateForce{Small,Large,Polar}, applyStiffness* and computeForce
                                     * (:356-398,400-602,604-743,840-1044,1046-1175) are statement for statement those of TetrahedronFEMForceField, so the
                                     * same kernels serve it.  Differences: no svd, no plasticity; updateStiffnessMatrix works with `large` too (there all
                                     * three copies of a cofactor are rewritten together, :920-937); sofab200_tetfem_get_rotations follows the class's own getRotation
                                     * (:779-820: mean of rotation * initialTransformation, Gram-Schmidt instead of a polar decomposition; large / polar);
                                     * its own computeVonMisesStress is not provided. */
    int compute_von_mises;          /* Data `computeVonMisesStress` (0 = off, 1 = corotational strain, 2 = Green-Lagrange strain):
                                     * non-zero makes init keep the shape-function matrices and Lame coefficients ([TFF].inl:278-282,1521-1541) */
    int fast_corotational;          /* 1: the component is a FastTetrahedralCorotationalForceField (…/fem/elastic/FastTetrahedralCorotationalForceField.inl
                                     * -- [FTC]): per tetrahedron six 3x3 edge blocks of the linear stiffness and a rotation ([FTC].inl:38-150, addForce
                                     * :296-399); addDForce runs over the EDGES with one 3x3 matrix per edge, re-assembled from the tetrahedra after every
                                     * addForce ([FTC].inl:402-470).  method: SOFAB200_TET_LARGE = "qr"/"large" (the class's default), _POLAR = "polar",
                                     * _POLAR2 = "polar2", _SMALL = "none"/"linear"/"small".  No plasticity / von Mises / updateStiffnessMatrix /
                                     * localStiffnessFactor Data; get_rotations / compute_von_mises / reset are not available; single GPU. */
    size_t n_edges;                 /* fast_corotational: the topology's edge list (n_edges x 2) when it holds one; 0 = number the edges as            */
    const uint32_t* edges;          /* TetrahedronSetTopologyContainer::createEdgesInTetrahedronArray does (first appearance, vertices sorted)        */
} sofab200_tetfem_desc;

/* init()+reinit() [TFF].inl:1257-1545: per-element material stiffness, rest rotation, rotated rest
 * shape and strain-displacement terms are computed in the reference's arithmetic, then laid out in
 * HBM as CTA tiles with a deterministic gather plan.  rest_position_host: n_nodes Vec3 of `real`
 * (the MechanicalObject's restPosition); tets_host: n_tets x 4 indices in topology order. (sync) */
int sofab200_tetfem_create(sofab200_ctx* ctx, sofab200_real real, size_t n_nodes, const void* rest_position_host,
                           size_t n_tets, const uint32_t* tets_host, const sofab200_tetfem_desc* desc, sofab200_tetfem** out);
int sofab200_tetfem_destroy(sofab200_tetfem* ff);
/* addForce(mparams, f, x, v) [TFF].inl:1548-1604: f += elastic forces at x; caches rotations[e].  With plastic_max_threshold > 0
 * the plasticity branch of computeForce ([TFF].inl:357-371) runs and updates the per-element plastic strain. */
int sofab200_tetfem_add_force(sofab200_tetfem* ff, void* f_dev, const void* x_dev);
/* addDForce(mparams, df, dx) [TFF].inl:1606-1636 with k_factor = kFactorIncludingRayleighDamping
 * (MechanicalParams.h:62): df -= R K_e R^T dx * k_factor using the cached rotations. */
int sofab200_tetfem_add_dforce(sofab200_tetfem* ff, void* df_dev, const void* dx_dev, double k_factor);
/* Element-ordered copies for inspection / parity (sync).  what:
 *   "rotations" (T x 9, rotations[e] = R^T, [TFF].inl:880), "initialRotations" (T x 9),
 *   "strainDisplacements" (T x 12: the 12 distinct cofactors), "materialsStiffnesses" (T x 3: K00,K01,K33),
 *   "rotatedInitialElements" (T x 12), "initialTransformation" (T x 9, svd only),
 *   "plasticStrains" (T x 6, _plasticStrains; only with plastic_max_threshold > 0).
 * fast_corotational: "rotations" (T x 9, tetraInfo.rotation of the last addForce), "restRotations" (T x 9), "shapeVectors" (T x 4 x 3),
 *   "linearDfDx" (T x 6 x 9), "linearDfDxDiag" (T x 4 x 9), "restEdgeVectors" (T x 6 x 3), "edgeOrientations" (T x 6),
 *   "edgeInfo" (E x 9, d_edgeInfo as of the last addDForce), "edges" (E x 2 uint32), "n_edges" (one uint64). */
int sofab200_tetfem_get(sofab200_tetfem* ff, const char* what, void* out_host);
/* computeVonMisesStress() [TFF].inl:2196-2416 at positions x (the reference runs it on AnimateEndEvent): von Mises stress per element
 * (d_vonMisesPerElement, topology order) and per node (d_vonMisesPerNode: mean over the tetrahedra around the node, ascending index).
 * Method 1 also rewrites rotations[e] from x, as the reference does.  Either output may be NULL.  Needs compute_von_mises != 0 at
 * creation (the value passed there selects the method).  The colour map of the reference is display code and is not computed. */
int sofab200_tetfem_compute_von_mises(sofab200_tetfem* ff, const void* x_dev, void* per_element_dev, void* per_node_dev);
/* reset() [TFF].inl:1380-1388: clears the plastic strains (nothing else is reset by the reference). */
int sofab200_tetfem_reset(sofab200_tetfem* ff);
/* getRotations(VecReal& vecR) [TFF].inl:781-833,2033-2042 (what WarpPreconditioner / RotationMatrix consumers read; SofaCUDA:
 * CudaTetrahedronFEMForceField.inl getRotations): per node the mean of rotations[t] * R0(t) over the tetrahedra around it, in
 * ascending tetrahedron index, made orthogonal by polarDecomposition; identity for method small.  vecR_dev: n_nodes x 9 `real`
 * (row-major 3x3), device memory. */
int sofab200_tetfem_get_rotations(sofab200_tetfem* ff, void* vecR_dev);
/* Layout statistics: out[0]=tiles, [1]=elements per tile, [2]=interior nodes, [3]=shared nodes,
 * [4]=staged (HBM) corner contributions, [5]=dynamic shared memory bytes, [6]=max valence, [7]=n_tets */
int sofab200_tetfem_stats(const sofab200_tetfem* ff, uint64_t out[8]);

/* ------------------------------------------------------------------------------------------------ */
/* HexahedronFEMForceField<B200Vec3Types>  -- [HFF]                                                  */
/* ------------------------------------------------------------------------------------------------ */
typedef struct sofab200_hexfem_desc {
    int method;                 /* sofab200_hex_method */
    size_t n_young;
    const double* young;
    size_t n_poisson;
    const double* poisson;
    int tile_elems;
} sofab200_hexfem_desc;
int sofab200_hexfem_create(sofab200_ctx* ctx, sofab200_real real, size_t n_nodes, const void* rest_position_host,
                           size_t n_hexas, const uint32_t* hexas_host, const sofab200_hexfem_desc* desc, sofab200_hexfem** out);
int sofab200_hexfem_destroy(sofab200_hexfem* ff);
/* addForce [HFF].inl:194-246 / addDForce [HFF].inl:248-286 */
int sofab200_hexfem_add_force(sofab200_hexfem* ff, void* f_dev, const void* x_dev);
int sofab200_hexfem_add_dforce(sofab200_hexfem* ff, void* df_dev, const void* dx_dev, double k_factor);
/* what: "rotations" (H x 9, _rotations[e] = R), "elementStiffnesses" (H x 576), "rotatedInitialElements" (H x 24) */
int sofab200_hexfem_get(sofab200_hexfem* ff, const char* what, void* out_host);
/* getRotations [HFF].inl:946-1023: getNodeRotation for every node -- identity + sum of _rotations[h] * _initialrotations[h]^T over the
 * hexahedra around the node (ascending index), divided by their number, polar-decomposed (the reference starts the sum from the identity;
 * reproduced as is).  vecR_dev: n_nodes x 9 `real`, row-major, device memory. */
int sofab200_hexfem_get_rotations(sofab200_hexfem* ff, void* vecR_dev);
/* as sofab200_tetfem_stats, except out[6] = number of unique (bit-identical) element stiffness matrices stored */
int sofab200_hexfem_stats(const sofab200_hexfem* ff, uint64_t out[8]);

/* ------------------------------------------------------------------------------------------------ */
/* One solver node resident on the device: EulerImplicitSolver + CGLinearSolver<GraphScattered> over  */
/* {MechanicalObject, DiagonalMass, one FEM force field, FixedProjectiveConstraint}  -- [EI][CG][GS]  */
/* ------------------------------------------------------------------------------------------------ */
typedef struct sofab200_node_desc {
    sofab200_tetfem* tetfem;        /* exactly one of tetfem / hexfem                                     */
    sofab200_hexfem* hexfem;
    const void* vertex_mass_host;   /* DiagonalMass `vertexMass` (n Reals) or NULL for no mass            */
    size_t n_fixed;                 /* FixedProjectiveConstraint `indices`                                */
    const uint32_t* fixed_host;
    int fix_all;                    /* Data `fixAll`                                                      */
    int mass_first;                 /* 1: the mass precedes the force field in the scene (all reference scenes) */
    int uniform_mass;               /* 1: the mass is a UniformMass (Mass/.../UniformMass.inl:403-496) with the MassType below;  */
    double uniform_vertex_mass;     /*    vertex_mass_host is then ignored                                       */
    const sofab200_plane_desc* plane; /* optional PlaneForceField, the node's LAST force field (as in the SofaCUDA benchmark scenes): its
                                     * addForce / addDForce are fused into the per-node epilogue of the element passes        */
    double plane_rayleigh_stiffness;
} sofab200_node_desc;

typedef struct sofab200_solver_params {
    double gravity[3];              /* context gravity                                                    */
    double dt;
    double rayleigh_stiffness;      /* EulerImplicitSolver Data [EI]:40-50                                */
    double rayleigh_mass;
    double vdamping;
    int first_order;
    int trapezoidal;
    unsigned iterations;            /* CGLinearSolver Data [CG]:35-45                                     */
    double tolerance;
    double threshold;
    int warm_start;
    double ff_rayleigh_stiffness;   /* BaseForceField::rayleighStiffness of the FEM component             */
    double mass_rayleigh_mass;      /* Mass::rayleighMass of the mass component                           */
} sofab200_solver_params;

int sofab200_node_create(sofab200_ctx* ctx, sofab200_real real, size_t n_nodes, const sofab200_node_desc* desc, sofab200_node** out);
int sofab200_node_destroy(sofab200_node* node);
int sofab200_node_set_params(sofab200_node* node, const sofab200_solver_params* p);
/* mop.computeForce: f = sum of addForce over the node's force fields in scene order
 * (MappingGraphMechanicalOperations.cpp:41-93): one fused pass, gravity term + element forces. */
int sofab200_node_compute_force(sofab200_node* node, void* f_dev, const void* x_dev);
/* GraphScatteredMatrix::apply [GS]:33-46: q = project((m M + b B + k K) p), fused in one pass. */
int sofab200_node_apply(sofab200_node* node, void* q_dev, const void* p_dev, double m_factor, double b_factor, double k_factor);
/* The visitor form of the same pass, MechanicalAddMBKdxVisitor / mop.addMBKv (MechanicalOperations.cpp:291-307,
 * MappingGraphMechanicalOperations.cpp:94-137): out = [init +] (m M + b B + k K) d, optionally scaled (b.teq(h)) and
 * projected.  init_dev may be NULL (start from 0) or alias out_dev. */
int sofab200_node_add_mbkdx(sofab200_node* node, void* out_dev, const void* init_dev, const void* d_dev, double m_factor, double b_factor, double k_factor,
                            int scale, double scale_factor, int project);
/* Replace the node's DiagonalMass vertexMass (n Reals, host).  Used by the multi-GPU host layer to give every rank the
 * global lumped mass of its nodes, zeroed where another rank adds the mass term of a shared node. */
int sofab200_node_set_vertex_mass(sofab200_node* node, const void* vertex_mass_host);
/* The node's mass component is a MeshMatrixMass (same real type and context; it must precede the force field in the scene): its addForce /
 * addMDx terms run as their own kernels and become the per-node start values of the element passes, so f, b and A*p keep the reference's
 * order of operations.  The CG loop of such a node is the multi-kernel one (the persistent kernel has no neighbour access to p).
 * Not available in a distributed node. */
int sofab200_node_set_mesh_mass(sofab200_node* node, sofab200_meshmass* mesh_mass);
/* CGLinearSolver::solve [CG]:73-315 for the matrix-free system (m M + b B + k K), entirely on the
 * device: no host round trip per iteration.  x_dev: solution (initial guess when warm_start).
 * nb_iter_host: NULL = leave the result on the device (async); else (sync) receives "CG iterations". */
int sofab200_node_cg_solve(sofab200_node* node, void* x_dev, const void* b_dev, double m_factor, double b_factor, double k_factor, int* nb_iter_host);
/* EulerImplicitSolver::solve [EI]:83-341 on device-resident x, v (async). */
int sofab200_node_step(sofab200_node* node, void* x_dev, void* v_dev);
/* The same step for HOST state vectors (pinned or pageable): H2D of x,v, step, D2H of x,v; the upload of v overlaps addForce,
 * which only needs x. (sync) */
int sofab200_node_step_host(sofab200_node* node, void* x_host, void* v_host);
/* The step for a host-owned POSITION vector only: x goes up, the step runs on it and on the node's device-resident velocities, x comes back.
 * The velocities never cross the bus unless asked: v_host_in (may be NULL) is uploaded first (initial velocities, or a host that changed
 * them; they start at zero otherwise), v_host_out (may be NULL) receives the new ones.  This is what [EI]:83-341 needs per step when the state
 * lives in a device-typed MechanicalObject: only a reader on the host (visual model, collision) makes x come back. (sync) */
int sofab200_node_step_host_x(sofab200_node* node, void* x_host, const void* v_host_in, void* v_host_out);
/* MechanicalObject's `externalForce` vector (Data `externalForce`, added to the freshly reset force by accumulateForce,
 * Sofa/Component/StateContainer/src/sofa/component/statecontainer/MechanicalObject.inl:1356-1375, before any force field's addForce): n Vec3 of
 * `real` in host memory, uploaded now; NULL removes it.  On a distributed node every rank passes the rows of ITS nodes; only the owner's copy of an
 * interface node keeps its row (the force sums over the sharing ranks would count it once per rank otherwise). (sync) */
int sofab200_node_set_external_force(sofab200_node* node, const void* ext_host);
/* One step of a DEVICE-RESIDENT state (x_dev, v_dev as for sofab200_node_step) coupled to a host loop: ext_host (may be NULL; pinned memory for
 * the copy to be asynchronous) holds this step's external forces and is uploaded first; the new positions are copied into x_out_host on a second
 * stream while the caller already submits the next step.  The call returns as soon as the PREVIOUS call's positions are complete on the host and
 * this call's forces have been read from ext_host (which may then be refilled), so the caller alternates two output buffers; sofab200_node_flush waits for everything outstanding.  This is the coupling of a device-typed
 * MechanicalObject with host-side readers (visual model, collision) and writers (interaction forces): nothing of the state crosses the bus twice. */
int sofab200_node_step_pipelined(sofab200_node* node, void* x_dev, void* v_dev, const void* ext_host, void* x_out_host);
int sofab200_node_flush(sofab200_node* node);
/* Which kernel the last CG solve of the node ran (diagnostics for the benchmark line): out = {grid, tiles per CTA, 1 = tile state cached in
 * shared memory / 0 = streamed from HBM scratch, dynamic shared memory bytes, element threads, dedicated shared-node threads,
 * 1 = fused single-reduction kernel enabled, 1 = a persistent kernel is enabled at all}; the first six are zero before the first fused solve. */
int sofab200_node_cg_kernel_info(const sofab200_node* node, int out[8]);
/* Results of the last solve (sync): nb_iter ("CG iterations" as the reference reports it), end condition
 * (0 iterations exhausted, 1 tolerance, 2 threshold, 3 den==0, 4 b==0; 99 = multi-GPU only: a wait on another GPU timed out
 * inside the CG kernel and the solve was abandoned), and the `graph` Data
 * (Error / Denominator histories, [CG]:109-116,148,213).  Any pointer may be NULL. */
int sofab200_node_last_solve(sofab200_node* node, int* nb_iter, int* end_cond, double* graph_error, size_t* n_error, double* graph_den, size_t* n_den, size_t cap);
/* Device vectors of the last step for parity checks (sync): "f" (force), "b" (right-hand side), "dx" (solution): n Vec3 of Real;
 * "plane_contacts": n bytes, the PlaneForceField contact flags of the last addForce. */
int sofab200_node_get(sofab200_node* node, const char* what, void* out_host);
/* ------------------------------------------------------------------------------------------------ */
/* Multi-GPU: one process per GPU, the mesh partitioned by contiguous element ranges (sofa_b200/parallel.py).   */
/* The reference has no distributed mode; these entry points are new.                                           */
/* ------------------------------------------------------------------------------------------------ */
typedef struct sofab200_comm sofab200_comm;   /* an NCCL communicator bound to a context */
#define SOFAB200_UNIQUE_ID_BYTES 128
/* Rank 0 calls get_unique_id and ships the bytes to the other ranks (torch.distributed / MPI / a file);
 * then every rank calls comm_create (collective, blocking). */
int sofab200_comm_get_unique_id(void* out_bytes);
int sofab200_comm_create(sofab200_ctx* ctx, int world, int rank, const void* unique_id_bytes, sofab200_comm** out);
int sofab200_comm_destroy(sofab200_comm* comm);

/* Interface (halo) description of the rank-local node set, all arrays on the host:
 *   owned[n_nodes]            1 where this rank owns the node (lowest sharing rank), else 0
 *   interface[n_interface]    local ids of the nodes shared with other ranks, ascending
 *   my_slot[n_interface]      position of this rank in the node's ascending list of sharing ranks
 *   for neighbour k: nb_rank[k], nb_count[k] nodes; nb_rows[k][i] = row in `interface`, nb_slot[k][i] = the neighbour's
 *   position in that node's list.  Both sides list the shared nodes of a pair in the same (global id) order. */
typedef struct sofab200_halo_desc {
    const unsigned char* owned;
    size_t n_interface;
    const uint32_t* interface;
    const int32_t* my_slot;
    int max_sharers;
    int n_neighbours;
    const int* nb_rank;
    const size_t* nb_count;
    const uint32_t* const* nb_rows;
    const int32_t* const* nb_slot;
} sofab200_halo_desc;
/* Attach a communicator and a halo plan to a node: from then on sofab200_node_step / _cg_solve / _apply / _compute_force run
 * the distributed algorithm -- rank-local fused element pass, NCCL send/recv of the interface partial sums added in
 * ascending rank order on every sharing rank, dot products over owned nodes all-reduced -- still without any host round
 * trip per iteration and still replayed from one CUDA graph per step.  sofab200_node_set_peer (below) then replaces the
 * NCCL traffic by stores into peer memory from inside one persistent kernel per GPU.  The node's vertexMass must hold the global lumped
 * mass on owned nodes and 0 elsewhere (sofab200_node_set_vertex_mass). */
int sofab200_node_set_distributed(sofab200_node* node, sofab200_comm* comm, const sofab200_halo_desc* halo);

/* ---- peer memory: the multi-GPU CG loop in ONE persistent kernel per GPU, exchanging over NVLink ------------------------
 * Every rank allocates a mailbox (sofab200_peer_alloc: cudaMalloc + CUDA IPC handle), ships the 64-byte handle to the other
 * ranks of the node by any means (tests and bench use torch.distributed), maps theirs (sofab200_peer_open) and hands all the
 * mapped base pointers to its solver node (sofab200_node_set_peer, after sofab200_node_set_distributed).  From then on
 * sofab200_node_cg_solve / _step run the persistent CG kernel on every GPU: the partial sums of the interface nodes are stored
 * straight into the neighbours' mailboxes, the dot products are all-reduced through the mailboxes (every rank adds the ranks'
 * values in rank order), and the GPUs synchronise with sequence-numbered flags -- no NCCL call and no kernel boundary inside
 * the loop.  The force field must have been created with the interface nodes flagged in sofab200_tetfem_desc::shared_nodes.
 * Falls back to the NCCL loop when the mesh does not fit the persistent kernel.  No reference counterpart. */
#define SOFAB200_IPC_HANDLE_BYTES 64
int sofab200_peer_alloc(sofab200_ctx* ctx, size_t bytes, void** dev_ptr, unsigned char handle[SOFAB200_IPC_HANDLE_BYTES]);
int sofab200_peer_open(sofab200_ctx* ctx, const unsigned char handle[SOFAB200_IPC_HANDLE_BYTES], void** dev_ptr);
int sofab200_peer_close(sofab200_ctx* ctx, void* dev_ptr);   /* a pointer obtained from sofab200_peer_open  */
int sofab200_peer_free(sofab200_ctx* ctx, void* dev_ptr);    /* a pointer obtained from sofab200_peer_alloc */
/* Bytes this node's mailbox needs (valid after sofab200_node_set_distributed).  inbox_rows: the largest number of interface rows
 * any rank receives (sum of its nb_count) -- the same value on every rank, also passed in sofab200_peer_desc. */
size_t sofab200_node_peer_bytes(const sofab200_node* node, size_t inbox_rows);
typedef struct sofab200_peer_desc {
    int rank, world;            /* world <= 8 (one NVSwitch domain)                                                        */
    void* const* peer_base;     /* [world] mailbox of every rank as mapped in THIS process (own allocation at [rank])      */
    size_t inbox_rows;          /* rows of one inbox buffer: max over the ranks of the rows a rank receives                */
    const size_t* remote_off;   /* [n_neighbours] (order of sofab200_halo_desc::nb_rank): first inbox row, in neighbour k's
                                 * mailbox, of the block it receives from this rank (= its cumulated nb_count before us)    */
} sofab200_peer_desc;
/* peer->peer_base == NULL leaves peer mode (every rank of a job must run the same loop).  SOFAB200_ERR_UNSUPPORTED when the
 * rank's partition does not fit the persistent kernel. */
int sofab200_node_set_peer(sofab200_node* node, const sofab200_peer_desc* peer);

/* CGLinearSolver keeps `timeStepCount` to silence first-step warnings ([CG]:161-176); reset() restores 0. */
int sofab200_node_reset(sofab200_node* node);

#ifdef __cplusplus
}
#endif
#endif /* SOFA_B200_H */
