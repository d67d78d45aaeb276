// The opaque handle behind sofab200_tetfem*: TetrahedronFEMForceField / TetrahedralCorotationalFEMForceField (tet_fem.cu, kind 0) and
// FastTetrahedralCorotationalForceField (fast_fem.cu, kind 1) share it, so that a solver node drives either through the same calls.
#pragma once
#include "common.cuh"

struct sofab200_tetfem {
    virtual ~sofab200_tetfem() {}
    sofab200_ctx* ctx = nullptr;
    int real = 0, method = 1;
    int kind = 0;      // 0: TetFF<R>; 1: FastFF<R>
    size_t n_nodes = 0, n_tets = 0;
};

namespace sb {
template <class R> struct NodeEpilogue;
template <class R> struct TileDev;
template <class R> struct FusedCG;
// fast_fem.cu
int fast_create(sofab200_ctx* ctx, int real, size_t n_nodes, const void* rest, size_t n_tets, const uint32_t* tets, const sofab200_tetfem_desc* desc, sofab200_tetfem** out);
template <class R> int fast_run(sofab200_tetfem* ff, bool dforce, const R* in, R k_factor, NodeEpilogue<R> ep, bool skip_gather);
template <class R> int fast_cg_fused(sofab200_tetfem* ff, R k_factor, FusedCG<R> a, size_t sync_capacity, bool dry_run, int* info);
template <class R> TileDev<R> fast_tiledev(sofab200_tetfem* ff);          // the plan of the addDForce pass (edges)
int fast_partial_count(sofab200_tetfem* ff);
size_t fast_tile_node_count(sofab200_tetfem* ff);
size_t fast_shared_slot_count(sofab200_tetfem* ff);
int fast_get(sofab200_tetfem* ff, const char* what, void* out_host);
int fast_stats(const sofab200_tetfem* ff, uint64_t out[8]);
}  // namespace sb
