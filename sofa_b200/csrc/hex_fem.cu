// HexahedronFEMForceField<B200Vec3Types> (placeholder until the hexa kernels land)
#include "fem_layout.cuh"
using namespace sb;
struct sofab200_hexfem { int real; size_t n_nodes; };
namespace sb {
template <class R> int hex_run(sofab200_hexfem*, bool, const R*, R, NodeEpilogue<R>) { return fail(SOFAB200_ERR_UNSUPPORTED, "hexa not built"); }
template int hex_run<float>(sofab200_hexfem*, bool, const float*, float, NodeEpilogue<float>);
template int hex_run<double>(sofab200_hexfem*, bool, const double*, double, NodeEpilogue<double>);
int hex_partial_count(sofab200_hexfem*) { return 0; }
int hex_real(sofab200_hexfem* ff) { return ff->real; }
size_t hex_nodes(sofab200_hexfem* ff) { return ff->n_nodes; }
}
extern "C" {
int sofab200_hexfem_create(sofab200_ctx*, sofab200_real, size_t, const void*, size_t, const uint32_t*, const sofab200_hexfem_desc*, sofab200_hexfem**) { return fail(SOFAB200_ERR_UNSUPPORTED, "hexa not built"); }
int sofab200_hexfem_destroy(sofab200_hexfem*) { return SOFAB200_OK; }
int sofab200_hexfem_add_force(sofab200_hexfem*, void*, const void*) { return fail(SOFAB200_ERR_UNSUPPORTED, "hexa not built"); }
int sofab200_hexfem_add_dforce(sofab200_hexfem*, void*, const void*, double) { return fail(SOFAB200_ERR_UNSUPPORTED, "hexa not built"); }
int sofab200_hexfem_get(sofab200_hexfem*, const char*, void*) { return fail(SOFAB200_ERR_UNSUPPORTED, "hexa not built"); }
int sofab200_hexfem_stats(const sofab200_hexfem*, uint64_t*) { return fail(SOFAB200_ERR_UNSUPPORTED, "hexa not built"); }
}
