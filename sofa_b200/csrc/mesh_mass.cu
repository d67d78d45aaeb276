// MeshMatrixMass<B200Vec3Types> on the device: the sparse (vertex + edge) mass matrix of
// Sofa/Component/Mass/src/sofa/component/mass/MeshMatrixMass.inl applied node by node.
//   addMDx   :1987-2048   res[i] += dx[i]*vertexMass[i]*Real(factor), then for every edge j in topology order
//                         res[e0] += dx[e1]*(edgeMass[j]*Real(factor)); res[e1] += dx[e0]*(...)
//            lumped:      res[i] += dx[i]*vertexMass[i]*m_massLumpingCoeff*Real(factor)
//   addForce :2072-2092   f[i] += gravity*vertexMass[i]*m_massLumpingCoeff
//   accFromF :2050-2069   a[i] = f[i]/(vertexMass[i]*m_massLumpingCoeff)   (lumped only)
// The reference scatters over the edge list; here every node gathers its own half-edges in ascending edge index starting from the
// vertex term, which is the order the sequential scatter adds them in: same bits, no atomics.  Half-edges are stored as a sliced
// ELL (slices of 32 nodes = one warp, entry j of lane k at base + 32 j + k) so a warp reads consecutive 8- or 12-byte records.
#include <algorithm>
#include <memory>

#include "common.cuh"

using namespace sb;

namespace sb {
constexpr int kMMSlice = 32;
template <class R> struct HalfEdge { uint32_t nb; R m; };

template <class R> __global__ void __launch_bounds__(128) meshmass_mdx_kernel(size_t n, R* __restrict__ res, const R* __restrict__ dx, const R* __restrict__ vm,
                                                                             const uint32_t* __restrict__ slice_base, const uint16_t* __restrict__ valence,
                                                                             const HalfEdge<R>* __restrict__ he, R factor, R coeff, int lumped) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    R rx = res[3 * i], ry = res[3 * i + 1], rz = res[3 * i + 2];
    const R dxx = dx[3 * i], dxy = dx[3 * i + 1], dxz = dx[3 * i + 2];
    const R m = vm[i];
    if (lumped) {
        rx += ((dxx * m) * coeff) * factor; ry += ((dxy * m) * coeff) * factor; rz += ((dxz * m) * coeff) * factor;
    } else {
        rx += (dxx * m) * factor; ry += (dxy * m) * factor; rz += (dxz * m) * factor;
        const HalfEdge<R>* p = he + slice_base[i / kMMSlice] + (i % kMMSlice);
        const int val = valence[i];
        for (int j = 0; j < val; ++j) {
            const HalfEdge<R> h = p[size_t(j) * kMMSlice];
            const R t = h.m * factor;   // tempMass = edgeMass[j] * Real(factor)
            rx += dx[3 * size_t(h.nb)] * t; ry += dx[3 * size_t(h.nb) + 1] * t; rz += dx[3 * size_t(h.nb) + 2] * t;
        }
    }
    res[3 * i] = rx; res[3 * i + 1] = ry; res[3 * i + 2] = rz;
}
template <class R> __global__ void __launch_bounds__(128) meshmass_force_kernel(size_t n, R* __restrict__ f, const R* __restrict__ vm, R gx, R gy, R gz, R coeff) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const R m = vm[i];
    f[3 * i] += (gx * m) * coeff; f[3 * i + 1] += (gy * m) * coeff; f[3 * i + 2] += (gz * m) * coeff;
}
template <class R> __global__ void __launch_bounds__(128) meshmass_acc_kernel(size_t n, R* __restrict__ a, const R* __restrict__ f, const R* __restrict__ vm, R coeff) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const R d = vm[i] * coeff;
    a[3 * i] = f[3 * i] / d; a[3 * i + 1] = f[3 * i + 1] / d; a[3 * i + 2] = f[3 * i + 2] / d;
}
}  // namespace sb

struct sofab200_meshmass {
    virtual ~sofab200_meshmass() {}
    sofab200_ctx* ctx = nullptr;
    int real = 0, lumping = 0;
    size_t n_nodes = 0, n_edges = 0;
    double coeff = 0;
};

namespace sb {
void meshmass_info(const sofab200_meshmass* mm, int* real, size_t* n_nodes, const sofab200_ctx** ctx) { *real = mm->real; *n_nodes = mm->n_nodes; *ctx = mm->ctx; }
template <class R> struct MeshMass : sofab200_meshmass {
    DevBuf<R> vm;
    DevBuf<uint32_t> slice_base;
    DevBuf<uint16_t> valence;
    DevBuf<HalfEdge<R>> he;
};

template <class R> static int meshmass_create(sofab200_ctx* ctx, size_t n, const R* vm, size_t E, const uint32_t* edges, const R* em, int lumping, double coeff,
                                              sofab200_meshmass** out) {
    std::unique_ptr<MeshMass<R>> mm(new MeshMass<R>());
    mm->ctx = ctx; mm->real = sizeof(R) == 4 ? SOFAB200_F32 : SOFAB200_F64; mm->lumping = lumping; mm->n_nodes = n; mm->n_edges = E; mm->coeff = coeff;
    cudaStream_t s = ctx->stream;
    SB_TRY(mm->vm.upload(std::vector<R>(vm, vm + n), s));
    if (!lumping) {
        std::vector<uint16_t> val(n, 0);
        for (size_t j = 0; j < E; ++j) {
            SB_CHECK(edges[2 * j] < n && edges[2 * j + 1] < n, "edge refers to a node index out of range");
            SB_CHECK(val[edges[2 * j]] < 0xFFFF && val[edges[2 * j + 1]] < 0xFFFF, "more than 65535 edges around a node");
            ++val[edges[2 * j]]; ++val[edges[2 * j + 1]];
        }
        const size_t n_slices = (n + kMMSlice - 1) / kMMSlice;
        std::vector<uint32_t> base(n_slices + 1, 0);
        for (size_t sl = 0; sl < n_slices; ++sl) {
            uint16_t mx = 0;
            for (size_t i = sl * kMMSlice; i < std::min(n, (sl + 1) * kMMSlice); ++i) mx = std::max(mx, val[i]);
            const size_t next = size_t(base[sl]) + size_t(mx) * kMMSlice;
            SB_CHECK(next < 0xFFFFFFFFull, "edge table too large for 32-bit offsets");
            base[sl + 1] = uint32_t(next);
        }
        std::vector<HalfEdge<R>> he(base[n_slices], HalfEdge<R>{0u, R(0)});
        std::vector<uint16_t> fill(n, 0);
        auto put = [&](uint32_t at, uint32_t nb, R m) { he[size_t(base[at / kMMSlice]) + size_t(fill[at]++) * kMMSlice + at % kMMSlice] = HalfEdge<R>{nb, m}; };
        // ascending edge index; within one edge the reference updates e[0] first, then e[1] (they are different nodes, so the order is per node anyway)
        for (size_t j = 0; j < E; ++j) { put(edges[2 * j], edges[2 * j + 1], em[j]); put(edges[2 * j + 1], edges[2 * j], em[j]); }
        SB_TRY(mm->valence.upload(val, s)); SB_TRY(mm->slice_base.upload(base, s)); SB_TRY(mm->he.upload(he, s));
    }
    SB_CUDA(cudaStreamSynchronize(s));
    *out = mm.release();
    return SOFAB200_OK;
}
template <class R> static int meshmass_mdx(MeshMass<R>& mm, R* res, const R* dx, double factor) {
    if (!mm.n_nodes) return SOFAB200_OK;
    meshmass_mdx_kernel<R><<<unsigned((mm.n_nodes + 127) / 128), 128, 0, mm.ctx->stream>>>(mm.n_nodes, res, dx, mm.vm.p, mm.slice_base.p, mm.valence.p, mm.he.p, R(factor), R(mm.coeff), mm.lumping);
    mm.ctx->launches++;
    SB_CUDA(cudaGetLastError());
    return SOFAB200_OK;
}
}  // namespace sb

extern "C" {
int sofab200_meshmass_create(sofab200_ctx* ctx, sofab200_real real, size_t n_nodes, const void* vertex_mass_host, size_t n_edges, const uint32_t* edges_host,
                             const void* edge_mass_host, int lumping, double mass_lumping_coeff, sofab200_meshmass** out) {
    SB_CHECK(ctx && out && (vertex_mass_host || n_nodes == 0), "null argument");
    SB_CHECK(lumping || n_edges == 0 || (edges_host && edge_mass_host), "the sparse mass matrix needs the edge list and the edge masses");
    SB_CUDA(cudaSetDevice(ctx->device));
    if (real == SOFAB200_F32) return meshmass_create<float>(ctx, n_nodes, static_cast<const float*>(vertex_mass_host), n_edges, edges_host, static_cast<const float*>(edge_mass_host), lumping, mass_lumping_coeff, out);
    return meshmass_create<double>(ctx, n_nodes, static_cast<const double*>(vertex_mass_host), n_edges, edges_host, static_cast<const double*>(edge_mass_host), lumping, mass_lumping_coeff, out);
}
int sofab200_meshmass_destroy(sofab200_meshmass* mm) { delete mm; return SOFAB200_OK; }
int sofab200_meshmass_add_mdx(sofab200_meshmass* mm, void* res_dev, const void* dx_dev, double factor) {
    SB_CHECK(mm, "null argument");
    if (!mm->n_nodes) return SOFAB200_OK;      // an empty state has no device arrays
    SB_CHECK(res_dev && dx_dev, "null argument");
    SB_CHECK(res_dev != dx_dev, "res and dx must be distinct vectors");
    if (mm->real == SOFAB200_F32) return meshmass_mdx(*static_cast<MeshMass<float>*>(mm), static_cast<float*>(res_dev), static_cast<const float*>(dx_dev), factor);
    return meshmass_mdx(*static_cast<MeshMass<double>*>(mm), static_cast<double*>(res_dev), static_cast<const double*>(dx_dev), factor);
}
int sofab200_meshmass_add_force(sofab200_meshmass* mm, void* f_dev, const double gravity[3]) {
    SB_CHECK(mm && gravity, "null argument");
    if (!mm->n_nodes) return SOFAB200_OK;
    SB_CHECK(f_dev, "null argument");
    const unsigned g = unsigned((mm->n_nodes + 127) / 128);
    if (mm->real == SOFAB200_F32) { auto* m = static_cast<MeshMass<float>*>(mm); meshmass_force_kernel<float><<<g, 128, 0, mm->ctx->stream>>>(mm->n_nodes, static_cast<float*>(f_dev), m->vm.p, float(gravity[0]), float(gravity[1]), float(gravity[2]), float(mm->coeff)); }
    else { auto* m = static_cast<MeshMass<double>*>(mm); meshmass_force_kernel<double><<<g, 128, 0, mm->ctx->stream>>>(mm->n_nodes, static_cast<double*>(f_dev), m->vm.p, gravity[0], gravity[1], gravity[2], mm->coeff); }
    mm->ctx->launches++;
    SB_CUDA(cudaGetLastError());
    return SOFAB200_OK;
}
int sofab200_meshmass_acc_from_f(sofab200_meshmass* mm, void* a_dev, const void* f_dev) {
    SB_CHECK(mm && (mm->n_nodes == 0 || (a_dev && f_dev)), "null argument");
    if (!mm->lumping) return fail(SOFAB200_ERR_UNSUPPORTED, "accFromF cannot be used with the sparse MeshMatrixMass (the reference refuses too, MeshMatrixMass.inl:2053-2058): use lumping");
    if (!mm->n_nodes) return SOFAB200_OK;
    const unsigned g = unsigned((mm->n_nodes + 127) / 128);
    if (mm->real == SOFAB200_F32) { auto* m = static_cast<MeshMass<float>*>(mm); meshmass_acc_kernel<float><<<g, 128, 0, mm->ctx->stream>>>(mm->n_nodes, static_cast<float*>(a_dev), static_cast<const float*>(f_dev), m->vm.p, float(mm->coeff)); }
    else { auto* m = static_cast<MeshMass<double>*>(mm); meshmass_acc_kernel<double><<<g, 128, 0, mm->ctx->stream>>>(mm->n_nodes, static_cast<double*>(a_dev), static_cast<const double*>(f_dev), m->vm.p, mm->coeff); }
    mm->ctx->launches++;
    SB_CUDA(cudaGetLastError());
    return SOFAB200_OK;
}
}  // extern "C"
