// MechanicalObject vector operations and the CG vector updates (streaming kernels).
// Vec3 operations of the reference are componentwise, so the kernels run on the flat 3n array.
#pragma once
#include <cooperative_groups.h>

#include "cg_persist.cuh"

namespace sb {

enum VOpKind { VOP_CLEAR = 0, VOP_SCALE, VOP_EQ_BF, VOP_COPY, VOP_PEQ, VOP_PEQ_BF, VOP_AVF, VOP_EQ_AB, VOP_EQ_ABF };

constexpr int kVecBlock = 256;
inline int vec_grid(size_t n, int sm_count) {
    size_t b = (n + kVecBlock - 1) / kVecBlock;
    const size_t cap = size_t(sm_count) * 8;  // a few CTAs per SM, grid-stride beyond that
    if (b > cap) b = cap;
    return int(b < 1 ? 1 : b);
}

// MechanicalObject.inl:1930-2072 helper forms
template <class R, int KIND> __global__ void __launch_bounds__(kVecBlock) vop_kernel(size_t n3, R* __restrict__ r, const R* a, const R* b, R k) {
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n3; i += size_t(gridDim.x) * blockDim.x) {
        if (KIND == VOP_CLEAR) r[i] = R(0);
        else if (KIND == VOP_SCALE) r[i] *= k;                       // vOp_vf
        else if (KIND == VOP_EQ_BF) r[i] = b[i] * k;                 // vOp_vbf
        else if (KIND == VOP_COPY) r[i] = a[i];                      // vOp_va
        else if (KIND == VOP_PEQ) r[i] += b[i];                      // vOp_vb
        else if (KIND == VOP_PEQ_BF) r[i] += b[i] * k;               // vOp_v_inc_bf
        else if (KIND == VOP_AVF) { R t = r[i]; t *= k; t += a[i]; r[i] = t; }  // vOp_avf
        else if (KIND == VOP_EQ_AB) r[i] = a[i] + b[i];              // vOp_vab
        else if (KIND == VOP_EQ_ABF) r[i] = a[i] + b[i] * k;         // vOp_vabf
    }
}

// MechanicalObject::accumulateForce, MechanicalObject.inl:1356-1375: f[i] += externalForce[i] for the rows that differ from Deriv()
template <class R> __global__ void __launch_bounds__(kVecBlock) accumulate_force_kernel(size_t n, R* __restrict__ f, const R* __restrict__ ext) {
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const R a = ext[3 * i], b = ext[3 * i + 1], c = ext[3 * i + 2];
        if (!(a == R(0) && b == R(0) && c == R(0))) { f[3 * i] += a; f[3 * i + 1] += b; f[3 * i + 2] += c; }
    }
}

// what the last CTA does with a finished dot product
enum DotFinish { DF_STORE = 0, DF_CG_NORMB = 1, DF_CG_RHO = 2 };

__device__ inline void dot_finish_action(int action, double s, double* result, CGDev* cg) {
    if (result) *result = s;
    if (action == DF_CG_NORMB) {
        cg->normb = sqrt(s);
        if (cg->normb == 0.0) { cg->done = 1; cg->nb_iter = 0; cg->end_cond = 4; }
    } else if (action == DF_CG_RHO) cg_after_rho(cg, s);
}

template <class R> __device__ __forceinline__ void dot_epilogue(double part, double* partials, unsigned* counter, int action, double* result, CGDev* cg) {
    __shared__ double red[32];
    __shared__ bool is_last;
    const double tot = block_sum(part, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = tot;
        __threadfence();
        const unsigned ticket = atomicInc(counter, gridDim.x - 1);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double s = 0.0;
    for (int i = threadIdx.x; i < int(gridDim.x); i += blockDim.x) s += __ldcg(partials + i);
    __syncthreads();
    s = block_sum(s, red);
    if (threadIdx.x == 0) dot_finish_action(action, s, result, cg);
}

// MechanicalObject::vDot in double with a fixed summation tree
template <class R> __global__ void __launch_bounds__(kVecBlock) vdot_kernel(size_t n3, const R* __restrict__ a, const R* __restrict__ b, double* partials,
                                                                             unsigned* counter, int action, double* result, CGDev* cg) {
    if (cg && cg->done) return;
    double part = 0.0;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n3; i += size_t(gridDim.x) * blockDim.x) part += double(a[i]) * double(b[i]);
    dot_epilogue<R>(part, partials, counter, action, result, cg);
}

// node-masked variant (distributed vDot: owned nodes only)
template <class R> __global__ void __launch_bounds__(kVecBlock) vdot_masked_kernel(size_t n, const R* __restrict__ a, const R* __restrict__ b, const unsigned char* __restrict__ mask,
                                                                                    double* partials, unsigned* counter, double* result, const CGDev* cg) {
    if (cg && cg->done) return;
    double part = 0.0;
    for (size_t g = size_t(blockIdx.x) * blockDim.x + threadIdx.x; g < n; g += size_t(gridDim.x) * blockDim.x) {
        if (mask && !mask[g]) continue;
        part += double(a[3 * g]) * double(b[3 * g]) + double(a[3 * g + 1]) * double(b[3 * g + 1]) + double(a[3 * g + 2]) * double(b[3 * g + 2]);
    }
    dot_epilogue<R>(part, partials, counter, DF_STORE, result, nullptr);
}

// vMultiOp integration fast path, MechanicalObject.inl:2208-2241
template <class R> __global__ void __launch_bounds__(kVecBlock) integrate_kernel(size_t n3, R* __restrict__ v, R* __restrict__ x, const R* __restrict__ a, R f_v_a, int f_v_a_is_one, R f_x_v) {
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n3; i += size_t(gridDim.x) * blockDim.x) {
        R vv = v[i];
        if (f_v_a_is_one) vv += a[i]; else vv += a[i] * f_v_a;
        v[i] = vv;
        x[i] += vv * f_x_v;
    }
}

// DiagonalMass pieces (DiagonalMass.inl:535-575,1392-1413)
template <class R> __global__ void __launch_bounds__(kVecBlock) mass_mdx_kernel(size_t n, R* __restrict__ res, const R* __restrict__ dx, const R* __restrict__ m, R factor, int is_one) {
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < 3 * n; i += size_t(gridDim.x) * blockDim.x) {
        const R mm = m[i / 3];
        if (is_one) res[i] += dx[i] * mm; else res[i] += (dx[i] * mm) * factor;
    }
}
template <class R> __global__ void __launch_bounds__(kVecBlock) mass_gravity_kernel(size_t n, R* __restrict__ f, const R* __restrict__ m, R gx, R gy, R gz) {
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < 3 * n; i += size_t(gridDim.x) * blockDim.x) {
        const int c = int(i % 3);
        const R g = c == 0 ? gx : (c == 1 ? gy : gz);
        f[i] += g * m[i / 3];
    }
}
// f[i] += (cx, cy, cz)   (UniformMass::addForce: the weight of a node is the same vector for every node)
template <class R> __global__ void __launch_bounds__(kVecBlock) add_const3_kernel(size_t n, R* __restrict__ f, R cx, R cy, R cz) {
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < 3 * n; i += size_t(gridDim.x) * blockDim.x) {
        const int c = int(i % 3);
        f[i] += c == 0 ? cx : (c == 1 ? cy : cz);
    }
}
template <class R> __global__ void __launch_bounds__(kVecBlock) mass_acc_kernel(size_t n, R* __restrict__ a, const R* __restrict__ f, const R* __restrict__ m) {
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < 3 * n; i += size_t(gridDim.x) * blockDim.x) a[i] = f[i] / m[i / 3];
}
// FixedProjectiveConstraint::projectResponse, indexed form
template <class R> __global__ void __launch_bounds__(kVecBlock) fixed_project_kernel(size_t n_idx, const uint32_t* __restrict__ idx, R* __restrict__ res) {
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_idx; i += size_t(gridDim.x) * blockDim.x) {
        const size_t g = idx[i];
        res[3 * g] = R(0); res[3 * g + 1] = R(0); res[3 * g + 2] = R(0);
    }
}
// PlaneForceField::addForce / addDForce (PlaneForceField.inl:158-226), one thread per node, the reference's operation order
template <class R> __global__ void __launch_bounds__(kVecBlock) plane_add_force_kernel(size_t n, PlaneDev<R> P, R* __restrict__ f, const R* __restrict__ x, const R* __restrict__ v,
                                                                                       unsigned char* __restrict__ contacts) {
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        R d = x[3 * i] * P.nx; d += x[3 * i + 1] * P.ny; d += x[3 * i + 2] * P.nz;     // p * planeN
        d = d - P.d;
        unsigned char hit = 0;
        if (P.bilateral || d < R(0)) {
            const R fi = -P.stiff * d, di = -P.damp * d;
            R f0 = P.nx * fi - v[3 * i] * di, f1 = P.ny * fi - v[3 * i + 1] * di, f2 = P.nz * fi - v[3 * i + 2] * di;
            R amp = f0 * f0; amp += f1 * f1; amp += f2 * f2;
            if (P.limit2 > R(0) && amp > P.limit2) { const R s = sqrt(P.limit2 / amp); f0 *= s; f1 *= s; f2 *= s; }
            f[3 * i] += f0; f[3 * i + 1] += f1; f[3 * i + 2] += f2;
            hit = 1;
        }
        contacts[i] = hit;
    }
}
template <class R> __global__ void __launch_bounds__(kVecBlock) plane_add_dforce_kernel(size_t n, PlaneDev<R> P, R fact, R* __restrict__ df, const R* __restrict__ dx,
                                                                                        const unsigned char* __restrict__ contacts) {
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        if (!contacts[i]) continue;
        R s = dx[3 * i] * P.nx; s += dx[3 * i + 1] * P.ny; s += dx[3 * i + 2] * P.nz;
        const R t = fact * s;
        df[3 * i] = df[3 * i] + P.nx * t; df[3 * i + 1] = df[3 * i + 1] + P.ny * t; df[3 * i + 2] = df[3 * i + 2] + P.nz * t;
    }
}

// generic per-node epilogue without element contributions (right-hand side when the stiffness term is absent)
template <class R> __global__ void __launch_bounds__(kVecBlock) node_only_kernel(size_t n, NodeEpilogue<R> ep) {
    for (size_t g = size_t(blockIdx.x) * blockDim.x + threadIdx.x; g < n; g += size_t(gridDim.x) * blockDim.x) {
        R ax, ay, az;
        node_pre(ep, uint32_t(g), ax, ay, az);
        node_mass(ep, ep.pre_kind, uint32_t(g), ax, ay, az);
        node_post(ep, uint32_t(g), ax, ay, az);
    }
}

// ---- CG vector steps -------------------------------------------------------------------------------
// 16-byte vector accesses on the flat 3n array (cudaMalloc'ed vectors are 256-byte aligned); the last n3 % 4 scalars
// are handled by the first threads of CTA 0.
// p = r (first iteration) or p = p*beta + r  (cgstep_beta -> vOp_avf), CGLinearSolver.inl:184-197
template <class R> __global__ void __launch_bounds__(kVecBlock) cg_p_update_kernel(size_t n3, R* __restrict__ p, const R* __restrict__ r, const CGDev* cg) {
    if (cg->done) return;
    typedef typename Vec4T<R>::T V;
    constexpr int N = Vec4T<R>::N;
    const bool first = cg->it == 1;
    const R beta = R(cg->rho / cg->rho_1);
    const size_t nv = n3 / N, stride = size_t(gridDim.x) * blockDim.x, t0 = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    V* pv = reinterpret_cast<V*>(p); const V* rv = reinterpret_cast<const V*>(r);
    for (size_t i = t0; i < nv; i += stride) {
        if (first) pv[i] = rv[i];
        else { V a = pv[i]; v4_avf(a, rv[i], beta); pv[i] = a; }
    }
    for (size_t i = nv * N + t0; i < n3; i += stride) {
        if (first) p[i] = r[i];
        else { R t = p[i]; t *= beta; t += r[i]; p[i] = t; }
    }
}
// x += p*alpha ; r += q*(-alpha), then rho' = r.r for the next iteration (xr_one, cg_persist.cuh)
template <class R> __global__ void __launch_bounds__(kVecBlock) cg_xr_update_kernel(size_t n3, R* __restrict__ x, R* __restrict__ r, const R* __restrict__ p, const R* __restrict__ q,
                                                                                     CGDev* cg, double* partials, unsigned* counter, int fused_rho) {
    if (cg->done) return;
    const double alpha_d = cg->alpha;
    const R alpha = R(alpha_d), malpha = R(-alpha_d);
    const bool a_one = (alpha_d == 1.0), ma_one = (-alpha_d == 1.0);
    const size_t stride = size_t(gridDim.x) * blockDim.x, t0 = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    double part = 0.0;
    if (sizeof(R) == 4) {
        const size_t nv = n3 / 4;
        float4* xv = reinterpret_cast<float4*>(x); float4* rv = reinterpret_cast<float4*>(r);
        const float4* pv = reinterpret_cast<const float4*>(p); const float4* qv = reinterpret_cast<const float4*>(q);
        for (size_t i = t0; i < nv; i += stride) {
            float4 xx = xv[i], rr = rv[i]; const float4 pp = pv[i], qq = qv[i];
            part += xr_one<float>(xx.x, rr.x, pp.x, qq.x, alpha, malpha, a_one, ma_one);
            part += xr_one<float>(xx.y, rr.y, pp.y, qq.y, alpha, malpha, a_one, ma_one);
            part += xr_one<float>(xx.z, rr.z, pp.z, qq.z, alpha, malpha, a_one, ma_one);
            part += xr_one<float>(xx.w, rr.w, pp.w, qq.w, alpha, malpha, a_one, ma_one);
            xv[i] = xx; rv[i] = rr;
        }
        for (size_t i = nv * 4 + t0; i < n3; i += stride) part += xr_one<R>(x[i], r[i], p[i], q[i], alpha, malpha, a_one, ma_one);
    } else {
        for (size_t i = t0; i < n3; i += stride) part += xr_one<R>(x[i], r[i], p[i], q[i], alpha, malpha, a_one, ma_one);
    }
    if (fused_rho) dot_epilogue<R>(part, partials, counter, DF_CG_RHO, nullptr, cg);   // distributed: rho comes from an all-reduced masked dot
}
// ---- multi-GPU halo kernels ------------------------------------------------------------------------
// pack the interface rows that go to the neighbours (concatenated per neighbour)
template <class R> __global__ void __launch_bounds__(kVecBlock) halo_pack_kernel(size_t n_send, const uint32_t* __restrict__ send_idx, const R* __restrict__ q, R* __restrict__ sendbuf, const CGDev* cg) {
    if (cg && cg->done) return;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_send; i += size_t(gridDim.x) * blockDim.x) {
        const size_t g = send_idx[i];
        sendbuf[3 * i] = q[3 * g]; sendbuf[3 * i + 1] = q[3 * g + 1]; sendbuf[3 * i + 2] = q[3 * g + 2];
    }
}
// every interface node: sum of the partial values of all sharing ranks in ascending rank order (src: -1 own value, -2 absent,
// >= 0 row of the receive buffer) => identical bits on every sharing rank
template <class R> __global__ void __launch_bounds__(kVecBlock) halo_sum_kernel(size_t n_if, int max_sh, const uint32_t* __restrict__ if_idx, const int32_t* __restrict__ src,
                                                                                 const R* __restrict__ recvbuf, R* __restrict__ q, const CGDev* cg) {
    if (cg && cg->done) return;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_if; i += size_t(gridDim.x) * blockDim.x) {
        const size_t g = if_idx[i];
        const R ox = q[3 * g], oy = q[3 * g + 1], oz = q[3 * g + 2];
        R ax = 0, ay = 0, az = 0;
        bool first = true;
        for (int j = 0; j < max_sh; ++j) {
            const int32_t sidx = src[i * max_sh + j];
            if (sidx == -2) continue;
            const R vx = sidx < 0 ? ox : recvbuf[3 * size_t(sidx)], vy = sidx < 0 ? oy : recvbuf[3 * size_t(sidx) + 1], vz = sidx < 0 ? oz : recvbuf[3 * size_t(sidx) + 2];
            if (first) { ax = vx; ay = vy; az = vz; first = false; } else { ax += vx; ay += vy; az += vz; }
        }
        q[3 * g] = ax; q[3 * g + 1] = ay; q[3 * g + 2] = az;
    }
}
// out[g] = owned[g] ? in[g] : 0  (the start value of a shared node enters the distributed sum on its owner only)
template <class R> __global__ void __launch_bounds__(kVecBlock) mask_rows_kernel(size_t n, const unsigned char* __restrict__ owned, const R* __restrict__ in, R* __restrict__ out) {
    for (size_t g = size_t(blockIdx.x) * blockDim.x + threadIdx.x; g < n; g += size_t(gridDim.x) * blockDim.x) {
        const bool o = owned[g] != 0;
        out[3 * g] = o ? in[3 * g] : R(0); out[3 * g + 1] = o ? in[3 * g + 1] : R(0); out[3 * g + 2] = o ? in[3 * g + 2] : R(0);
    }
}
// ---- halo sum over peer memory (multi-GPU, outside the CG kernel: addForce and the right-hand side) ---------------------------
// One CTA: every interface row of q is stored into the sharing neighbours' inboxes as self-validating 8-byte words (payload +
// sequence number, see cg_persist.cuh), then the rows are rebuilt from the own value and the neighbours' words in ascending rank
// order.  Two inbox buffers alternate from call to call (a neighbour may start call k+1 before this rank has read call k).
template <class R> __global__ void __launch_bounds__(1024) halo_peer_kernel(PeerDev<R> P, size_t n_if, const uint32_t* __restrict__ if_idx, R* q,
                                                                             unsigned long long* hcount, size_t buf_words, int* fail_flag, CGDev* cg) {
    // a late or missing neighbour must not yield silently wrong rows: the flag is sticky, the step's solve is marked failed (end_cond 99),
    // sofab200_node_last_solve / _step_host report it, and the call counter is not advanced past a failed exchange
    __shared__ int s_fail;
    if (threadIdx.x == 0) s_fail = 0;
    const unsigned long long c = *hcount + 1;
    const unsigned seq = unsigned(c);
    const size_t boff = (1 + (c & 1ull)) * buf_words;       // buffer 0 belongs to the CG kernel
    const int ms1 = P.max_sh - 1;
    __syncthreads();                                          // everyone has read hcount before thread 0 bumps it
    for (size_t row = threadIdx.x; row < n_if; row += blockDim.x) {
        const size_t g3 = 3 * size_t(if_idx[row]);
        const R v0 = q[g3], v1 = q[g3 + 1], v2 = q[g3 + 2];
        for (int e = 0; e < ms1; ++e) {
            const int2 to = P.if_send[row * ms1 + e];
            if (to.x >= 0) { unsigned long long* d = P.nb_inbox[to.x] + boff + size_t(InboxWords<R>::N) * size_t(to.y); inbox_put(d, 0, v0, seq); inbox_put(d, 1, v1, seq); inbox_put(d, 2, v2, seq); }
        }
    }
    for (size_t row = threadIdx.x; row < n_if; row += blockDim.x) {
        const size_t g3 = 3 * size_t(if_idx[row]);
        R s0 = R(0), s1 = R(0), s2 = R(0);
        for (int j = 0; j < P.max_sh; ++j) {
            const int sj = P.src[row * P.max_sh + j];
            R c0 = R(0), c1 = R(0), c2 = R(0);
            if (sj == -1) { c0 = q[g3]; c1 = q[g3 + 1]; c2 = q[g3 + 2]; }
            else if (sj >= 0) {
                const unsigned long long* w = P.inbox + boff + size_t(InboxWords<R>::N) * size_t(sj);
                const long long t0 = poll_clock();
                while (!(inbox_get(w, 0, seq, c0) && inbox_get(w, 1, seq, c1) && inbox_get(w, 2, seq, c2)))
                    if (poll_clock() - t0 > kSyncTimeoutCycles) { s_fail = 1; break; }
            }
            if (j == 0) { s0 = c0; s1 = c1; s2 = c2; } else { s0 += c0; s1 += c1; s2 += c2; }
        }
        q[g3] = s0; q[g3 + 1] = s1; q[g3 + 2] = s2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_fail) { *fail_flag = 1; if (cg) { cg->done = 1; cg->end_cond = 99; } }
        else *hcount = c;
    }
}

// the scalar bookkeeping of the CG after an all-reduced dot product
__global__ void cg_scalar_kernel(CGDev* cg, const double* value, int action) {
    if (cg->done) return;
    if (action == DF_CG_NORMB || action == DF_CG_RHO) dot_finish_action(action, *value, nullptr, cg);
    else cg_after_den(cg, *value);
}

// ---- fused tail of a CG iteration (cooperative launch, two grid barriers) --------------------------------------------
//   A  boundary gather of q = A p (shared nodes) + the rest of den = p.q      | grid barrier, every CTA sums the partials
//   B  alpha = rho/den ; x += alpha p ; r -= alpha q ; partial rho' = r.r     | grid barrier, every CTA sums the partials
//   C  tolerance test ; beta = rho'/rho ; p = p*beta + r  (the NEXT iteration's direction)
// Every CTA adds the same partials in the same order, so all take the same branch; CTA 0 records the scalars in CGDev.
constexpr int kTailBlock = kGatherChunk;
template <class R> __global__ void __launch_bounds__(kTailBlock) cg_tail_kernel(TileDev<R> d, NodeEpilogue<R> ep, size_t n3, R* __restrict__ x, R* __restrict__ r, R* __restrict__ p,
                                                                                 const R* __restrict__ q, CGDev* cg, double* partials_den, int n_tile_partials, double* partials_rho) {
    namespace cgp = cooperative_groups;
    __shared__ double red[32];
    __shared__ double bcast;
    if (cg->done) return;
    cgp::grid_group grid = cgp::this_grid();
    // snapshot of the scalars this iteration starts from (CTA 0 rewrites CGDev after each barrier)
    const double rho = cg->rho, normb = cg->normb, tol = cg->tolerance, thr = cg->threshold;
    const int it = cg->it;
    const unsigned tsc = cg->time_step_count, max_iter = cg->max_iter;
    trace_mark(ep.trace, kTraceTail, 0);
    // ---- A
    double part = 0.0;
    for (int chunk = blockIdx.x; chunk < d.n_chunks; chunk += gridDim.x) part += gather_chunk<R>(d, ep, chunk, threadIdx.x);
    __syncthreads();
    part = block_sum(part, red);
    if (threadIdx.x == 0) partials_den[n_tile_partials + blockIdx.x] = part;
    trace_mark(ep.trace, kTraceTail, 1);
    grid.sync();
    trace_mark(ep.trace, kTraceTail, 2);
    const double den = sum_partials_all<R>(partials_den, n_tile_partials + int(gridDim.x), red, &bcast);
    bool stop = false;
    if (den != 0.0) { if (fabs(den) <= thr && !(it == 1 && tsc == 0)) stop = true; } else stop = true;
    if (blockIdx.x == 0 && threadIdx.x == 0) cg_after_den(cg, den);
    if (stop) return;
    const double alpha_d = rho / den;
    trace_mark(ep.trace, kTraceTail, 3);
    // ---- B
    const R alpha = R(alpha_d), malpha = R(-alpha_d);
    const bool a_one = (alpha_d == 1.0), ma_one = (-alpha_d == 1.0);
    const size_t stride = size_t(gridDim.x) * blockDim.x, t0 = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    double prr = 0.0;
    if (sizeof(R) == 4) {
        const size_t nv = n3 / 4;
        float4* xv = reinterpret_cast<float4*>(x); float4* rv = reinterpret_cast<float4*>(r);
        const float4* pv = reinterpret_cast<const float4*>(p); const float4* qv = reinterpret_cast<const float4*>(q);
        for (size_t i = t0; i < nv; i += stride) {
            float4 xx = xv[i], rr = rv[i]; const float4 pp = pv[i], qq = __ldcg(qv + i);
            prr += xr_one<float>(xx.x, rr.x, pp.x, qq.x, alpha, malpha, a_one, ma_one);
            prr += xr_one<float>(xx.y, rr.y, pp.y, qq.y, alpha, malpha, a_one, ma_one);
            prr += xr_one<float>(xx.z, rr.z, pp.z, qq.z, alpha, malpha, a_one, ma_one);
            prr += xr_one<float>(xx.w, rr.w, pp.w, qq.w, alpha, malpha, a_one, ma_one);
            xv[i] = xx; rv[i] = rr;
        }
        for (size_t i = nv * 4 + t0; i < n3; i += stride) { R qq = __ldcg(q + i); prr += xr_one<R>(x[i], r[i], p[i], qq, alpha, malpha, a_one, ma_one); }
    } else {
        for (size_t i = t0; i < n3; i += stride) { R qq = __ldcg(q + i); prr += xr_one<R>(x[i], r[i], p[i], qq, alpha, malpha, a_one, ma_one); }
    }
    __syncthreads();
    prr = block_sum(prr, red);
    if (threadIdx.x == 0) partials_rho[blockIdx.x] = prr;
    trace_mark(ep.trace, kTraceTail, 4);
    grid.sync();
    trace_mark(ep.trace, kTraceTail, 5);
    const double rho_new = sum_partials_all<R>(partials_rho, int(gridDim.x), red, &bcast);
    const int it2 = it + 1;
    bool stop2 = unsigned(it2) > max_iter;
    if (!stop2) { const double err = sqrt(rho_new) / normb; if (err <= tol && !(it2 == 1 && tsc == 0)) stop2 = true; }
    if (blockIdx.x == 0 && threadIdx.x == 0) cg_after_rho(cg, rho_new);
    if (stop2) return;
    // ---- C : p = p*beta + r   (cgstep_beta)
    const R beta = R(rho_new / rho);
    if (sizeof(R) == 4) {
        const size_t nv = n3 / 4;
        float4* pv = reinterpret_cast<float4*>(p); const float4* rv = reinterpret_cast<const float4*>(r);
        for (size_t i = t0; i < nv; i += stride) { float4 a = pv[i]; v4_avf(a, rv[i], float(beta)); pv[i] = a; }
        for (size_t i = nv * 4 + t0; i < n3; i += stride) { R t = p[i]; t *= beta; t += r[i]; p[i] = t; }
    } else {
        for (size_t i = t0; i < n3; i += stride) { R t = p[i]; t *= beta; t += r[i]; p[i] = t; }
    }
    if (ep.trace && threadIdx.x == 0) {   // duration of phase C (the last iteration of a solve returns before it)
        unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        ep.trace[kTraceTail + blockIdx.x * kTraceWords + 6] = t - ep.trace[kTraceTail + blockIdx.x * kTraceWords + 5];
    }
}

struct CGBegin { unsigned max_iter; double tolerance, threshold; };
// fail: sticky flag of the peer-memory halo exchange (null outside peer mode); a solve whose inputs went through a failed exchange does not run
__global__ void cg_begin_kernel(CGDev* cg, CGBegin b, const int* fail) {
    cg->done = 0; cg->nb_iter = 0; cg->end_cond = 0; cg->it = 0;
    if (fail && *fail) { cg->done = 1; cg->end_cond = 99; }
    cg->max_iter = b.max_iter; cg->tolerance = b.tolerance; cg->threshold = b.threshold;
    cg->rho = 0; cg->rho_1 = 0; cg->den = 0; cg->alpha = 0; cg->normb = 0;
    cg->n_err = 1; cg->graph_error[0] = 1.0; cg->n_den = 0;
}
__global__ void cg_end_kernel(CGDev* cg) { cg->time_step_count++; }

}  // namespace sb
