// SOFA-side glue (compiles only inside a SOFA build tree).
// Device DataTypes "B200Vec3f" / "B200Vec3d": same Coord/Deriv/Real as Vec3Types, vectors are
// sofa::type::vector_device with a memory manager that allocates through the CUDA runtime, so that every
// Data<VecCoord>/Data<VecDeriv> of a MechanicalObject<B200Vec3fTypes> owns a lazily synchronised host/device pair
// (Sofa/framework/Type/src/sofa/type/vector_device.h:50-155).  Kernels only ever see the raw device pointer
// obtained with deviceRead()/deviceWrite() at call time; the AoS Vec3 layout is the one libsofa_b200 expects.
#pragma once
#include <cuda_runtime.h>
#include <sofa/defaulttype/VecTypes.h>
#include <sofa/type/vector_device.h>

#include <sofa_b200.h>

namespace sofa::b200 {

/// MemoryManager concept of vector_device (the role CudaMemoryManager plays for SofaCUDA,
/// applications/plugins/SofaCUDA/Core/src/sofa/gpu/cuda/CudaMemoryManager.h:39-170)
template <class T> class B200MemoryManager : public sofa::type::MemoryManager<T> {
public:
    typedef T* host_pointer;
    typedef void* device_pointer;
    typedef unsigned int gl_buffer;
    enum { MAX_DEVICES = 8, BSIZE = 64, SUPPORT_GL_BUFFER = 0 };
    static int numDevices() { int n = 0; cudaGetDeviceCount(&n); return n; }
    static void hostAlloc(void** p, int n) { cudaMallocHost(p, n); }            // pinned: H2D/D2H of state vectors are async
    static void hostFree(const host_pointer p) { cudaFreeHost(p); }
    static void deviceAlloc(int d, void** p, int n) { cudaSetDevice(d); cudaMalloc(p, n); }
    static void deviceFree(int d, const device_pointer p) { cudaSetDevice(d); cudaFree(p); }
    static void memcpyHostToDevice(int d, device_pointer dst, const void* src, size_t n) { cudaSetDevice(d); cudaMemcpy(dst, src, n, cudaMemcpyHostToDevice); }
    static void memcpyDeviceToHost(int d, void* dst, const void* src, size_t n) { cudaSetDevice(d); cudaMemcpy(dst, src, n, cudaMemcpyDeviceToHost); }
    static void memcpyDeviceToDevice(int d, device_pointer dst, const void* src, size_t n) { cudaSetDevice(d); cudaMemcpy(dst, src, n, cudaMemcpyDeviceToDevice); }
    static void memsetDevice(int d, device_pointer dst, int v, size_t n) { cudaSetDevice(d); cudaMemset(dst, v, n); }
    static int getBufferDevice() { int d = 0; cudaGetDevice(&d); return d; }
    static bool bufferAlloc(gl_buffer*, int, bool) { return false; }
    static void bufferFree(const gl_buffer) {}
    static bool bufferRegister(const gl_buffer) { return false; }
    static void bufferUnregister(const gl_buffer) {}
    static bool bufferMapToDevice(device_pointer*, const gl_buffer) { return false; }
    static void bufferUnmapToDevice(device_pointer*, const gl_buffer) {}
    static device_pointer deviceOffset(device_pointer p, size_t off) { return static_cast<char*>(p) + off; }
    static device_pointer null() { return nullptr; }
    static bool isNull(device_pointer p) { return p == nullptr; }
};

template <class T> using B200Vector = sofa::type::vector_device<T, B200MemoryManager<T>>;

/// DataTypes concept (Sofa/framework/DefaultType/src/sofa/defaulttype/VecTypes.h:45-234): only the containers change.
template <class TReal> class B200Vec3Types : public sofa::defaulttype::StdVectorTypes<sofa::type::Vec<3, TReal>, sofa::type::Vec<3, TReal>, TReal> {
public:
    typedef sofa::type::Vec<3, TReal> Coord;
    typedef Coord Deriv;
    typedef TReal Real;
    typedef B200Vector<Coord> VecCoord;
    typedef B200Vector<Deriv> VecDeriv;
    typedef B200Vector<Real> VecReal;
    static constexpr sofab200_real abiReal = sizeof(TReal) == 4 ? SOFAB200_F32 : SOFAB200_F64;
    static const char* Name() { return sizeof(TReal) == 4 ? "B200Vec3f" : "B200Vec3d"; }
};
typedef B200Vec3Types<float> B200Vec3fTypes;
typedef B200Vec3Types<double> B200Vec3dTypes;

/// One libsofa_b200 context per simulation thread (DefaultAnimationLoop parallelODESolving runs solver nodes on
/// task-scheduler threads, Sofa/framework/Simulation/Core/src/sofa/simulation/SolveVisitor.cpp:141-150).
inline sofab200_ctx* threadContext() {
    thread_local sofab200_ctx* ctx = nullptr;
    if (!ctx) { int dev = 0; cudaGetDevice(&dev); sofab200_ctx_create(dev, nullptr, &ctx); }
    return ctx;
}

}  // namespace sofa::b200
