// SOFA-side glue (compiles only inside a SOFA build tree).
// Device DataTypes "B200Vec3f" / "B200Vec3d": same Coord/Deriv/Real as Vec3Types, vectors are
// sofa::type::vector_device with a memory manager that allocates through the CUDA runtime, so that every
// Data<VecCoord>/Data<VecDeriv> of a MechanicalObject<B200Vec3fTypes> owns a lazily synchronised host/device pair
// (Sofa/framework/Type/src/sofa/type/vector_device.h:50-155).  Kernels only ever see the raw device pointer
// obtained with deviceRead()/deviceWrite() at call time; the AoS Vec3 layout is the one libsofa_b200 expects.
#pragma once
#include <cuda_runtime.h>
#include <sofa/defaulttype/DataTypeInfo.h>
#include <sofa/defaulttype/VecTypes.h>
#include <sofa/helper/MemoryManager.h>
#include <sofa/type/vector_device.h>

#include <sofa_b200.h>

namespace sofa::b200 {

/// MemoryManager concept of vector_device (the role CudaMemoryManager plays for SofaCUDA,
/// applications/plugins/SofaCUDA/Core/src/sofa/gpu/cuda/CudaMemoryManager.h:39-170)
template <class T> class B200MemoryManager : public sofa::helper::MemoryManager<T> {
public:
    typedef T* host_pointer;
    typedef void* device_pointer;
    typedef unsigned int buffer_id_type;
    template <class T2> struct rebind { typedef B200MemoryManager<T2> other; };
    enum { MAX_DEVICES = 8, BSIZE = 64, SUPPORT_GL_BUFFER = 0 };
    static int numDevices() { int n = 0; cudaGetDeviceCount(&n); return n; }
    static void hostAlloc(void** p, int n) { cudaMallocHost(p, n); }            // pinned: H2D/D2H of state vectors are async
    static void memsetHost(host_pointer p, int v, size_t n) { memset(static_cast<void*>(p), v, n); }
    static void hostFree(const host_pointer p) { cudaFreeHost(p); }
    static void deviceAlloc(int d, device_pointer* p, int n) { cudaSetDevice(d); cudaMalloc(p, n); }
    static void deviceFree(int d, const device_pointer p) { cudaSetDevice(d); cudaFree(p); }
    static void memcpyHostToDevice(int d, device_pointer dst, const void* src, size_t n) { cudaSetDevice(d); cudaMemcpy(dst, src, n, cudaMemcpyHostToDevice); }
    static void memcpyDeviceToHost(int d, void* dst, const void* src, size_t n) { cudaSetDevice(d); cudaMemcpy(dst, src, n, cudaMemcpyDeviceToHost); }
    static void memcpyDeviceToDevice(int d, device_pointer dst, const void* src, size_t n) { cudaSetDevice(d); cudaMemcpy(dst, src, n, cudaMemcpyDeviceToDevice); }
    static void memsetDevice(int d, device_pointer dst, int v, size_t n) { cudaSetDevice(d); cudaMemset(dst, v, n); }
    static int getBufferDevice() { int d = 0; cudaGetDevice(&d); return d; }
    static bool bufferAlloc(buffer_id_type*, int, bool = true) { return false; }
    static void bufferFree(const buffer_id_type) {}
    static bool bufferRegister(const buffer_id_type) { return false; }
    static void bufferUnregister(const buffer_id_type) {}
    static bool bufferMapToDevice(device_pointer*, const buffer_id_type) { return false; }
    static void bufferUnmapToDevice(device_pointer*, const buffer_id_type) {}
    static device_pointer deviceOffset(device_pointer p, size_t off) { return static_cast<char*>(p) + off; }
    static device_pointer null() { return nullptr; }
    static bool isNull(device_pointer p) { return p == nullptr; }
};

/// third template argument of vector_device (the role of SofaCUDA's DataTypeInfoManager, CudaTypes.h:56-61)
template <class T> struct B200DataTypeInfoManager {
    static const bool ZeroConstructor = sofa::defaulttype::DataTypeInfo<T>::ZeroConstructor;
    static const bool SimpleCopy = sofa::defaulttype::DataTypeInfo<T>::SimpleCopy;
};
template <class T> using B200Vector = sofa::type::vector_device<T, B200MemoryManager<T>, B200DataTypeInfoManager<T>>;

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

/// raw device pointers of a Data<VecCoord/VecDeriv/VecReal> (the transfer, if the host copy is the valid one, happens inside vector_device)
template <class V> inline const void* devRead(const V& v) { return v.size() ? v.deviceRead() : nullptr; }
template <class V> inline void* devWrite(V& v) { return v.size() ? v.deviceWrite() : nullptr; }

/// One libsofa_b200 context per simulation thread (DefaultAnimationLoop parallelODESolving runs solver nodes on
/// task-scheduler threads, Sofa/framework/Simulation/Core/src/sofa/simulation/SolveVisitor.cpp:141-150).
inline sofab200_ctx* threadContext() {
    thread_local sofab200_ctx* ctx = nullptr;
    if (!ctx) { int dev = 0; cudaGetDevice(&dev); sofab200_ctx_create(dev, nullptr, &ctx); }
    return ctx;
}

}  // namespace sofa::b200

// ---- the trait specialisations a device vector needs to live inside Data<> (what SofaCUDA provides for CudaVector in
// applications/plugins/SofaCUDA/Core/src/sofa/gpu/cuda/CudaTypes.h:878-1040): host-side accessors go through hostRead()/hostWrite() so that
// the lazily synchronised pair stays coherent, Data<> serialisation sees an ordinary vector, and the mass components find their MassType.
#include <sofa/helper/accessor.h>
namespace sofa::helper {
template <class T> class ReadAccessorVector<sofa::b200::B200Vector<T>> {
public:
    typedef sofa::b200::B200Vector<T> container_type;
    typedef typename container_type::Size Size;
    typedef typename container_type::value_type value_type;
    typedef typename container_type::reference reference;
    typedef typename container_type::const_reference const_reference;
    typedef typename container_type::iterator iterator;
    typedef typename container_type::const_iterator const_iterator;
    ReadAccessorVector(const container_type& c) : m_ref(c), m_host(c.hostRead()) {}
    Size size() const { return m_ref.size(); }
    bool empty() const { return m_ref.empty(); }
    const container_type& ref() const { return m_ref; }
    const_reference operator[](Size i) const { return m_host[i]; }
    const_iterator begin() const { return m_host; }
    const_iterator end() const { return m_host + m_ref.size(); }
protected:
    const container_type& m_ref;
    const value_type* m_host;
};
template <class T> class WriteAccessorVector<sofa::b200::B200Vector<T>> {
public:
    typedef sofa::b200::B200Vector<T> container_type;
    typedef typename container_type::Size Size;
    typedef typename container_type::value_type value_type;
    typedef typename container_type::reference reference;
    typedef typename container_type::const_reference const_reference;
    typedef typename container_type::iterator iterator;
    typedef typename container_type::const_iterator const_iterator;
    WriteAccessorVector(container_type& c) : m_ref(c), m_host(c.hostWrite()) {}
    bool empty() const { return m_ref.empty(); }
    Size size() const { return m_ref.size(); }
    void reserve(Size n) { m_ref.reserve(n); m_host = m_ref.hostWrite(); }
    const_reference operator[](Size i) const { return m_host[i]; }
    reference operator[](Size i) { return m_host[i]; }
    const container_type& ref() const { return m_ref; }
    container_type& wref() { return m_ref; }
    const_iterator begin() const { return m_host; }
    iterator begin() { return m_host; }
    const_iterator end() const { return m_host + m_ref.size(); }
    iterator end() { return m_host + m_ref.size(); }
    void clear() { m_ref.clear(); }
    void resize(Size n, bool init = true) { if (init) m_ref.resize(n); else m_ref.fastResize(n); m_host = m_ref.hostWrite(); }
    iterator erase(iterator pos) { iterator it = m_ref.erase(pos); m_host = m_ref.hostWrite(); return it; }
    void push_back(const_reference v) { m_ref.push_back(v); m_host = m_ref.hostWrite(); }
    void pop_back() { m_ref.pop_back(); m_host = m_ref.hostWrite(); }
protected:
    container_type& m_ref;
    T* m_host;
};
}  // namespace sofa::helper
namespace sofa::defaulttype {
template <class T> struct DataTypeInfo<sofa::b200::B200Vector<T>> : public VectorTypeInfo<sofa::b200::B200Vector<T>> {
    static std::string name() { return std::string("B200Vector<") + DataTypeName<T>::name() + ">"; }
};
}  // namespace sofa::defaulttype
#include <sofa/component/mass/MassType.h>
namespace sofa::component::mass {
template <class TReal> struct MassType<sofa::b200::B200Vec3Types<TReal>> { using type = TReal; };
}  // namespace sofa::component::mass
