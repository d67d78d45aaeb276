#!/usr/bin/env bash
# Syntax-checks every file of sofa_b200/plugin against the reference's own headers where they lie under $SOFA_REF (default /root/reference):
# g++ -fsyntax-only with include paths to every SOFA module's src directory, the config headers cmake would generate (same emulation as
# oracle/build_ref.sh) and three small stand-ins for the Boost headers SOFA's core includes (Boost is not in this image).  SOFA itself cannot be
# built here; this keeps the glue from rotting: every virtual the glue overrides, every Data it reads and every C-ABI call it makes must exist
# with the right signature.  Output: one line per file; exit code 1 if any file fails.
set -uo pipefail
REF="${SOFA_REF:-/root/reference}"
HERE="$(cd "$(dirname "$0")/.." && pwd)"
OUT="${PLUGIN_CHECK_DIR:-/tmp/sofa_b200_plugin_check}"
if [ ! -d "$REF/Sofa/framework" ]; then echo "plugin_syntax_check: reference tree not found at $REF"; exit 2; fi
mkdir -p "$OUT/gen" "$OUT/stub/boost/container"
gen() { mkdir -p "$(dirname "$2")"; sed -E 's/^#cmakedefine01 ([A-Za-z0-9_]+).*/#define \1 0/; s/^#cmakedefine .*/\/\/ &/; s/@[A-Za-z0-9_]+@/0/g' "$1" > "$2"; }
INC=()
while IFS= read -r srcdir; do INC+=("-I$srcdir"); done < <(find "$REF/Sofa" -type d -name src -not -path "*/extlibs/*")
while IFS= read -r f; do
  rel="${f#*/src/}"; gen "$f" "$OUT/gen/${rel%.in}"
done < <(find "$REF/Sofa" -name "*.h.in" -path "*/src/*")
cat > "$OUT/stub/boost/intrusive_ptr.hpp" <<'EOS'
#pragma once
#include <cstddef>
#include <functional>
namespace boost {
template <class T> class intrusive_ptr {
public:
    typedef T element_type;
    intrusive_ptr() : p(nullptr) {}
    intrusive_ptr(std::nullptr_t) : p(nullptr) {}
    intrusive_ptr(T* q, bool add = true) : p(q) { if (p && add) intrusive_ptr_add_ref(p); }
    intrusive_ptr(const intrusive_ptr& o) : p(o.p) { if (p) intrusive_ptr_add_ref(p); }
    template <class U> intrusive_ptr(const intrusive_ptr<U>& o) : p(o.get()) { if (p) intrusive_ptr_add_ref(p); }
    ~intrusive_ptr() { if (p) intrusive_ptr_release(p); }
    intrusive_ptr& operator=(const intrusive_ptr& o) { intrusive_ptr(o).swap(*this); return *this; }
    intrusive_ptr& operator=(T* q) { intrusive_ptr(q).swap(*this); return *this; }
    void reset() { intrusive_ptr().swap(*this); }
    void reset(T* q) { intrusive_ptr(q).swap(*this); }
    T* get() const { return p; }
    T& operator*() const { return *p; }
    T* operator->() const { return p; }
    explicit operator bool() const { return p != nullptr; }
    void swap(intrusive_ptr& o) { T* t = p; p = o.p; o.p = t; }
private:
    T* p;
};
template <class T, class U> bool operator==(const intrusive_ptr<T>& a, const intrusive_ptr<U>& b) { return a.get() == b.get(); }
template <class T, class U> bool operator!=(const intrusive_ptr<T>& a, const intrusive_ptr<U>& b) { return a.get() != b.get(); }
template <class T, class U> bool operator==(const intrusive_ptr<T>& a, U* b) { return a.get() == b; }
template <class T, class U> bool operator!=(const intrusive_ptr<T>& a, U* b) { return a.get() != b; }
template <class T, class U> bool operator==(T* a, const intrusive_ptr<U>& b) { return a == b.get(); }
template <class T, class U> bool operator!=(T* a, const intrusive_ptr<U>& b) { return a != b.get(); }
template <class T> bool operator==(const intrusive_ptr<T>& a, std::nullptr_t) { return a.get() == nullptr; }
template <class T> bool operator!=(const intrusive_ptr<T>& a, std::nullptr_t) { return a.get() != nullptr; }
template <class T> bool operator<(const intrusive_ptr<T>& a, const intrusive_ptr<T>& b) { return std::less<T*>()(a.get(), b.get()); }
template <class T> T* get_pointer(const intrusive_ptr<T>& p) { return p.get(); }
template <class T, class U> intrusive_ptr<T> static_pointer_cast(const intrusive_ptr<U>& p) { return static_cast<T*>(p.get()); }
template <class T, class U> intrusive_ptr<T> dynamic_pointer_cast(const intrusive_ptr<U>& p) { return dynamic_cast<T*>(p.get()); }
template <class T, class U> intrusive_ptr<T> const_pointer_cast(const intrusive_ptr<U>& p) { return const_cast<T*>(p.get()); }
}  // namespace boost
namespace std { template <class T> struct hash<boost::intrusive_ptr<T>> { size_t operator()(const boost::intrusive_ptr<T>& p) const { return hash<T*>()(p.get()); } }; }
EOS
cat > "$OUT/stub/boost/shared_ptr.hpp" <<'EOS'
#pragma once
#include <memory>
namespace boost { template <class T> using shared_ptr = std::shared_ptr<T>; }
EOS
cat > "$OUT/stub/boost/container/stable_vector.hpp" <<'EOS'
#pragma once
#include <deque>
namespace boost { namespace container { template <class T, class A = std::allocator<T>> using stable_vector = std::deque<T, A>; } }
EOS
rc=0
for f in "$HERE"/sofa_b200/plugin/${1:-*}.cpp; do
  log="$OUT/$(basename "$f").log"
  if grep -q "NEEDS_EIGEN" "$f" && ! echo '#include <Eigen/Sparse>' | g++ -x c++ -fsyntax-only - > /dev/null 2>&1; then echo "SKIP $(basename "$f") (needs Eigen, absent from this image)"; continue; fi
  if g++ -std=c++20 -fsyntax-only -w -DSOFA_BUILD_SOFAB200 -I"$OUT/gen" -I"$OUT/stub" -I"$HERE/include" -I"$HERE/sofa_b200/plugin" -I/usr/local/cuda/include "${INC[@]}" "$f" > "$log" 2>&1; then
    echo "OK   $(basename "$f")"
  else
    echo "FAIL $(basename "$f") ($(grep -c 'error' "$log") errors, first: $(grep -m1 'error' "$log" | cut -c1-220))"; rc=1
  fi
done
exit $rc
