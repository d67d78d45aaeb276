#!/usr/bin/env bash
# TEST INFRASTRUCTURE (oracle side).  Builds oracle/_ref/libsofa_ref.so from the
# reference's OWN sources where they lie under $SOFA_REF (default /root/reference):
#   Sofa/framework/Type/src/sofa/type/{Mat.cpp,Vec.h,Mat.h,...}
#   Sofa/framework/Helper/src/sofa/helper/{decompose.cpp,decompose.inl,rmath.h}
#   Sofa/framework/Geometry/src/sofa/geometry/{Tetrahedron.h,Hexahedron.h}
# plus oracle/ref_shim.cpp (ours: a C ABI over those templates).  Nothing from the
# reference is copied into the repo: the only outputs are generated config headers
# and the .so, all under oracle/_ref/ (git-ignored; it still ships to the GPU box).
#
# Full SOFA cannot be configured here (Boost/Eigen3/TinyXML2 absent, no network), so
# only this math layer -- which is header/2-file self-contained -- is "the reference
# compiled here".  See DESIGN.md "Oracle".
set -euo pipefail
REF="${SOFA_REF:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
FW="$REF/Sofa/framework"
if [ ! -d "$FW/Type/src/sofa/type" ]; then
  echo "build_ref: reference tree not found at $REF (nothing built)"; exit 0
fi
mkdir -p "$OUT/inc/sofa/config" "$OUT/inc/sofa/type" "$OUT/inc/sofa/helper/logging" "$OUT/inc/sofa/geometry"
gen() { # cmake configure_file emulation: all options off, all @VARS@ -> 0
  sed -E 's/^#cmakedefine01 ([A-Za-z0-9_]+).*/#define \1 0/; s/^#cmakedefine .*/\/\/ &/; s/@[A-Za-z0-9_]+@/0/g' "$1" > "$2"
}
gen "$FW/Config/src/sofa/config.h.in"                           "$OUT/inc/sofa/config.h"
gen "$FW/Config/src/sofa/config/sharedlibrary_defines.h.in"     "$OUT/inc/sofa/config/sharedlibrary_defines.h"
gen "$FW/Config/src/sofa/config/build_option_bbox.h.in"         "$OUT/inc/sofa/config/build_option_bbox.h"
gen "$FW/Config/src/sofa/config/build_option_dump_visitor.h.in" "$OUT/inc/sofa/config/build_option_dump_visitor.h"
gen "$FW/Type/src/sofa/type/config.h.in"                        "$OUT/inc/sofa/type/config.h"
gen "$FW/Helper/src/sofa/helper/config.h.in"                    "$OUT/inc/sofa/helper/config.h"
gen "$FW/Geometry/src/sofa/geometry/config.h.in"                "$OUT/inc/sofa/geometry/config.h"
# logging stub (the real one needs boost::shared_ptr): swallow every message stream
cat > "$OUT/inc/sofa/helper/logging/Messaging.h" <<'EOF'
#pragma once
#include <iosfwd>
namespace sofa_ref_stub { struct Null { template<class T> Null& operator<<(const T&) { return *this; } }; }
#define msg_info(...)      if (true) {} else ::sofa_ref_stub::Null()
#define msg_warning(...)   if (true) {} else ::sofa_ref_stub::Null()
#define msg_error(...)     if (true) {} else ::sofa_ref_stub::Null()
#define msg_deprecated(...) if (true) {} else ::sofa_ref_stub::Null()
#define dmsg_info(...)     if (true) {} else ::sofa_ref_stub::Null()
#define dmsg_warning(...)  if (true) {} else ::sofa_ref_stub::Null()
#define dmsg_error(...)    if (true) {} else ::sofa_ref_stub::Null()
EOF
# SOFA's default Release flags are -O3 -DNDEBUG; no -march, no fast-math
# (Sofa/framework/Config/CMakeLists.txt:62,141-146,163-165).  -ffp-contract=off keeps
# the x86-64 result independent of the host (no FMA contraction), as on a stock build.
g++ -std=c++20 -O3 -DNDEBUG -ffp-contract=off -fPIC -shared \
    -DSOFA_BUILD_HELPER -DSOFA_BUILD_SOFA_TYPE \
    -I"$OUT/inc" -I"$FW/Config/src" -I"$FW/Type/src" -I"$FW/Helper/src" -I"$FW/Geometry/src" \
    "$FW/Type/src/sofa/type/Mat.cpp" "$FW/Helper/src/sofa/helper/decompose.cpp" \
    "$HERE/ref_shim.cpp" -o "$OUT/libsofa_ref.so"
echo "build_ref: built $OUT/libsofa_ref.so"
