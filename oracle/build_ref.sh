#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Compiles the UNMODIFIED reference renderer, from the sources where they lie under
# /root/reference/SourceCode, into oracle/_ref/ (git-ignored; travels to the GPU box like any built artefact).
# The reference's own CMake build cannot be used offline: cmake/FindRapidJSON.cmake:3-17 and cmake/FindSTB.cmake:3-14
# git-clone at configure time.  We therefore compile its 13 translation units directly with its Release flags
# (CMakeLists.txt:5,12: C++20, -DMEASURE_TIME), baseline x86-64 (no -march=native => no FMA contraction, which is
# a parity variable), plus our driver oracle/ref_driver.cpp.
#   rapidjson: third-party, not vendored by the reference (GIT_TAG master, unpinned).  A header-only v1.1.0 copy
#              ships inside this image's site-packages (tilelang/3rdparty/composable_kernel/include).
#   stb_image: vendored by the reference (SourceCode/external/stb_image.h).
# Two flavours, because USE_TEXTURES changes struct layouts (Vertex.h:10-19, Material.h:10-25):
#   oracle/_ref/crt_ref      plain           (configs 1, 2, 4, 5)
#   oracle/_ref/crt_ref_tex  -DUSE_TEXTURES=1 (config 3)
# and INTEGRATION.md section B compiled for real (oracle/ref_b200_binding.cpp):
#   oracle/_ref/crt_ref_b200 the reference's loader, tree build and PPM writer around crtb200_render (needs libcrtb200.so)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${CRT_REFERENCE_ROOT:-/root/reference}/SourceCode"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_ref: $REF not present (GPU box?) -- keeping prebuilt binaries in $OUT" >&2
  exit 0
fi
RJ="$(python3 - <<'PY'
import importlib.util, os, sys
cands = []
spec = importlib.util.find_spec("tilelang")
if spec and spec.submodule_search_locations:
    cands.append(os.path.join(list(spec.submodule_search_locations)[0], "3rdparty/composable_kernel/include"))
for c in cands:
    if os.path.exists(os.path.join(c, "rapidjson/document.h")):
        print(c); sys.exit(0)
sys.exit(1)
PY
)"
mkdir -p "$OUT"
SRCS=$(ls "$REF"/src/*.cpp)
WRAP="-Wl,--wrap=_ZNK6KDTreeI19ObjectKDTreeSubTreeE9intersectERK3Ray -Wl,--wrap=_ZNK12ObjectKDTree20checkForIntersectionERK3Rayfb"
COMMON="-std=gnu++20 -O3 -DNDEBUG -DMEASURE_TIME -ffp-contract=off -w -include $HERE/ref_shim.h -I$REF/include -I$REF/external -I$RJ"
g++ $COMMON                  "$HERE/ref_driver.cpp" $SRCS $WRAP -lpthread -o "$OUT/crt_ref" &
g++ $COMMON -DUSE_TEXTURES=1 "$HERE/ref_driver.cpp" $SRCS $WRAP -lpthread -o "$OUT/crt_ref_tex" &
CSRC="$HERE/../course-assignment-danielhalachev_b200/csrc"
if [ -f "$CSRC/libcrtb200.so" ]; then
  g++ $COMMON "$HERE/ref_b200_binding.cpp" $SRCS -L"$CSRC" -lcrtb200 -Wl,-rpath,'$ORIGIN/../../course-assignment-danielhalachev_b200/csrc' \
      -lpthread -o "$OUT/crt_ref_b200" &
fi
wait
echo "build_ref: built $OUT/crt_ref, $OUT/crt_ref_tex$([ -f "$OUT/crt_ref_b200" ] && echo ", $OUT/crt_ref_b200")"
