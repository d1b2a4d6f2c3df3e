#!/usr/bin/env bash
# Builds the UNMODIFIED-ALGORITHM reference encoder (Nuos/jpgEnc) into oracle/_ref/ so the C oracle
# (oracle/jpgenc_oracle.c) can be pinned against it.  TEST INFRASTRUCTURE ONLY.
#
#   oracle/_ref/jpgEnc_ref          the reference CLI (src/main.cpp)
#   oracle/_ref/libjpgenc_ref.so    the reference library + oracle/ref_probe.cpp (C-ABI stage dump)
#
# The reference is MSVC-2013 + Boost code.  Boost is not in this image, so the build uses the
# from-scratch stand-in headers in oracle/boost_shim/ (ublas::matrix as a dense array; no encode-path
# arithmetic lives in Boost).  Three MSVC-isms are patched with sed on a THROW-AWAY copy under
# $TMPDIR (never committed, never written next to the reference, deleted on exit):
#   1. BitstreamGeneric.hpp: friend templates re-declare the class's own parameter name
#   2. JpegSegments.hpp:     template<int> vs std::array<_, size_t> deduction
#   3. JpegSegments.hpp:     std::array::assign (MSVC-only) -> fill
# None of them changes behaviour.  Only binaries land in oracle/_ref/ (git-ignored).
set -euo pipefail
REF="${JPGENC_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
    echo "build_ref.sh: $REF not present; keeping whatever is already in $OUT" >&2
    exit 0
fi
mkdir -p "$OUT"
WORK="$(mktemp -d)"
trap 'rm -rf "$WORK"' EXIT
mkdir -p "$WORK/include" "$WORK/src"
cp "$REF"/include/*.hpp "$WORK/include/"
cp "$REF"/src/Image.cpp "$REF"/src/Huffman.cpp "$REF"/src/main.cpp "$WORK/src/"
sed -i -E '/^    template<typename BlockType>$/{N;s/BlockType/BlockTypeF/g}' "$WORK/include/BitstreamGeneric.hpp"
sed -i -E 's/template <int T>/template <std::size_t T>/; s/template <int sz>/template <std::size_t sz>/; s/HTinfo\.assign\(/HTinfo.fill(/; s/QT\.QT_info\.assign\(/QT.QT_info.fill(/' "$WORK/include/JpegSegments.hpp"
CXX="${JPGENC_CXX:-/usr/bin/g++}"; [ -x "$CXX" ] || CXX=g++
# -O2, no -march: keeps double arithmetic un-contracted (no FMA on baseline x86-64), like the survey's pins
FLAGS=(-std=c++20 -O2 -fopenmp -w -include algorithm -include cmath -include functional -include array
       -I"$HERE/boost_shim" -I"$WORK/include")
"$CXX" "${FLAGS[@]}" -c "$WORK/src/Image.cpp" -fPIC -o "$WORK/Image.o"
"$CXX" "${FLAGS[@]}" -c "$WORK/src/Huffman.cpp" -fPIC -o "$WORK/Huffman.o"
"$CXX" "${FLAGS[@]}" -c "$WORK/src/main.cpp" -o "$WORK/main.o"
"$CXX" "${FLAGS[@]}" -fno-access-control -c "$HERE/ref_probe.cpp" -fPIC -o "$WORK/probe.o"
"$CXX" -fopenmp -o "$OUT/jpgEnc_ref" "$WORK/main.o" "$WORK/Image.o" "$WORK/Huffman.o" -lpthread
"$CXX" -fopenmp -shared -o "$OUT/libjpgenc_ref.so" "$WORK/probe.o" "$WORK/Image.o" "$WORK/Huffman.o" -lpthread
echo "build_ref.sh: built $OUT/jpgEnc_ref and $OUT/libjpgenc_ref.so"
