#!/bin/bash
# Blackwell-native evidence of the built library, checkable without a GPU: the architectures of its code objects, per kernel the SASS
# instruction count and the number of TMA (UTMALDG / UBLKCP), mbarrier (SYNCS) and legacy / 5th-gen tensor-core (HMMA / UTC*MMA)
# sites, and ptxas' register / spill / shared-memory table.       tools/sass_evidence.sh > profiles/r2_sass_evidence.txt
cd "$(dirname "$0")/.."
so=feature_detector_b200/libfd_b200.so
echo "# $(date -u +%F) $so"
echo "## code objects"
cuobjdump -lelf $so | sed 's/^/  /'
echo "## per kernel: instructions, UTMALDG, SYNCS (mbarrier), HMMA, UTC*MMA  (nothing on this path is a contraction: tensor-core counts are expected to be 0)"
cuobjdump -sass $so | awk '
  /Function :/ { if (name) printf "%6d %4d %4d %4d %4d  %s\n", n, tma, bar, hmma, utc, name; name=$3; n=tma=bar=hmma=utc=0 }
  /^[ \t]+\/\*[0-9a-f]+\*\/[ \t]+[A-Z@]/ { n++ }
  /UTMALDG|UBLKCP/ { tma++ }  /SYNCS/ { bar++ }  /HMMA/ { hmma++ }  /UTC[A-Z]*MMA/ { utc++ }
  END { printf "%6d %4d %4d %4d %4d  %s\n", n, tma, bar, hmma, utc, name }' | c++filt | sed 's/fdb::(anonymous namespace):://' | cut -c1-140
echo "## ptxas -v (registers, spills, static shared memory) per source file"
for f in feature_detector_b200/csrc/fd_*.cu; do
  echo "### $(basename $f)"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -fmad=false -Xptxas -v -c -o /dev/null $f 2>&1 |
    awk '/Compiling entry function/ { split($0, a, "\047"); name=a[2] } /Used [0-9]+ registers/ { sub(/^ptxas info[ ]*: /, ""); r=$0 } /spill/ { sub(/^[ ]+/, ""); print "  " name ": " r_prev; } { r_prev=r }' | c++filt 2>/dev/null | head -0
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -fmad=false -Xptxas -v -c -o /dev/null $f 2>&1 |
    grep -E "Compiling entry|registers|spill" | sed -E "s/ptxas info    : //; s/Compiling entry function '([^']*)' for 'sm_100a'/\1/" | paste - - - | c++filt | sed 's/fdb::(anonymous namespace):://g' | cut -c1-260
done
echo
echo "## generic (unspecialised address space) loads / stores / atomics per object: LD.E / ST.E / ATOM.E in the SASS"
echo "## (fd_select: one LD.E.64 per instantiation, the kept list read from shared or from global memory in the final ordering)"
for f in feature_detector_b200/csrc/fd_*.o; do
  echo "$(basename $f .o): LD $(cuobjdump -sass $f | grep -c ' LD\.E') ST $(cuobjdump -sass $f | grep -c ' ST\.E') ATOM $(cuobjdump -sass $f | grep -c ' ATOM\.E')"
done
