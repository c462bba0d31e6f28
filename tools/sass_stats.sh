#!/bin/bash
# Per-kernel SASS instruction counts (code size = 16 B per instruction) of the built library.
cuobjdump -sass "${1:-feature_detector_b200/libfd_b200.so}" | awk '/Function :/{if(name)print n, name; name=$3; n=0} /^[ \t]+\/\*[0-9a-f]+\*\/[ \t]+[A-Z@]/{n++} END{print n, name}' | c++filt | sed 's/fdb::(anonymous namespace):://' | cut -c1-90
