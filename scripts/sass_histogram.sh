#!/bin/bash
# SASS opcode histogram of every kernel object (cuobjdump works without a GPU): the evidence that the quadrature
# kernels use DMMA.8x8x4 / UBLKCP (TMA bulk) / UTMALDG / SYNCS mbarriers.  Writes profiles/sass_<object>.txt.
#     scripts/sass_histogram.sh [round-tag]
cd "$(dirname "$0")/.."
tag=${1:-r02}
for o in quad_mma quad_kernel fast_kernel resonant rel_kernel setup_kernels nhds_kernel multi_gpu; do
  f=alps_b200/csrc/$o.o
  [ -f $f ] || continue
  out=profiles/${tag}_sass_$o.txt
  {
    echo "# cuobjdump -sass $f : opcode histogram (count opcode), all kernels of the object"
    cuobjdump -sass $f | grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?//' \
      | awk '{print $1}' | sed 's/;$//' | sort | uniq -c | sort -rn
    echo "# kernels:"
    cuobjdump -sass $f | grep -E '^\s+Function :' | sed 's/^\s*//'
  } > $out
done
ls -la profiles/${tag}_sass_*.txt
