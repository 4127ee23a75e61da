#!/bin/bash
# tuning builds of the tensor-pipe forward: same objects, roi_align.cu recompiled with other constants
cd "$(dirname "$0")/.."
L=htd_b200/_lib; V=$L/variants; mkdir -p $V
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Wno-deprecated-gpu-targets"
others=$(ls $L/*.o | grep -v hooks | grep -v roi_align)
build() { # name tiles desc ahead seg
  nvcc $F -DHTD_MF_TILES=$2 -DHTD_MF_DESC=$3 -DHTD_MF_AHEAD=$4 -DHTD_MF_SEG=$5 -c htd_b200/csrc/roi_align.cu -o $V/roi_$1.o &&
  nvcc -shared -o $V/lib_$1.so $V/roi_$1.o $others -lcudart -lcuda && echo built $1
}
build t5d4a2s2 5 4 2 2 &
build t4d4a2s4 4 4 2 4 &
build t5d4a2s1 5 4 2 1 &
build t5d4a2s4 5 4 2 4 &
build t4d4a2s2 4 4 2 2 &
build t5d4a3s2 5 4 3 2 &
wait
ls -la $V/*.so
