#!/bin/bash
# Memory-safety check of the kernel sources without a GPU (compute-sanitizer is closed on the GPU pool):
# builds the debug emulator with AddressSanitizer and runs every emulator test under it.  numpy's buffers
# come from ASan's allocator (LD_PRELOAD), shared memory is a heap block, so out-of-bounds global and shared
# accesses of the kernels trap.  Round 1 (final build): 218 passed, no reports.
set -e
cd "$(dirname "$0")/.."
CS=medical-vision-textural-bias_b200/mvtb/csrc
mkdir -p tests/cuemu/_build
cp tests/cuemu/_build/libmvtb_emu.so /tmp/libmvtb_emu.orig.so 2>/dev/null || true
g++ -std=c++17 -O1 -g -fsanitize=address -fno-omit-frame-pointer -shared -fPIC -DMVTB_EMU -Itests/cuemu -x c++ \
    $CS/plan.cu $CS/kspace_chain.cu $CS/voxel_ops.cu $CS/bandlimited.cu $CS/spike_fast.cu tests/cuemu/cuemu.cpp -o tests/cuemu/_build/libmvtb_emu.so
ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:verify_asan_link_order=0 \
LD_PRELOAD=$(gcc -print-file-name=libasan.so) \
    python -m pytest tests/test_emu_bandlimited.py tests/test_emu_inplace_sp.py tests/test_emu_kernels.py -x -q -p no:cacheprovider
rm -f tests/cuemu/_build/libmvtb_emu.so
[ -f /tmp/libmvtb_emu.orig.so ] && cp /tmp/libmvtb_emu.orig.so tests/cuemu/_build/libmvtb_emu.so
