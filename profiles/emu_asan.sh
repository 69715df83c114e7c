#!/bin/bash
# compute-sanitizer is closed on this GPU pool; the closest available memory check: the product kernels compiled for the CPU by
# the test-only CUDA emulation (tests/emu) with AddressSanitizer + UBSan, run through the emulated parity tests.
# Usage: bash profiles/emu_asan.sh  (writes profiles/r2_emu_asan.log)
set -e
cd "$(dirname "$0")/.."
export PLF_EMU_EXTRA_FLAGS="-fsanitize=address,undefined -fno-omit-frame-pointer"
rm -f tests/emu/libplf_emu.so
ASAN=$(gcc -print-file-name=libasan.so)
LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0:halt_on_error=0:log_path=/tmp/plf_asan UBSAN_OPTIONS=print_stacktrace=1:log_path=/tmp/plf_ubsan \
  python -m pytest tests/test_emu_parity.py -x -q 2>&1 | tail -5 | tee profiles/r2_emu_asan.log
echo "ASan reports: $(ls /tmp/plf_asan.* 2>/dev/null | wc -l), UBSan reports: $(ls /tmp/plf_ubsan.* 2>/dev/null | wc -l)" | tee -a profiles/r2_emu_asan.log
for f in /tmp/plf_asan.* /tmp/plf_ubsan.*; do [ -f "$f" ] && head -40 "$f" >> profiles/r2_emu_asan.log; done
rm -f tests/emu/libplf_emu.so   # the next test run rebuilds the plain emulation
