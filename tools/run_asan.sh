#!/bin/bash
# Host-side suite under AddressSanitizer + UBSan (reference: the gtest suite is run ASAN-clean in CI).
# usage: tools/run_asan.sh [pytest args]     (defaults to the whole CPU suite)
set -e
cd "$(dirname "$0")/.."
make -j8 asan > /dev/null
GCC=/usr/bin/gcc
rm -f /tmp/ucc_b200_san.*
export LD_PRELOAD=$($GCC -print-file-name=libasan.so):$($GCC -print-file-name=libubsan.so)
export ASAN_OPTIONS=detect_leaks=0:log_path=/tmp/ucc_b200_san UBSAN_OPTIONS=print_stacktrace=1:log_path=/tmp/ucc_b200_san
export UCC_B200_LIB=$PWD/build-asan/lib/libucc.so
if [ $# -eq 0 ]; then set -- tests -q -m "not gpu"; fi
python -m pytest -p no:cacheprovider "$@"
n=$(ls /tmp/ucc_b200_san.* 2>/dev/null | wc -l)
echo "sanitizer reports: $n"
[ "$n" -eq 0 ]
