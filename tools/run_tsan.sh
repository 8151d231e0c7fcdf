#!/bin/bash
# THREAD_MULTIPLE tests under ThreadSanitizer.  Known false positives: memcpy "races" between rndv_fetch / send_push and the
# producer of the buffer - the ring that orders them is a POSIX shm segment mapped twice in the single-process harness,
# so TSAN does not connect the release (through one mapping) with the acquire (through the other).
set -e
cd "$(dirname "$0")/.."
make -j8 tsan > /dev/null 2>&1
rm -f /tmp/ucc_b200_tsan.*
LD_PRELOAD=$(/usr/bin/gcc -print-file-name=libtsan.so) TSAN_OPTIONS=log_path=/tmp/ucc_b200_tsan:halt_on_error=0:report_signal_unsafe=0 \
  UCC_B200_LIB=$PWD/build-tsan/lib/libucc.so python -m pytest -p no:cacheprovider -q tests/test_core.py -k "shared_context or concurrent_progress" "$@" | tee /tmp/ucc_b200_tsan_pytest.log || true   # TSAN makes python exit 66 when it reported anything
grep -q " passed" /tmp/ucc_b200_tsan_pytest.log && ! grep -q " failed" /tmp/ucc_b200_tsan_pytest.log
real=$(cat /tmp/ucc_b200_tsan.* 2>/dev/null | grep SUMMARY | grep -v "in memcpy" | wc -l)
echo "thread sanitizer reports (excluding double-mapped-ring memcpy false positives): $real"
[ "$real" -eq 0 ]
