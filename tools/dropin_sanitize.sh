#!/bin/bash
# Host-side sanitizers over the drop-in layer (no GPU): libecc_b200/csrc/dropin.cpp is rebuilt with AddressSanitizer +
# UndefinedBehaviorSanitizer, then with ThreadSanitizer, linked against the host-built engine stub, and the C harness
# (real reference structs, verdicts judged by the unmodified reference) runs its direct / kats / threads / bench modes.
# Needs: python __graft_entry__.py build (harness, reference) and the stub (pytest tests/test_dropin_host.py builds it).
set -u
cd "$(dirname "$0")/.."
STUB=tests/hostsim/_build
H=oracle/_ref/dropin_harness
build() { ( cd libecc_b200/csrc && g++ -O1 -g -std=c++17 -fPIC -shared "$1" -fno-omit-frame-pointer -Wno-unknown-pragmas -x c++ dropin.cpp \
  -o "$2" -L../../$STUB -lecc_b200_stub -Wl,-rpath,"$PWD/../../$STUB" -ldl -lpthread ); }
build -fsanitize=address,undefined /tmp/libecc_b200_dropin_asan.so
build -fsanitize=thread /tmp/libecc_b200_dropin_tsan.so
ASAN="$(g++ -print-file-name=libasan.so) $(g++ -print-file-name=libubsan.so)"
TSAN="$(g++ -print-file-name=libtsan.so)"
run() { echo "== $*"; "$@" 2>&1 | grep -i "HARNESS\|ERROR: AddressSanitizer\|runtime error\|WARNING: ThreadSanitizer" | sort | uniq -c; }
export ASAN_OPTIONS=detect_leaks=0 HARNESS_POOL=64 HARNESS_REPS=2 HARNESS_KATS_THIN=1
HARNESS_CURVES=FRP256V1,SECP521R1 LD_PRELOAD="$ASAN $STUB/libecc_b200_stub.so" run $H direct /tmp/libecc_b200_dropin_asan.so
LD_PRELOAD="$ASAN $STUB/libecc_b200_stub.so" run $H kats /tmp/libecc_b200_dropin_asan.so
LD_PRELOAD="$ASAN $STUB/libecc_b200_stub.so" run $H threads /tmp/libecc_b200_dropin_asan.so
for c in FRP256V1 SECP521R1 SECP224R1; do LD_PRELOAD="$ASAN $STUB/libecc_b200_stub.so" run $H fuzz /tmp/libecc_b200_dropin_asan.so $c 300; done
for sc in ECDSA ECFSDSA BIP0340 ECSDSA ECKCDSA ECGDSA ECRDSA SM2 BIGN; do
  LD_PRELOAD="$ASAN $STUB/libecc_b200_stub.so" run $H bench /tmp/libecc_b200_dropin_asan.so FRP256V1 2600 $sc 64
  LD_PRELOAD="$TSAN $STUB/libecc_b200_stub.so" run $H bench /tmp/libecc_b200_dropin_tsan.so FRP256V1 2600 $sc 64
done
LD_PRELOAD="$TSAN $STUB/libecc_b200_stub.so" run $H threads /tmp/libecc_b200_dropin_tsan.so
