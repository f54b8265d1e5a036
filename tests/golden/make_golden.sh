#!/bin/bash
# Regenerates tests/golden/*.json from the reference's own test-vector headers.  Needs /root/reference (build
# container only) and oracle/_ref/libecc_ref.so (make -C oracle ref).  The JSON files are committed.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
root="$(cd "$here/../.." && pwd)"
REF="${REF:-/root/reference}"
make -C "$root/oracle" ref >/dev/null
mkdir -p "$root/oracle/_ref"
gcc -O0 -std=gnu11 -w -DWITH_STDLIB -I"$REF/src" "$here/dump_golden.c" -o "$root/oracle/_ref/dump_golden" \
    -L"$root/oracle/_ref" -lecc_ref -Wl,-rpath,"$root/oracle/_ref"
"$root/oracle/_ref/dump_golden" "$here"
gzip -9 -n -f "$here/wycheproof_ecdsa.json" "$here/wycheproof_ecdh.json"
ls -la "$here"/*.json*
