#!/usr/bin/env bash
# AddressSanitizer + UndefinedBehaviorSanitizer over the host-side System state machine (alvaar_b200/csrc/system_core.h -- the code
# that drives the CUDA kernels in the product) instantiated over the CPU oracle: 100 frames at 640x480 and 40 at 1280x720, each with
# a blackout frame (reset, re-initialisation) and findPlane calls.  Prints the sanitizer's findings, if any.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
B="${TMPDIR:-/tmp}/alva_sanitize"; mkdir -p "$B"
SAN="-O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer"
for f in alva_oracle ba_oracle klt_oracle pose_oracle detect_oracle match_oracle init_oracle; do
  gcc $SAN -ffp-contract=off -std=gnu11 -w -c "$ROOT/oracle/$f.c" -o "$B/$f.o"
done
g++ $SAN -std=c++17 -I"$ROOT/tests/host" -o "$B/sys" "$ROOT/tests/host/system_sanitize_main.cpp" "$B"/*.o -lm
cd "$ROOT"
python -c "
from alvaar_b200 import synth
import sys
for (w,h,nf) in ((640,480,100),(1280,720,40)):
    fr,_=synth.make_frames(nf,w,h,seed=7,rgba=True); fr.tofile('$B/frames_%d.bin'%w)
    K=synth.intrinsics(w,h); open('$B/args_%d.txt'%w,'w').write('%d %d %d %r %r %r %r'%(w,h,nf,K[0],K[1],K[2],K[3]))
"
for w in 640 1280; do "$B/sys" "$B/frames_$w.bin" $(cat "$B/args_$w.txt"); done
echo "sanitizers: clean"
