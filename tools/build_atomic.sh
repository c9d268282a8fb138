#!/bin/bash
# Build both libraries into a staging directory and move them into the package in one step, so that a repository snapshot
# taken at any moment (gpurun) never sees a half-written .so.
set -e
ST=${ST:-/tmp/f5stage}
mkdir -p $ST
make -C f5_tts_b200/csrc -j4 OUT=$ST/libf5tts_b200.so BUILD=$ST/build > $ST/build.log 2>&1 || { grep -E "error" $ST/build.log | head; exit 1; }
make -C f5_tts_b200/csrc -j4 TRACE=1 OUT=$ST/libf5tts_b200_trace.so BUILD=$ST/build_trace > $ST/build_trace.log 2>&1 || { grep -E "error" $ST/build_trace.log | head; exit 1; }
grep -E "warning" $ST/build.log | head -5 || true
cp $ST/libf5tts_b200.so f5_tts_b200/.libf5tts_b200.so.new && mv f5_tts_b200/.libf5tts_b200.so.new f5_tts_b200/libf5tts_b200.so
cp $ST/libf5tts_b200_trace.so f5_tts_b200/.libf5tts_b200_trace.so.new && mv f5_tts_b200/.libf5tts_b200_trace.so.new f5_tts_b200/libf5tts_b200_trace.so
echo "built $(date)"
