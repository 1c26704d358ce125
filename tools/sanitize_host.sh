#!/bin/bash
# Host code (hx_host.cpp: reader with read-ahead threads, packer, file pipeline) under the sanitizers, no GPU needed:
#   tools/sanitize_host.sh            thread sanitizer, then address + undefined-behaviour sanitizer, on the reader driver;
#                                     then tests/test_host_pipeline.py against an ASAN/UBSAN build of the stub library
set -e
cd "$(dirname "$0")/.."
mkdir -p tests/_build /tmp/hx_sanitize
[ -f tests/_build/hx_oracle.o ] || gcc -O2 -fPIC -std=c11 -pthread -c oracle/hx_oracle.c -o tests/_build/hx_oracle.o
SRC="tools/reader_sanitize_driver.cpp tests/stub/hx_lane_stub.cpp hyperex_b200/csrc/hx_host.cpp tests/_build/hx_oracle.o"
for san in thread address,undefined; do
    g++ -O1 -g -fsanitize=$san -fno-omit-frame-pointer -std=c++17 -I/usr/local/cuda/include -o /tmp/hx_sanitize/drv $SRC -lz -ldl -lpthread
    echo "== reader driver, -fsanitize=$san"
    /tmp/hx_sanitize/drv | tail -3
done
g++ -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -std=c++17 -fPIC -shared -I/usr/local/cuda/include \
    -o tests/_build/libhxhost_stub.so tests/stub/hx_lane_stub.cpp hyperex_b200/csrc/hx_host.cpp tests/_build/hx_oracle.o -lz -ldl -lpthread
echo "== tests/test_host_pipeline.py, ASAN + UBSAN stub library"
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python -m pytest tests/test_host_pipeline.py -x -q | tail -2
rm -f tests/_build/libhxhost_stub.so  # the next test run rebuilds the plain one
