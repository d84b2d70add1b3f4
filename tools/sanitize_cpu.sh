#!/bin/bash
# AddressSanitizer + UndefinedBehaviorSanitizer run of everything that executes on the CPU (memory / UB
# checking): the oracle (oracle/_san/liboracle_san.so) and the HOST side of the product library (partitioner, parameter
# parsers, struct helpers, ABI checks; build/san/libpnpb200.so) under the whole `-m "not gpu"` suite, the world-2 gloo
# processes included (they inherit the environment). Needs the distribution's g++ (libasan); nothing here runs on a GPU.
set -e
cd "$(dirname "$0")/.."
make -s -C oracle san CXX=/usr/bin/g++
make -s -j8 -C modular-pnp_b200 BUILD=build/san LIB=build/san/libpnpb200.so \
     HOSTX="-fsanitize=address -fsanitize=undefined -fno-omit-frame-pointer" EXTRA="-ccbin /usr/bin/g++" build/san/libpnpb200.so
# libstdc++ is preloaded next to libasan: python itself does not link it, and ASan's __cxa_throw interceptor must
# resolve before the first C++ exception of the library (the ERROR_DEVICE path of test_no_cpu_fallback)
PRE="$(/usr/bin/gcc -print-file-name=libasan.so) $(/usr/bin/gcc -print-file-name=libstdc++.so.6)"
PNPB200_LIB=$PWD/modular-pnp_b200/build/san/libpnpb200.so PNP_ORACLE_LIB=$PWD/oracle/_san/liboracle_san.so LD_PRELOAD="$PRE" \
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1:protect_shadow_gap=0 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
python -m pytest tests -q -m "not gpu" -x
rm -rf modular-pnp_b200/build/san        # the variant library must not travel to the GPU box
