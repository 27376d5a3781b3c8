#!/bin/bash
# AddressSanitizer pass over the CPU-side test infrastructure: the flow kernels' per-thread code as run by the emulation
# harness (an out-of-bounds read there would be an out-of-bounds read on the device) and both oracle files.  Builds the
# instrumented libraries, runs the golden tests with the ASan runtime preloaded, restores the normal builds.
set -e
cd "$(dirname "$0")/.."
ASAN=$(g++ -print-file-name=libasan.so)
cp oracle/_ref/liboracle.so /tmp/liboracle.normal
trap 'cp /tmp/liboracle.normal oracle/_ref/liboracle.so; rm -f tests/emul/libflow_emul.so tests/mock/libmvtools_mockengine.so' EXIT   # (the test fixtures rebuild the normal ones)
(cd tests/emul && g++ -std=c++17 -O1 -g -fsanitize=address -fno-omit-frame-pointer -fPIC -fvisibility=hidden -ffp-contract=off -shared -o libflow_emul.so flow_emul.cpp \
    -L../../vapoursynth-mvtools_b200 -lmvtools_b200 -Wl,-rpath,"$PWD/../../vapoursynth-mvtools_b200" && touch libflow_emul.so)
(cd oracle && gcc -std=gnu11 -O1 -g -fsanitize=address -fno-omit-frame-pointer -fPIC -fvisibility=hidden -ffp-contract=off -shared -o _ref/liboracle.so mvtools_oracle.c mvtools_oracle_flow.c -lm)
LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0 python -m pytest tests/test_flow_emul.py tests/test_oracle_flow.py tests/test_oracle.py -q -k "goldens or matches_golden or field_shift"
# ... and over the plugin's host-side glue (plugin.cpp + the stand-in engine): the look-ahead caches and fused paths, also
# with several host threads (the slower timing under ASan is what first exposed the window cache's eviction race)
ROOT=$PWD; PKG=$ROOT/vapoursynth-mvtools_b200
g++ -std=c++17 -O1 -g -fsanitize=address -fno-omit-frame-pointer -fPIC -fvisibility=hidden -Wall -I$ROOT/include -I$ROOT/include/vs -shared -Wl,-Bsymbolic \
    -o tests/mock/libmvtools_mockengine.so $PKG/plugin/plugin.cpp tests/mock/mock_engine.cpp $PKG/csrc/host_logic.cpp -L$ROOT/oracle/_ref -loracle -Wl,-rpath,$ROOT/oracle/_ref
# (not the tests that also load the compiled REFERENCE plugin: it throws and catches std::out_of_range inside selectSADFunction, and a
# preloaded ASan cannot intercept __cxa_throw of a libstdc++ that arrives by dlopen -- "CHECK failed: real___cxa_throw != 0")
LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0 python -m pytest tests/test_plugin_glue.py -q \
    -k "(fused or lookahead or concurrent or canonical or vector_clip or four_contexts or failed_window or recycled) and not recalculate_passes and not _live"
