#!/bin/bash
# ThreadSanitizer pass over the plugin's host-side glue: the product's plugin.cpp + the stand-in engine of tests/mock built with
# -fsanitize=thread, the glue tests (window state machine, cooperative packing, several contexts / devices, eviction, the injected
# window failure) run with the TSan runtime preloaded.  Children of the subprocess tests inherit it (a report makes them exit 66).
# Not the tests that also load the compiled REFERENCE plugin (uninstrumented; its own races are not the subject).
set -e
cd "$(dirname "$0")/.."
TSAN=$(g++ -print-file-name=libtsan.so)
ROOT=$PWD; PKG=$ROOT/vapoursynth-mvtools_b200
trap 'rm -f tests/mock/libmvtools_mockengine.so' EXIT   # (the test fixture rebuilds the normal one)
g++ -std=c++17 -O1 -g -fsanitize=thread -fno-omit-frame-pointer -fPIC -fvisibility=hidden -Wall -I$ROOT/include -I$ROOT/include/vs -shared -Wl,-Bsymbolic \
    -o tests/mock/libmvtools_mockengine.so $PKG/plugin/plugin.cpp tests/mock/mock_engine.cpp $PKG/csrc/host_logic.cpp -L$ROOT/oracle/_ref -loracle -Wl,-rpath,$ROOT/oracle/_ref
LD_PRELOAD=$TSAN TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" python -u -m pytest tests/test_plugin_glue.py -q -s \
    -k "(fused or lookahead or concurrent or canonical or vector_clip or four_contexts or failed_window or recycled) and not recalculate_passes and not _live" 2>&1 | tee /tmp/tsan_glue.log | tail -3
echo "ThreadSanitizer warnings: $(grep -c 'WARNING: ThreadSanitizer' /tmp/tsan_glue.log || true)"
