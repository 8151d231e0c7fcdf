#!/bin/bash
# Builds a tiny consumer against an installed tree twice - through pkg-config and through CMake's find_package(ucc) - and runs it.
set -e
PREFIX=${1:?usage: test_consumer_export.sh <install prefix>}
T=$(mktemp -d)
cat > $T/main.c <<'EOC'
#include <ucc/api/ucc.h>
#include <stdio.h>
int main(void)
{
    ucc_lib_config_h cfg; ucc_lib_h lib; ucc_lib_params_t p = {.mask = UCC_LIB_PARAM_FIELD_THREAD_MODE, .thread_mode = UCC_THREAD_SINGLE};
    if (ucc_lib_config_read(NULL, NULL, &cfg) != UCC_OK) return 1;
    if (ucc_init(&p, cfg, &lib) != UCC_OK) return 2;
    ucc_lib_config_release(cfg);
    printf("consumer ok: %s\n", ucc_get_version_string());
    return ucc_finalize(lib) == UCC_OK ? 0 : 3;
}
EOC
export PKG_CONFIG_PATH=$PREFIX/lib/pkgconfig
cc $T/main.c $(pkg-config --cflags --libs ucc) -Wl,-rpath,$PREFIX/lib -o $T/pc_consumer
UCC_MODULE_DIR=$PREFIX/lib/ucc $T/pc_consumer
if command -v cmake > /dev/null; then
  cat > $T/CMakeLists.txt <<'EOC'
cmake_minimum_required(VERSION 3.16)
project(ucc_consumer C)
find_package(ucc REQUIRED)
add_executable(cm_consumer main.c)
target_link_libraries(cm_consumer ucc::ucc)
EOC
  cmake -S $T -B $T/build -DCMAKE_PREFIX_PATH=$PREFIX > /dev/null
  cmake --build $T/build > /dev/null
  UCC_MODULE_DIR=$PREFIX/lib/ucc LD_LIBRARY_PATH=$PREFIX/lib $T/build/cm_consumer
fi
rm -rf $T
echo CONSUMER_EXPORT_OK
