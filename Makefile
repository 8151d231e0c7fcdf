# ucc_b200 build: C core (gcc) + sm_100a CUDA plugin modules (nvcc).
# Everything lands in-tree under ucc_b200/lib so the GPU box sees the same binaries.
#   make            - core library, plugins, tools
#   make core       - libucc.so only (no CUDA toolchain needed)
#   make sass       - SASS / PTX listings of the collective kernels into profiles/
CC      ?= gcc
CXX     ?= g++
NVCC    ?= /usr/local/cuda/bin/nvcc
CUDA_HOME ?= /usr/local/cuda
OUT     := ucc_b200/lib
MODDIR  := $(OUT)/ucc
BINDIR  := ucc_b200/bin
BUILD   := build

CFLAGS  := -O2 -g -std=gnu11 -fPIC -Wall -Wextra -Wno-unused-parameter -Wno-missing-field-initializers -Wno-sign-compare \
           -Iinclude -Isrc -D_GNU_SOURCE $(EXTRA_CFLAGS)
LDFLAGS := -shared -Wl,--no-undefined -lpthread -ldl -lrt -lm
NVFLAGS := -O3 -std=c++17 --extended-lambda -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC \
           -Iinclude -Isrc -D_GNU_SOURCE --expt-relaxed-constexpr $(EXTRA_NVFLAGS)
CUDA_LIBS := -L$(CUDA_HOME)/lib64 -lcudart -Xlinker -rpath,$(CUDA_HOME)/lib64

CORE_DIRS := src/utils src/utils/arch src/utils/profile src/core src/schedule src/coll_score \
             src/components/base src/components/cl src/components/cl/basic src/components/cl/hier \
             src/components/tl src/components/tl/self src/components/tl/shm \
             src/components/mc src/components/mc/cpu src/components/ec src/components/ec/cpu src/components/topo
CORE_SRCS := $(foreach d,$(CORE_DIRS),$(wildcard $(d)/*.c))
CORE_OBJS := $(patsubst %.c,$(BUILD)/%.o,$(CORE_SRCS))

PLUGINS :=
ifneq ($(wildcard src/components/mc/cuda/*.c*),)
PLUGINS += $(MODDIR)/libucc_mc_cuda.so
endif
ifneq ($(wildcard src/components/ec/cuda/*.c*),)
PLUGINS += $(MODDIR)/libucc_ec_cuda.so
endif
ifneq ($(wildcard src/components/topo/cuda/*.c*),)
PLUGINS += $(MODDIR)/libucc_sysinfo_cuda.so
endif
ifneq ($(wildcard src/components/tl/nvl/*.c*),)
PLUGINS += $(MODDIR)/libucc_tl_nvl.so
endif
ifneq ($(wildcard src/components/tl/nccl/*.c*),)
PLUGINS += $(MODDIR)/libucc_tl_nccl.so
endif

TOOLS :=
ifneq ($(wildcard tools/info/*.c),)
TOOLS += $(BINDIR)/ucc_info
endif
ifneq ($(wildcard tools/perf/*.c*),)
TOOLS += $(BINDIR)/ucc_perftest
endif

# host-side algorithm plugins of tl/shm (libucc_tlcp_shm_<name>.so, plain C)
TLCP_SRCS := $(wildcard src/components/tl/shm/coll_plugins/*/*.c)
TLCPS     := $(foreach s,$(TLCP_SRCS),$(MODDIR)/libucc_tlcp_shm_$(notdir $(patsubst %/,%,$(dir $(s)))).so)

.PHONY: all core plugins tools clean sass asan tsan
all: core plugins tools
core: $(OUT)/libucc.so $(TLCPS)
$(MODDIR)/libucc_tlcp_shm_%.so: src/components/tl/shm/coll_plugins/%/*.c $(OUT)/libucc.so
	@mkdir -p $(MODDIR)
	$(CC) $(CFLAGS) -shared -o $@ $(wildcard src/components/tl/shm/coll_plugins/$*/*.c) -L$(OUT) -lucc -Wl,-rpath,'$$ORIGIN/..'
plugins: $(PLUGINS)
tools: $(TOOLS)

$(BUILD)/%.o: %.c
	@mkdir -p $(dir $@)
	$(CC) $(CFLAGS) -MMD -MP -c $< -o $@

$(OUT)/libucc.so: $(CORE_OBJS)
	@mkdir -p $(OUT) $(MODDIR)
	$(CC) -o $@ $(CORE_OBJS) $(LDFLAGS) -Wl,-soname,libucc.so

# ---- CUDA plugin modules: every .c / .cu of the component dir goes into one module ----
define plugin_rule
$(1)_SRCS := $$(wildcard $(2)/*.c) $$(wildcard $(2)/*.cu) $$(wildcard $(2)/kernels/*.cu) src/utils/cuda/ucc_cuda_util.c
$(1)_OBJS := $$(patsubst %,$(BUILD)/%.o,$$($(1)_SRCS))
$(MODDIR)/libucc_$(1).so: $$($(1)_OBJS) $(OUT)/libucc.so
	@mkdir -p $(MODDIR)
	$(NVCC) -shared -o $$@ $$($(1)_OBJS) -L$(OUT) -lucc $(CUDA_LIBS) -Xlinker -rpath,'$$$$ORIGIN/..' -lpthread -ldl $(3)
endef
$(BUILD)/%.cu.o: %.cu
	@mkdir -p $(dir $@)
	$(NVCC) $(NVFLAGS) -MMD -MP -c $< -o $@
$(BUILD)/%.c.o: %.c
	@mkdir -p $(dir $@)
	$(CC) $(CFLAGS) -I$(CUDA_HOME)/include -MMD -MP -c $< -o $@

$(eval $(call plugin_rule,mc_cuda,src/components/mc/cuda,))
$(eval $(call plugin_rule,ec_cuda,src/components/ec/cuda,))
$(eval $(call plugin_rule,sysinfo_cuda,src/components/topo/cuda,))
$(eval $(call plugin_rule,tl_nvl,src/components/tl/nvl,))
$(eval $(call plugin_rule,tl_nccl,src/components/tl/nccl,))

$(BINDIR)/ucc_info: $(wildcard tools/info/*.c) $(OUT)/libucc.so
	@mkdir -p $(BINDIR)
	$(CC) $(CFLAGS) -fvisibility=default -o $@ $(wildcard tools/info/*.c) -L$(OUT) -lucc -Wl,-rpath,'$$ORIGIN/../lib' -ldl -lpthread

$(BINDIR)/ucc_perftest: $(wildcard tools/perf/*.c) $(wildcard tools/perf/*.cc) $(OUT)/libucc.so
	@mkdir -p $(BINDIR)
	$(CXX) -O2 -g -std=c++17 -Iinclude -Isrc -I$(CUDA_HOME)/include -o $@ $(wildcard tools/perf/*.cc) -L$(OUT) -lucc -Wl,-rpath,'$$ORIGIN/../lib' -ldl -lpthread

sass: plugins
	@mkdir -p profiles/sass
	@for m in tl_nvl ec_cuda; do if [ -f $(MODDIR)/libucc_$$m.so ]; then \
	  cuobjdump -sass $(MODDIR)/libucc_$$m.so > profiles/sass/$$m.sass 2>/dev/null; \
	  cuobjdump -ptx $(MODDIR)/libucc_$$m.so > profiles/sass/$$m.ptx 2>/dev/null; fi; done

# AddressSanitizer + UBSan build of the host side (core, tl/shm plugins) into build-asan/; run the suite against it with
#   LD_PRELOAD=$$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 UCC_B200_LIB=build-asan/lib/libucc.so pytest ...
asan:
	$(MAKE) core CC=/usr/bin/gcc BUILD=build-asan/obj OUT=build-asan/lib BINDIR=build-asan/bin \
	  EXTRA_CFLAGS="-fsanitize=address,undefined -fno-sanitize=alignment -fno-omit-frame-pointer -O1" \
	  LDFLAGS="-shared -fsanitize=address,undefined -lpthread -ldl -lrt -lm"

# ThreadSanitizer build (THREAD_MULTIPLE paths): LD_PRELOAD=$$(gcc -print-file-name=libtsan.so) UCC_B200_LIB=build-tsan/lib/libucc.so
tsan:
	$(MAKE) core CC=/usr/bin/gcc BUILD=build-tsan/obj OUT=build-tsan/lib BINDIR=build-tsan/bin \
	  EXTRA_CFLAGS="-fsanitize=thread -fno-omit-frame-pointer -O1" LDFLAGS="-shared -fsanitize=thread -lpthread -ldl -lrt -lm"

# ---- install: headers, libraries, plugin modules, tools, pkg-config file and CMake package (what a consumer of
#      openucx/ucc finds after `make install`: include/ucc/api/*.h, lib/libucc.so, lib/ucc/*.so, lib/pkgconfig/ucc.pc, lib/cmake/ucc) ----
PREFIX  ?= /usr/local
VERSION := $(shell sed -n 's/.*UCC_API_MAJOR *\([0-9][0-9]*\).*/\1/p' include/ucc/api/ucc_version.h | head -1).$(shell sed -n 's/.*UCC_API_MINOR *\([0-9][0-9]*\).*/\1/p' include/ucc/api/ucc_version.h | head -1).0
.PHONY: install
install: core tools
	@mkdir -p $(PREFIX)/include/ucc/api $(PREFIX)/lib/ucc $(PREFIX)/lib/pkgconfig $(PREFIX)/lib/cmake/ucc $(PREFIX)/bin
	cp include/ucc/api/*.h $(PREFIX)/include/ucc/api/
	cp $(OUT)/libucc.so $(PREFIX)/lib/
	@if ls $(MODDIR)/*.so > /dev/null 2>&1; then cp $(MODDIR)/*.so $(PREFIX)/lib/ucc/; fi
	@if ls $(BINDIR)/* > /dev/null 2>&1; then cp $(BINDIR)/* $(PREFIX)/bin/; fi
	sed -e 's|@PREFIX@|$(abspath $(PREFIX))|g' -e 's|@VERSION@|$(VERSION)|g' ucc.pc.in > $(PREFIX)/lib/pkgconfig/ucc.pc
	sed -e 's|@VERSION@|$(VERSION)|g' cmake/ucc-config.cmake.in > $(PREFIX)/lib/cmake/ucc/ucc-config.cmake
	sed -e 's|@VERSION@|$(VERSION)|g' cmake/ucc-config-version.cmake.in > $(PREFIX)/lib/cmake/ucc/ucc-config-version.cmake

clean:
	rm -rf $(BUILD) $(OUT) $(BINDIR) build-asan build-tsan

-include $(CORE_OBJS:.o=.d)
-include $(shell find $(BUILD) -name '*.d' 2>/dev/null)
