# Builds libdeepnet_b200.so (sm_100a only) in-tree: deepnet_b200/lib/libdeepnet_b200.so
# Usage: make -j8            (product library)
#        make oracle         (CPU parity oracle, test infrastructure)
NVCC      ?= /usr/local/cuda/bin/nvcc
# The image exports CXX=/opt/gcc/bin/g++ (a trimmed toolchain); use the distro compiler as nvcc's host compiler.
HOSTCXX   := $(shell test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin $(HOSTCXX) --expt-relaxed-constexpr \
             -Xcudafe --diag_suppress=177 -Xptxas -warn-spills -Xfatbin=-compress-all
SRC       := deepnet_b200/csrc
OBJ       := build/obj
LIB       := deepnet_b200/lib/libdeepnet_b200.so

CONVERT_TYPES := f32 f64 i8 u8 i16 u16 i32 u32 i64 u64 bool
CTYPE_f32 := float
CTYPE_f64 := double
CTYPE_i8 := int8_t
CTYPE_u8 := uint8_t
CTYPE_i16 := int16_t
CTYPE_u16 := uint16_t
CTYPE_i32 := int32_t
CTYPE_u32 := uint32_t
CTYPE_i64 := int64_t
CTYPE_u64 := uint64_t
CTYPE_bool := bool8

SOURCES := $(wildcard $(SRC)/*.cu)
OBJECTS := $(patsubst $(SRC)/%.cu,$(OBJ)/%.o,$(SOURCES)) \
           $(foreach t,$(CONVERT_TYPES),$(OBJ)/gen_ew_convert_$(t).o)
HEADERS := $(wildcard $(SRC)/*.cuh) include/dn_tensor.h

all: $(LIB)

$(LIB): $(OBJECTS)
	@mkdir -p $(dir $@)
	$(NVCC) $(ARCH) -shared -ccbin $(HOSTCXX) -o $@ $(OBJECTS) -lcudart

$(OBJ)/%.o: $(SRC)/%.cu $(HEADERS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVCCFLAGS) $(EXTRA_$*) -c $< -o $@

# the fused-expression interpreter must round every instruction like the single operator it stands for
EXTRA_ew_fused := -fmad=false

# one generated translation unit per conversion target type
build/gen/ew_convert_%.cu: $(SRC)/ew_convert.cu.in
	@mkdir -p build/gen
	sed -e 's/@TT@/$*/g' -e 's/@CTYPE@/$(CTYPE_$*)/g' $< > $@

$(OBJ)/gen_ew_convert_%.o: build/gen/ew_convert_%.cu $(HEADERS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVCCFLAGS) -I$(SRC) -c $< -o $@

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIB)

.PHONY: all oracle clean
.SECONDARY:
