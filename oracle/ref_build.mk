# Builds the reference's own C++ sources (where they lie, read-only, under
# /root/reference) into oracle/_ref/libxara_ref.a.  Test infrastructure only:
# nothing under xara_b200/ links against or loads this archive.
#
# We do NOT run the reference's cmake build; every translation unit in the
# whitelisted directories is compiled stand-alone with g++ using the
# reference's own Linux release flags (CMakeLists.txt:79,93-113), and the ones
# that need an interpreter / MPI / an external solver package simply fail and
# are left out.  The harness (oracle/ref_harness.cpp) then links only the
# archive members it needs.
REF      ?= /root/reference
SRC      := $(REF)/SRC
OUT      := _ref
OBJ      := $(OUT)/obj
TOPDIRS  := matrix utility tagged actor domain element material coordTransformation \
            analysis graph handler system_of_eqn quadrature interpolate damping \
            recorder logging damage
EXCLUDE  := -not -path '*/petsc/*' -not -path '*/mumps/*' -not -path '*/pardiso/*' \
            -not -path '*/itpack/*' -not -path '*/feap/*' -not -iname '*tcl*' \
            -not -iname '*distributed*' -not -iname '*mpi*' -not -iname '*cula*' \
            -not -iname '*cusp*' -not -path '*/umfGEN/*' -not -path '*/sparseSYM/*' \
            -not -path '*/community/*' -not -path '*/IGA/*' -not -path '*/PFEM*' -not -iname 'PFEM*' -not -name 'Test*.cpp'
INCS     := $(shell find $(SRC) -type d | sed 's/^/-I/') -I$(REF)/OTHER/AMD -I$(REF)/OTHER/UMFPACK -I$(REF)/OTHER/CSPARSE
CXXFLAGS := -std=c++17 -O3 -march=haswell -mavx2 -ffloat-store -fPIC -w -D_LINUX -D_UNIX -D_TCL85 $(INCS)

SRCS := $(shell for d in $(TOPDIRS); do find $(SRC)/$$d -name '*.cpp' $(EXCLUDE); done | sort -u)
OBJS := $(patsubst $(SRC)/%.cpp,$(OBJ)/%.o,$(SRCS))

all: $(OUT)/libxara_ref.a

$(OUT)/libxara_ref.a: $(OBJS)
	@rm -f $@
	@find $(OBJ) -name '*.o' | xargs ar rcs $@
	@echo "archive: $$(ar t $@ | wc -l) members"

$(OBJ)/%.o: $(SRC)/%.cpp
	@mkdir -p $(dir $@)
	@$(CXX) $(CXXFLAGS) -c $< -o $@ 2> $@.err || (echo "FAIL $<" >> $(OUT)/failed.txt; echo "" | $(CXX) -x c++ -c - -o $@)

# ---- plain-C closed-form small inverses used by MatrixND (matrix/routines/*.c) ----
CSRCS := $(wildcard $(SRC)/matrix/routines/*.c)
COBJS := $(patsubst $(SRC)/%.c,$(OBJ)/%.o,$(CSRCS))
$(OUT)/libxara_ref.a: $(COBJS)
$(OBJ)/%.o: $(SRC)/%.c
	@mkdir -p $(dir $@)
	@$(CC) -O3 -march=haswell -ffloat-store -fPIC -w -I$(SRC)/matrix/routines -c $< -o $@ 2> $@.err || (echo "" | $(CC) -x c -c - -o $@)

# ---- the reference's vendored METIS 4 (OTHER/METIS, plain C): graph/partitioner/Metis.cpp calls
# METIS_PartGraphKway on the element graph; the partition tests feed its result to xb_setup_partitioned ----
metis: $(OUT)/libmetis_ref.so
$(OUT)/libmetis_ref.so: $(wildcard $(REF)/OTHER/METIS/*.c)
	@mkdir -p $(OUT)
	@echo "build $@"
	@$(CC) -O2 -fPIC -shared -std=gnu89 -w -fcommon -I$(REF)/OTHER/METIS $^ -o $@ -lm

# ---- the harness shared library the tests and the reference bench arm load ----
harness: $(OUT)/libref_harness.so $(OUT)/libmetis_ref.so $(OUT)/libref_glue.so

# ---- the reference-side binding of INTEGRATION.md, built for real: the reference's own Newton loop on top of
# the C ABI (tests only; links the product library, so it is kept apart from the harness the CPU arm loads) ----
XB_DIR := ../xara_b200
$(OUT)/libref_glue.so: ref_glue.cpp ref_harness.cpp ref_shims.cpp ../include/xara_b200.h $(OUT)/libxara_ref.a
	@echo "link $@"
	@$(CXX) -std=c++17 -O2 -fPIC -w -D_LINUX -D_UNIX $(INCS) -shared ref_glue.cpp ref_shims.cpp \
	    -o $@ $(OUT)/libxara_ref.a -L$(XB_DIR) -lxara_b200 -Wl,-rpath,'$$ORIGIN/../../xara_b200' -Wl,--no-undefined
$(OUT)/libref_harness.so: ref_harness.cpp ref_shims.cpp $(OUT)/libxara_ref.a
	@echo "link $@"
	@$(CXX) -std=c++17 -O2 -fPIC -w -D_LINUX -D_UNIX $(INCS) -shared ref_harness.cpp ref_shims.cpp \
	    -o $@ $(OUT)/libxara_ref.a -Wl,--no-undefined
