# Makefile — builds the B200 D2Q9-BGK solver behind the reference's make surface
# (ag14774/OpenCL-Lattice-Boltzmann Makefile:1-32): `make` -> ./d2q9-bgk (and the
# README's ./d2q9-bgk.exe), `make check` with the four *_FILE variables, `make clean`.
#
#   make                 library (nvcc, sm_100a) + C host + oracle
#   make run DECK=128x128            run a shipped deck from decks/
#   make check [REF_AV_VELS_FILE=check/128x256.av_vels.dat REF_FINAL_STATE_FILE=...]
#   make check-all       run + check the four decks
EXE  = d2q9-bgk
PKG  = opencl-lattice-boltzmann_b200
LIB  = $(PKG)/liblbm_b200.so

# the image exports CC=/opt/gcc/bin/gcc (a wrapper without OpenMP specs); use the system gcc
HOSTCC   ?= /usr/bin/gcc
NVCC     ?= nvcc
PYTHON   ?= python
CFLAGS    = -std=c99 -Wall -O3 -fopenmp -D_DEFAULT_SOURCE
NVCCFLAGS = -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall
LIBS      = -lm

FINAL_STATE_FILE=./final_state.dat
AV_VELS_FILE=./av_vels.dat
REF_FINAL_STATE_FILE=check/128x128.final_state.dat
REF_AV_VELS_FILE=check/128x128.av_vels.dat
DECK=128x128

all: $(LIB) $(EXE) $(EXE).exe oracle

$(LIB): $(PKG)/csrc/lbm_cuda.cu $(PKG)/csrc/lbm_kernels.cuh $(PKG)/csrc/lbm_fuse2.cuh $(PKG)/csrc/lbm_fuse2p.cuh $(PKG)/csrc/lbm_fuse2q.cuh $(PKG)/csrc/lbm_tile.cuh include/lbm.h
	$(NVCC) $(NVCCFLAGS) -shared -Iinclude $< -o $@

$(EXE): $(PKG)/host/d2q9_bgk_main.c $(PKG)/host/lbm_io.c $(PKG)/host/lbm_io.h include/lbm.h $(LIB)
	$(HOSTCC) $(CFLAGS) -Iinclude -I$(PKG)/host $(PKG)/host/d2q9_bgk_main.c $(PKG)/host/lbm_io.c \
	    -L$(PKG) -llbm_b200 -Wl,-rpath,'$$ORIGIN/$(PKG)' $(LIBS) -o $@

$(EXE).exe: $(EXE)
	ln -sf $(EXE) $@

oracle:
	$(MAKE) -C oracle --no-print-directory

run: $(EXE)
	./$(EXE) decks/input_$(DECK).params decks/obstacles_$(DECK).dat

# VELOCITY_TOLERANCE=<percent> (extension, off by default): also check the u_x, u_y, |u| columns
check:
	$(PYTHON) check/check.py --ref-av-vels-file=$(REF_AV_VELS_FILE) --ref-final-state-file=$(REF_FINAL_STATE_FILE) --av-vels-file=$(AV_VELS_FILE) --final-state-file=$(FINAL_STATE_FILE) $(if $(VELOCITY_TOLERANCE),--velocity-tolerance $(VELOCITY_TOLERANCE))

# the 1024x1024 final_state golden is committed as its checked column only (90 MB as text)
check/1024x1024.final_state.dat: check/1024x1024.final_state.pressure.npz
	$(PYTHON) check/regenerate_missing_goldens.py --expand 1024x1024

check-all: $(EXE) check/1024x1024.final_state.dat
	@for d in 128x128 128x256 256x256 1024x1024; do \
	  echo "== $$d"; ./$(EXE) decks/input_$$d.params decks/obstacles_$$d.dat || exit 1; \
	  $(PYTHON) check/check.py --ref-av-vels-file=check/$$d.av_vels.dat --ref-final-state-file=check/$$d.final_state.dat \
	    --av-vels-file=$(AV_VELS_FILE) --final-state-file=$(FINAL_STATE_FILE) || exit 1; done

.PHONY: all check check-all clean oracle run

clean:
	rm -f $(EXE) $(EXE).exe $(LIB) final_state.dat av_vels.dat
	$(MAKE) -C oracle --no-print-directory clean
