# Convenience targets; the driver-facing entry points are __graft_entry__.py (build, smoke) and bench.py.
PY ?= python

.PHONY: build drivers test-cpu test-gpu bench smoke clean

build:            ## libs2mv.so (nvcc, sm_100a), the drivers, the CPU oracle (and oracle/_ref where /root/reference exists)
	$(PY) __graft_entry__.py

drivers: build
	$(MAKE) -C drivers

test-cpu:         ## oracle vs goldens, host logic (gloo world size 2), ABI, drivers' command lines
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu:         ## parity through the C ABI (needs a B200)
	$(PY) -m pytest tests -x -q -m gpu

smoke:
	$(PY) __graft_entry__.py --smoke

bench:
	$(PY) bench.py

clean:
	rm -f stereo-to-multiview-cuda_b200/libs2mv.so oracle/libs2mv_oracle.so
	$(MAKE) -C drivers clean
