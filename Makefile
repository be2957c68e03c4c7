# Convenience targets; the build itself is reinforcementlearning_blokus_b200/build.py (nvcc, sm_100a only).
PY ?= python

.PHONY: build test test-gpu bench bench-reference oracle golden clean

build:            ## compile libblokus_b200.so, the pipe-throughput probe and the CPU oracle
	$(PY) -c "import __graft_entry__ as g; g.build()"

test:             ## CPU suite (oracle vs reference fixtures, host logic, ABI exports, gloo)
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu:         ## parity suite through the C ABI (needs a B200)
	$(PY) -m pytest tests -x -q -m gpu

bench:            ## one JSON line (playouts/s, roofline, cpu_baseline, extras)
	$(PY) bench.py

bench-reference:  ## the CPU arm: C port of the reference's engine on all host threads
	$(PY) bench.py --impl reference --steps 3 --warmup 1

oracle:
	$(MAKE) -C oracle

golden:           ## regenerate tests/golden from the unmodified reference (build container only)
	PYTHONDONTWRITEBYTECODE=1 $(PY) tests/golden/generate_golden.py

clean:
	rm -f reinforcementlearning_blokus_b200/*.so reinforcementlearning_blokus_b200/*.srchash oracle/liborc.so
