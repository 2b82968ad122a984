.PHONY: build test test-gpu bench clean
build:
	python -c "import __graft_entry__ as g; g.build()"
test:
	python -m pytest tests/ -x -q -m "not gpu"
test-gpu:
	python -m pytest tests/ -x -q -m gpu
bench:
	python bench.py --gpus 1
clean:
	rm -rf sparktorch_b200/csrc/_build sparktorch_b200/_C*.so
