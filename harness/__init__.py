"""harness -- synthetic inputs (SURVEY §8d generators) and the front-end call-pattern replay used by tests/ and bench.py.
Not product code: the product is practical-multi-view_b200/ (libpmv_cuda.so + the host adapters)."""
