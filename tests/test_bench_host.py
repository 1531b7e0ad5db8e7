"""Host-side logic of bench.py that needs no GPU: the default decompositions, the workload / config objects both arms must
agree on (the driver compares them), and the per-rank NUMA binding degrading to a no-op where NVML is absent."""
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(m)
    finally:
        sys.argv = argv
    return m


def _args(bench, *argv):
    old, sys.argv = sys.argv, ["bench.py", *argv]
    try:
        return bench.parse()
    finally:
        sys.argv = old


def test_default_blocks_never_split_x_and_multiply_to_n(bench):
    """x is never split by default: the 128-cell x-tiles stay full and the halo pushes are contiguous rows"""
    for n, b in bench.BLOCKS_FOR.items():
        assert b[0] * b[1] * b[2] == n and b[0] == 1


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_strong_scaling_workload_is_the_headline_grid(bench, n):
    a = _args(bench, "--gpus", str(n))
    cells, blocks, extent = bench.workload(a, n)
    assert cells == (512, 512, 512) and blocks == bench.BLOCKS_FOR[n]
    assert all(c % b == 0 for c, b in zip(cells, blocks))
    assert extent[1] - extent[0] == pytest.approx(12.0)


def test_weak_scaling_multiplies_the_grid_by_the_blocks(bench):
    a = _args(bench, "--gpus", "8", "--scaling", "weak", "--grid", "256")
    cells, blocks, _ = bench.workload(a, 8)
    assert cells == tuple(256 * b for b in blocks)


def test_both_arms_describe_the_same_config(bench):
    """the driver marks a pair of lines `same_config` only if the config objects are identical"""
    a = _args(bench)
    cells, _, _ = bench.workload(a, 1)
    c1, c2 = bench.config_of(a, cells), bench.config_of(_args(bench, "--impl", "reference"), cells)
    assert c1 == c2 and "workload" in c1 and "model" not in c1


def test_numa_binding_is_a_no_op_without_a_gpu(bench):
    before = os.sched_getaffinity(0)
    assert bench.bind_near_gpu(0) is None or isinstance(bench.bind_near_gpu(0), int)
    if bench.bind_near_gpu(0) is None:
        assert os.sched_getaffinity(0) == before
