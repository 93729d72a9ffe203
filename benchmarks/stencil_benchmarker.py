"""Timing harness: the reference's ``StencilBenchmarker`` / ``StencilTest`` / ``PrimedStencilTest``
(``benchmarks/stencil_benchmarker.py:7-114``) for the CUDA backend.

A *test* times one way of running a stencil over a matrix; the *benchmarker* feeds every test the
same sequence of random matrices and prints ``shape,name,average seconds,checksum`` per shape, as
the reference does.  "Unprimed" tests build a fresh ``Stencil`` per trial, so they pay the
specialisation (here: lowering + NVRTC, mostly served from the shim's module cache after the first
trial); "primed" tests reuse one instance -- the difference is the reference's "with / without
compile" column (``README.md:29-34``).
"""
from __future__ import print_function

import time

import numpy


class Timer(object):
    """Wall-clock context manager (stands in for ``ctree.util.Timer``)."""
    def __enter__(self):
        self.start = time.perf_counter()
        return self

    def __exit__(self, *exc):
        self.interval = time.perf_counter() - self.start


class StencilTest(object):
    """One named way of running ``stencil`` with ``backend``; unprimed: new instance per trial."""

    def __init__(self, name, stencil, backend, setup_args=()):
        self.name = name
        self.stencil_class = stencil
        self.setup_args = setup_args
        self.backend = backend
        self.trial_times = dict()
        self.stencil_instance = None
        self.sum = 0.0

    def setup(self):
        pass

    def run(self, test_matrix):
        return self.stencil_class(*self.setup_args, backend=self.backend)(test_matrix)

    def run_trial(self, test_matrix, test_id=0):
        self.trial_times.setdefault(test_id, [])
        self.setup()
        with Timer() as timer:
            output = self.run(test_matrix)
        self.trial_times[test_id].append(timer.interval)
        self.sum += float(numpy.sum(output))

    def average_time(self, test_id):
        if test_id not in self.trial_times:
            raise Exception("Test {} does not contain data for test_id {}".format(self.name, test_id))
        times = self.trial_times[test_id]
        return sum(times) / float(len(times)) if times else 0


class PrimedStencilTest(StencilTest):
    """Reuses one stencil instance: specialisation is paid once."""

    def setup(self):
        if self.stencil_instance is None:
            self.stencil_instance = self.stencil_class(*self.setup_args, backend=self.backend)

    def run(self, test_matrix):
        return self.stencil_instance(test_matrix)


class StencilBenchmarker(object):
    """Runs every test on every matrix of ``matrix_iterator`` and prints the reference's table."""

    def __init__(self, tests_to_run, iterations=1, min_power=10, max_power=12):
        self.tests_to_run = tests_to_run
        self.iterations = iterations
        self.min_power = min_power
        self.max_power = max_power

    def matrix_iterator(self):
        sizes = [2 ** power for power in range(self.min_power, self.max_power)]
        for width in sizes:
            for height in sizes:
                yield numpy.random.random([height, width]).astype(numpy.float32)

    def run(self, out=print):
        numpy.random.seed(1)
        shapes = []
        for matrix in self.matrix_iterator():
            shapes.append(matrix.shape)
            for test in self.tests_to_run:
                for _ in range(self.iterations):
                    test.run_trial(matrix, matrix.shape)
        rows = []
        for shape in shapes:
            for test in self.tests_to_run:
                rows.append((shape, test.name, test.average_time(shape), test.sum))
                out("{},{},{},{:e}".format(*rows[-1]))
        return rows
