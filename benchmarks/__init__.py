"""Benchmark drivers with the shape of reference ``benchmarks/`` running on ``backend='cuda'``."""
