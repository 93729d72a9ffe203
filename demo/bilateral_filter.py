"""Bilateral-filter an 8-bit grayscale raw image on the GPU (reference ``demo/bilateral_filter.py``).

    python demo/bilateral_filter.py IN.raw WIDTH HEIGHT OUT.raw      # e.g. the reference's mallard.raw 800 800
    python demo/bilateral_filter.py --synthetic 800 800 OUT.raw      # no input file needed

Like the reference demo: sigma_d = 3 (radius 9), sigma_i = 70, and -- because the kernel does not
normalise by the weight sum -- the output is rescaled so its mean intensity matches the input's
(reference ``demo/bilateral_filter.py:17,32-34``) and clipped to 0..255.
"""
from __future__ import print_function

import os
import sys
import time

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stencil_code_b200.library.better_bilateral_filter import BetterBilateralFilter  # noqa: E402


def synthetic_image(width, height, seed=3):
    """Piecewise-smooth test card with edges and noise (stands in for a photograph)."""
    rng = numpy.random.RandomState(seed)
    y, x = numpy.mgrid[0:height, 0:width].astype(numpy.float32)
    image = 96 + 64 * numpy.sin(x / 37.0) * numpy.cos(y / 53.0)
    image[(x // 100 + y // 100) % 2 == 0] += 48                         # checkerboard edges
    image += rng.normal(0, 12, size=image.shape)                        # sensor noise
    return numpy.clip(image, 0, 255).astype(numpy.uint8)


def rescale_to_intensity(filtered, intensity):
    """The reference's post-step: match the mean intensity of the input, clip to 8 bits."""
    out_intensity = float(filtered.mean())
    return numpy.clip((filtered * (intensity / out_intensity)).astype(numpy.int64), 0, 255).astype(numpy.uint8)


def run(pixels, backend='cuda'):
    in_grid = pixels.astype(numpy.float32)
    stencil = BetterBilateralFilter(3, 70, backend=backend)
    begin = time.perf_counter()
    out_grid = numpy.asarray(stencil(in_grid))
    seconds = time.perf_counter() - begin
    return rescale_to_intensity(out_grid, float(pixels.mean())), seconds


def main(argv):
    if argv[1] == "--synthetic":
        width, height, out_path = int(argv[2]), int(argv[3]), argv[4]
        pixels = synthetic_image(width, height)
    else:
        width, height, out_path = int(argv[2]), int(argv[3]), argv[4]
        pixels = numpy.fromfile(argv[1], dtype=numpy.uint8, count=width * height).reshape(height, width)
    result, seconds = run(pixels)
    result.tofile(out_path)
    print("intensity {:.2f} -> {:.2f}; {}x{} filtered in {:.1f} ms (first call includes specialisation)".format(
        float(pixels.mean()), float(result.mean()), width, height, seconds * 1e3))


if __name__ == "__main__":
    if len(sys.argv) < 5:
        print(__doc__)
        sys.exit(1)
    main(sys.argv)
