"""The reference-shaped benchmark harness."""
import numpy
import pytest

from benchmarks.stencil_benchmarker import StencilBenchmarker, StencilTest, PrimedStencilTest
from stencil_code_b200.library.jacobi_stencil import Jacobi


class TinyBenchmarker(StencilBenchmarker):
    def matrix_iterator(self):
        for size in (6, 9):
            yield numpy.random.random([size, size]).astype(numpy.float32)


def test_harness_with_the_python_interpreter():
    lines = []
    rows = TinyBenchmarker([StencilTest("unprimed", stencil=Jacobi, backend='python'),
                            PrimedStencilTest("primed", stencil=Jacobi, backend='python')],
                           iterations=2).run(out=lines.append)
    assert len(rows) == 4 and len(lines) == 4
    assert rows[0][0] == (6, 6) and rows[0][1] == "unprimed" and rows[0][2] > 0
    assert lines[0].startswith("(6, 6),unprimed,")
    with pytest.raises(Exception):
        StencilTest("x", Jacobi, 'python').average_time("missing")


@pytest.mark.gpu
def test_reference_benchmark_drivers_on_cuda():
    from benchmarks import laplacian, convolve
    rows = laplacian.main(iterations=1)
    assert len(rows) == 3 * 2 * 2 and all(row[2] > 0 for row in rows)
    results = convolve.main(sizes=(256, 1024), iterations=2, out=lambda line: None)
    assert [r[0] for r in results] == [256, 1024]


def test_demo_helpers():
    from demo.bilateral_filter import synthetic_image, rescale_to_intensity
    image = synthetic_image(64, 48)
    assert image.shape == (48, 64) and image.dtype == numpy.uint8
    scaled = rescale_to_intensity(image.astype(numpy.float32) * 0.01, float(image.mean()))
    assert scaled.dtype == numpy.uint8 and abs(float(scaled.mean()) - float(image.mean())) < 1.5


@pytest.mark.gpu
def test_demo_on_cuda_matches_oracle():
    from demo.bilateral_filter import synthetic_image, run
    from cases import oracle_output
    from stencil_code_b200.library.better_bilateral_filter import BetterBilateralFilter
    image = synthetic_image(128, 96)
    result, _ = run(image)
    expected = oracle_output(BetterBilateralFilter(3, 70, backend='python'), image.astype(numpy.float32))
    from demo.bilateral_filter import rescale_to_intensity
    numpy.testing.assert_array_equal(result, rescale_to_intensity(expected, float(image.mean())))


@pytest.mark.gpu
def test_demo_on_the_reference_image_crop():
    """``demo/mallard.raw`` (64x64 crop committed as a golden vector together with the unmodified
    reference's output): backend='cuda' equals the oracle bit for bit, the reference's python
    backend on the interior to 6 decimals, and the demo's normalised 8-bit image
    (``demo/bilateral_filter.py:32-34``) to one grey level."""
    import os
    from cases import oracle_output
    from demo.bilateral_filter import rescale_to_intensity
    from stencil_code_b200.library.better_bilateral_filter import BetterBilateralFilter
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mallard_crop__clamp.npz")
    data = numpy.load(golden)
    image = data["image_in"]
    filtered = numpy.asarray(BetterBilateralFilter(backend='cuda')(image))
    expected = oracle_output(BetterBilateralFilter(backend='python'), image)
    assert (filtered.view(numpy.uint32) == expected.view(numpy.uint32)).all()
    interior = BetterBilateralFilter(backend='python').interior_points_slice()
    scale = float(numpy.abs(data["image_out"]).max())
    numpy.testing.assert_array_almost_equal(filtered[interior] / scale, data["image_out"][interior] / scale, decimal=6)
    ours = rescale_to_intensity(filtered[interior], float(image[interior].mean())).astype(numpy.int64)
    theirs = rescale_to_intensity(data["image_out"][interior], float(image[interior].mean())).astype(numpy.int64)
    assert numpy.abs(ours - theirs).max() <= 1
