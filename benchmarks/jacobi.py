"""2-d Jacobi benchmark (reference ``benchmarks/jacobi.py``): 1024..2048 squared float32."""
from benchmarks.stencil_benchmarker import StencilBenchmarker, StencilTest, PrimedStencilTest
from stencil_code_b200.library.jacobi_stencil import Jacobi


def main(backend='cuda', iterations=3):
    return StencilBenchmarker([
        StencilTest("unprimed-" + backend, stencil=Jacobi, backend=backend),
        PrimedStencilTest("primed-" + backend, stencil=Jacobi, backend=backend),
    ], iterations=iterations).run()


if __name__ == '__main__':
    main()
