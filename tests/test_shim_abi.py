"""The C-ABI library loads on a GPU-less host and exports every symbol the header declares."""
import ctypes
import os
import re

from stencil_code_b200.backend.cuda import runtime

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "stencil_cuda.h")) as handle:
        text = handle.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 40
    raw = ctypes.CDLL(runtime.LIBRARY_PATH)
    for name in names:
        assert hasattr(raw, name), name
    assert set(names) == set(runtime.SYMBOLS), set(names) ^ set(runtime.SYMBOLS)


def test_abi_version_and_error_string():
    library = runtime.load_library()
    assert library.sc_abi_version() == 1
    module = ctypes.c_void_p()
    code = library.sc_compile(b"this is not CUDA", b"bad.cu", None, 0, ctypes.byref(module))
    assert code == -3                                   # SC_ECOMPILE
    assert b"bad.cu" in library.sc_last_error()


def test_nvrtc_compiles_sm100a_cubin_without_a_gpu():
    module = runtime.compile_source('extern "C" __global__ void k(float* p) { p[threadIdx.x] = 1.0f; }',
                                    "tiny.cu", ("--fmad=false",))
    cubin = module.cubin()
    assert cubin[:4] == b"\x7fELF" and len(cubin) > 1000


def test_on_disk_cubin_cache(tmp_path, monkeypatch):
    monkeypatch.setenv("STENCIL_CUDA_CACHE", str(tmp_path))
    source = 'extern "C" __global__ void cached_kernel(float* p) { p[threadIdx.x] = 3.0f; }'
    first = runtime.compile_source(source, "cached.cu", ("--fmad=false",))
    assert not first.from_disk_cache and len(list(tmp_path.glob("cached_*.cubin"))) == 1
    second = runtime.compile_source(source, "cached.cu", ("--fmad=false",))
    assert second.from_disk_cache and second.cubin() == first.cubin()
    other = runtime.compile_source(source, "cached.cu", ("--fmad=true",))
    assert not other.from_disk_cache
    monkeypatch.setenv("STENCIL_CUDA_CACHE", "off")
    assert not runtime.compile_source(source, "cached.cu", ("--fmad=false",)).from_disk_cache
