import os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "fullsize: BASELINE-config-size input (GPU only; skipped on the emulator / mock)")
    if config.getoption("--simt"):
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host", "simt"))
        import build as simt_build
        import genozip_b200.lib as lib
        lib.LIBPATH = simt_build.build()
        lib._lib = None
        lib._TESTS_MAY_LOAD_EMULATION = True
    if config.getoption("--dry-gpu"):
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import mock_gzb
        mock_gzb.install()


def pytest_addoption(parser):
    parser.addoption("--simt", action="store_true", default=False,
                     help="run the -m gpu tests against tests/host/_build/libgzb200_simt.so: the product's .cu sources compiled by g++ and "
                          "executed by the lock-step SIMT emulator (tests/host/simt). Checks kernel LOGIC on a machine without a GPU.")
    parser.addoption("--dry-gpu", action="store_true", default=False,
                     help="run the -m gpu tests' own logic on a machine WITHOUT a GPU: genozip_b200.Engine's marshalling code on top of "
                          "tests/mock_gzb.py (CPU checkers behind the C-ABI's entry points). Checks the tests and the binding, not the kernels.")


def pytest_collection_modifyitems(config, items):
    if config.getoption("--dry-gpu") or config.getoption("--simt"):
        skip = pytest.mark.skip(reason="--dry-gpu / --simt: needs torch device memory (covered by tests/test_fastq_path_cpu.py / test_bam_path_cpu.py)")
        for it in items:
            if "test_gpu_fastq_path" in it.nodeid or "test_zz_gpu_bam_path" in it.nodeid:
                it.add_marker(skip)
            if "fullsize" in it.keywords:
                it.add_marker(pytest.mark.skip(reason="--dry-gpu / --simt: BASELINE-size inputs are for the GPU"))
