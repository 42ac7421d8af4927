"""The product's CUDA sources on the SIMT emulator (tests/host/simt), loaded NEXT TO the real library — test infrastructure.

`simt_engine_class()` returns a genozip_b200.lib.Engine subclass whose C-ABI calls go to tests/host/_build/libgzb200_simt.so:
the same .cu files compiled by g++ and executed lane by lane on the host, "device" memory being host memory."""
import ctypes as C
import os, sys

HERE = os.path.dirname(os.path.abspath(__file__))
_L = None


def simt_lib():
    global _L
    if _L is None:
        sys.path.insert(0, os.path.join(HERE, "host", "simt"))
        import build as simt_build
        import genozip_b200.lib as lib
        saved = (lib.LIBPATH, lib._lib)
        try:                                                               # lib.load() sets the argtypes; borrow it for the other path
            lib.LIBPATH, lib._lib, lib._TESTS_MAY_LOAD_EMULATION = simt_build.build(), None, True
            _L = lib.load()
        finally:
            lib.LIBPATH, lib._lib = saved
            lib._TESTS_MAY_LOAD_EMULATION = False
    return _L


def simt_engine_class():
    import genozip_b200.lib as lib
    L = simt_lib()

    class SimtEngine(lib.Engine):
        torch_device = "cpu"                                                # FastqCodecPath: buffers are ordinary host tensors

        def __init__(self, device=0):
            h = C.c_void_p()
            rc = L.gzb_engine_create(device, C.byref(h))
            assert rc == 0, L.gzb_last_error(None).decode()
            self.h, self.L, self.device = h, L, device

    return SimtEngine
