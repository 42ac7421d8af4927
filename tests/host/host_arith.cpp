// host_arith.cpp — builds genozip_b200/csrc/arith_model.cuh (the product's arithmetic-coder chain logic) for the HOST with a
// one-lane "warp", so that the CPU test suite can check its model/range-coder logic against the oracle without a GPU.
// Test infrastructure only.  The device-only pieces (float reciprocal division, 8-entries-per-lane warp search) are
// covered by the -m gpu parity tests.
#include <stdint.h>
#include <stdlib.h>
#include <vector>
#include "../../genozip_b200/csrc/arith_model.cuh"

using namespace gzb;

static std::vector<uint32_t> make_models (uint32_t maxs, bool o1, bool rle)
{
    const uint32_t nctx = o1 ? 256 : 1, st = ar_stride (maxs);
    std::vector<uint32_t> m ((size_t)nctx * st + 258 * AR_RUN_STRIDE + 16);
    for (uint32_t c = 0; c < nctx; c++) ar_model_init (m.data () + (size_t)c * st, maxs);
    if (rle) for (uint32_t c = 0; c < 258; c++) ar_model_init (m.data () + (size_t)nctx * st + c * AR_RUN_STRIDE, 4);
    return m;
}

extern "C" uint32_t har_encode (const uint8_t *in, uint32_t n, int o1, int rle, uint8_t *out)
{
    uint32_t maxs = 0;
    for (uint32_t i = 0; i < n; i++) if (in[i] > maxs) maxs = in[i];
    maxs++;
    std::vector<uint32_t> m = make_models (maxs, o1, rle);
    return o1 ? ar_encode_leaf<true> (m.data (), maxs, rle, in, n, out, 0) : ar_encode_leaf<false> (m.data (), maxs, rle, in, n, out, 0);
}

extern "C" void har_decode (const uint8_t *body, uint32_t body_len, int o1, int rle, uint8_t *out, uint32_t n)
{
    const uint32_t maxs = body[0] ? body[0] : 256;
    std::vector<uint32_t> m = make_models (maxs, o1, rle);
    if (o1) ar_decode_leaf<true> (m.data (), maxs, rle, body, body_len, out, n, 0);
    else    ar_decode_leaf<false> (m.data (), maxs, rle, body, body_len, out, n, 0);
}
