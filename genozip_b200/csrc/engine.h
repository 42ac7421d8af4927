// engine.h — the per-(host thread, GPU) engine behind the opaque gzb_engine of include/gzb200.h
#pragma once
#include <string>
#include <vector>
#include <cuda_runtime.h>

struct gzb_engine {
    int          device = 0;
    int          sm_count = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr, stream3 = nullptr, stream4 = nullptr, stream_copy = nullptr, stream_copy2 = nullptr;   // stream2 / stream3 / stream4: the arithmetic chain kernels (general / order-0 / split encoder) run beside the rANS one; stream_copy / stream_copy2: gzb_stage_upload / gzb_stage_fetch transfers
    cudaEvent_t  ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, ev4 = nullptr, ev5 = nullptr, ev6 = nullptr, ev7 = nullptr;   // ev0..ev1: rANS chain kernel, ev3..ev2: arithmetic chain kernel, ev3..ev4: order-0 arithmetic, ev3..ev5: split arithmetic encoder
    uint8_t     *ws = nullptr;   size_t ws_cap = 0; // device workspace (grow-only)
    uint8_t     *pin = nullptr;  size_t pin_cap = 0;// pinned host staging (grow-only)
    std::string  err;
    uint64_t     launches = 0;
    float        last_chain_ms = 0, last_rans_ms = 0, last_arith_ms = 0, last_o0_ms = 0, last_split_ms = 0, last_arith_all_ms = 0, last_split_part_ms[3] = { 0, 0, 0 };
    float        last_domain_ms = 0;                // dominant kernel of the last PBWT / LONGR batch call (k_pbwt_rows, k_longr_channels, k_longr_decode)
    size_t       arena_hint = 0, arena_hint_dec = 0;
    // DOMQ session: device state kept between gzb_domq_prepare and gzb_domq_split of the same batch
    uint8_t     *dq_buf = nullptr; size_t dq_cap = 0;
    void        *dq_session = nullptr;
    void       (*dq_free)(void *) = nullptr;
    uint32_t     dq_n_vbs = 0; bool dq_devptr = false;
    // staging (stage.cu): the two feeder threads behind gzb_stage_upload / gzb_stage_fetch, made on first use
    void        *stager = nullptr;
    void       (*stager_free)(void *) = nullptr;
};

int engine_reserve (gzb_engine *e, size_t ws_bytes, size_t pin_bytes);
