// Helpers shared by the translation units of the C-ABI layer (capi.cu: context, stage calls, one image; pipeline.cu:
// batches of frames).  Not part of the boundary: include/jpgenc_b200.h is.
#pragma once

#include "common.cuh"

namespace jpgenc { namespace detail {

constexpr size_t kStatsBytes = 4096 + 8192;   // per frame: symbol histogram u32[4][256] + first-occurrence keys u64[4][256] (stats.cu)

int fail(jpgenc_ctx* c, int code, const char* what);

// device buffer of at least need_bytes, grown on demand.  `headroom`: for buffers whose size follows the data (scan sizes
// differ from pass to pass): a reallocation frees device memory, which synchronises the whole device, so it should happen
// a few times, not whenever a pass is slightly larger than the last one
int ensure_bytes(jpgenc_ctx* c, void** ptr, size_t* cap, size_t need_bytes, bool headroom);
template <class T>
int ensure(jpgenc_ctx* c, T** ptr, size_t* cap, size_t need_bytes, bool headroom = false) {
    void* p = *ptr;
    const int rc = ensure_bytes(c, &p, cap, need_bytes, headroom);
    *ptr = static_cast<T*>(p);
    return rc;
}

int set_geometry(jpgenc_ctx* c, uint32_t w, uint32_t h, uint32_t maxval);
int ensure_pinned(jpgenc_ctx* c, size_t bytes);                 // c->h_pinned, grown on demand
int ensure_coef(jpgenc_ctx* c);                                  // coefficients + refinement list of the bound frames
int ensure_stats_buffers(jpgenc_ctx* c);                         // K2's outputs for the bound frames (+ pinned staging)
int ensure_entropy_buffers(jpgenc_ctx* c, uint64_t raw_total, uint64_t out_total, uint64_t k4_tiles);
void ensure_host_pool(jpgenc_ctx* c);
void leave_batch_state(jpgenc_ctx* c);
int flush_entropy_time(jpgenc_ctx* c);
int poll_mailbox(jpgenc_ctx* c, int word, unsigned long long* value);

// pinned staging of a context: [statistics F * kStatsBytes + 16][device tables F * 8 KB][PassMeta block (common.cuh)]
size_t stage_tables_off(uint32_t F);
size_t stage_meta_off(uint32_t F);
size_t stage_bytes(uint32_t F);

uint32_t env_u32(const char* name, uint32_t dflt);
bool trace_on();                                                 // JPGENC_TRACE=1: host wall-clock of the phases on stderr
double now_us();

}}  // namespace jpgenc::detail
