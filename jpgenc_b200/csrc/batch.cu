// Batches of independent images on one GPU (BASELINE config 4: frames of a batch are sharded by image).
//
// An image is one unit of work with its own DC prediction chains, Huffman tables and file, so a batch needs no new
// kernels: `workers` contexts (each with its own streams and device buffers) are driven by as many host threads that
// pull frame indices from a shared counter.  While one frame waits for its symbol statistics or builds its tables on
// the host, the copies and kernels of the other frames keep PCIe and the SMs busy.
#include <atomic>
#include <thread>
#include <vector>

#include "common.cuh"

struct jpgenc_batch {
    int device = 0;
    std::vector<jpgenc_ctx*> ctx;
    std::string error;
};

namespace {

template <class Frame>
int run_batch(jpgenc_batch* b, uint32_t n, Frame&& encode_one) {
    std::atomic<uint32_t> next{0};
    std::atomic<int> first_error{JPGENC_OK};
    auto work = [&](jpgenc_ctx* c) {
        for (;;) {
            const uint32_t i = next.fetch_add(1, std::memory_order_relaxed);
            if (i >= n || first_error.load(std::memory_order_relaxed) != JPGENC_OK) return;
            const int rc = encode_one(c, i);
            if (rc != JPGENC_OK) {
                int expected = JPGENC_OK;
                if (first_error.compare_exchange_strong(expected, rc)) b->error = jpgenc_last_error(c);
                return;
            }
        }
    };
    std::vector<std::thread> threads;
    for (size_t k = 1; k < b->ctx.size() && k < n; ++k) threads.emplace_back(work, b->ctx[k]);
    work(b->ctx[0]);
    for (std::thread& t : threads) t.join();
    return first_error.load();
}

}  // namespace

extern "C" {

int jpgenc_batch_create(int device, int workers, jpgenc_batch** out) {
    if (!out || workers < 1 || workers > 64) return JPGENC_ERR_ARG;
    *out = nullptr;
    jpgenc_batch* b = new jpgenc_batch();
    b->device = device;
    for (int k = 0; k < workers; ++k) {
        jpgenc_ctx* c = nullptr;
        const int rc = jpgenc_create(device, &c);
        if (rc != JPGENC_OK) {
            for (jpgenc_ctx* x : b->ctx) jpgenc_destroy(x);
            delete b;
            return rc;
        }
        c->parallel_tables = false;          // the frames themselves run in parallel; no extra table threads per context
        b->ctx.push_back(c);
    }
    *out = b;
    return JPGENC_OK;
}

void jpgenc_batch_destroy(jpgenc_batch* b) {
    if (!b) return;
    for (jpgenc_ctx* c : b->ctx) jpgenc_destroy(c);
    delete b;
}

const char* jpgenc_batch_last_error(const jpgenc_batch* b) { return b ? b->error.c_str() : jpgenc_last_error(nullptr); }

int jpgenc_batch_set_qtables(jpgenc_batch* b, const uint8_t qy[64], const uint8_t qc[64]) {
    if (!b) return JPGENC_ERR_ARG;
    for (jpgenc_ctx* c : b->ctx) {
        const int rc = jpgenc_set_qtables(c, qy, qc);
        if (rc) { b->error = jpgenc_last_error(c); return rc; }
    }
    return JPGENC_OK;
}

int jpgenc_batch_encode(jpgenc_batch* b, uint32_t n, const uint8_t* const* frames, uint32_t w, uint32_t h, uint32_t maxval,
                        uint8_t* const* out, const uint64_t* caps, uint64_t* sizes) {
    if (!b || !frames || !out || !caps || !sizes) return JPGENC_ERR_ARG;
    return run_batch(b, n, [&](jpgenc_ctx* c, uint32_t i) {
        return jpgenc_encode_rgb(c, frames[i], w, h, maxval, out[i], caps[i], &sizes[i]);
    });
}

int jpgenc_batch_encode_device(jpgenc_batch* b, uint32_t n, const void* const* dev_frames, uint32_t w, uint32_t h,
                               uint32_t maxval, uint8_t* const* out, const uint64_t* caps, uint64_t* sizes) {
    if (!b || !dev_frames || !sizes) return JPGENC_ERR_ARG;
    return run_batch(b, n, [&](jpgenc_ctx* c, uint32_t i) {
        const int rc = jpgenc_bind_device_rgb(c, dev_frames[i], w, h, maxval);
        if (rc) return rc;
        return jpgenc_encode_bound(c, out ? out[i] : nullptr, caps ? caps[i] : 0, &sizes[i]);
    });
}

}  // extern "C"
