// Batches of equally sized frames: every pass (a slice of the batch) goes through ALL kernels as one asynchronous chain
//
//     K1 (+ exact refinement) -> K2 -> build_tables_kernel -> finalize_tables_kernel -> K3a -> K3b -> K4
//
// with no host step in between: the Huffman tables are built on the device (tables_device.cu), the sizes and offsets that
// follow from them are computed on the device (finalize.cu), K4 lays the complete files of the pass -- header, stuffed scan,
// EOI per frame -- back to back into one buffer, and ONE small device-to-host copy at the end of the chain tells the host
// what happened.  The host synchronises once per pass.
//
// What this replaces: a caller looping over Image::writeJPEG (src/Image.cpp:831-976) for a sequence of frames; every frame
// keeps its own DC prediction chains, its own four image-optimal tables and its own file, results are
// byte-identical to encoding the frames one by one.
//
// Overlap.  A table build is a long, thin kernel (one working thread per table, ~0.2 ms for a 30-symbol alphabet whatever
// the number of frames); K1/K2/K3 are short and wide.  Passes therefore rotate over several SLOTS -- the context itself
// and lazily created further contexts on the same device, each with its own stream and buffers -- so that the table build
// of one pass runs beside the wide kernels of the others.  All slots are driven by the calling thread alone: per pass it
// enqueues about a dozen asynchronous operations and later waits for one event.  No worker threads, nothing spinning beside
// the caller -- eight such processes on a 32-core host do not compete for cores (the first design -- host-built tables,
// three lanes with a thread each, a pool of table builders -- was host-bound there: SCALE_r01 / VERDICT round 1).
#include <algorithm>
#include <cstring>
#include <vector>

#include "internal.hpp"

using namespace jpgenc;
using namespace jpgenc::detail;

namespace {

constexpr uint32_t kDefaultSlots = 4, kMaxSlots = 6;
constexpr uint32_t kMaxHeaderBytes = 2 + 18 + 2 * 69 + 19 + 4 * (21 + 256) + 14;   // SOI APP0 DQTx2 SOF0 DHTx4 SOS

// JPGENC_TRACE=2: a device-side timeline of the call -- timing events between the stages of every pass, printed relative to
// the first one when the call ends (development aid; the extra event records cost a few microseconds per pass)
struct Timeline {
    struct Mark { cudaEvent_t ev; uint32_t pass; const char* what; };
    std::vector<Mark> marks;
    bool on = false;
    void mark(jpgenc_ctx* l, uint32_t pass, const char* what) {
        if (!on) return;
        cudaEvent_t ev = nullptr;
        if (cudaEventCreate(&ev) != cudaSuccess) return;
        cudaEventRecord(ev, l->stream);
        marks.push_back({ev, pass, what});
    }
    void print() {
        if (marks.empty()) return;
        cudaEventSynchronize(marks.back().ev);
        for (const Mark& m : marks) {
            float ms = 0;
            cudaEventSynchronize(m.ev);
            cudaEventElapsedTime(&ms, marks.front().ev, m.ev);
            std::fprintf(stderr, "[jpgenc timeline] pass %u %-10s %8.1f us\n", m.pass, m.what, ms * 1e3);
        }
        for (const Mark& m : marks) cudaEventDestroy(m.ev);
        marks.clear();
    }
};
Timeline g_timeline;

struct Job {
    uint32_t n = 0;
    uint32_t w = 0, h = 0, maxval = 255;
    // output: per-frame buffers (legacy), one packed buffer, or none (sizes only)
    uint8_t* const* out_ptrs = nullptr;
    const uint64_t* caps = nullptr;
    bool packed_mode = false;              // files back to back: offsets[] are reported, `packed` (may be null) receives them
    uint8_t* packed = nullptr;
    uint64_t packed_cap = 0, packed_at = 0;
    uint64_t* offsets = nullptr;
    uint64_t* sizes = nullptr;
};

struct Pass {
    uint32_t f0 = 0, F = 0;
    uint32_t k4_grid = 0;
    int attempts = 0;
};

// the largest pass one slot can hold: item slabs (worst-case reservation, 96 KB per tile) within ~2 GB, block ids in 31 bits
uint32_t frames_per_pass_cap(const jpgenc_ctx* c) {
    const size_t nblocks = static_cast<size_t>(c->mcu_w) * c->mcu_h * kBlocksPerMcu, tiles = (nblocks + 383) / 384;
    size_t per_pass = std::max<size_t>(1, (2ull << 30) / (tiles * 384 * 64 * 4));
    per_pass = std::min<size_t>(per_pass, 0x7FFFFFFFull / nblocks);
    return static_cast<uint32_t>(std::min<size_t>(per_pass, 1024));
}

// Frames per pass.  A pass costs ~0.1 ms of launch gaps and one host wait whatever its size, its wide kernels run the better
// the larger they are, and its table build (~0.2 ms, independent of the number of frames) only hides behind OTHER passes' wide
// kernels: passes should be large, but a batch should still give every slot about two of them.  Measured on 1920x1080
// frames, 4 slots, device tables at 0.21 ms (frames per pass -> frames/s):
// 1024 frames: 64 -> 147 k, 128 -> 159 k, 256 -> 157 k;  128 frames: 32 -> 117 k, 64 -> 120 k, 128 -> 120 k.
uint32_t pass_frames(const jpgenc_ctx* c, uint32_t n, uint32_t slots) {
    const size_t px = static_cast<size_t>(c->mcu_w) * c->mcu_h * 256;
    const uint32_t lo = static_cast<uint32_t>(std::max<size_t>(1, (128u << 20) / px));    // ~128 Mpx: 64 frames of 1920x1080
    const uint32_t hi = static_cast<uint32_t>(std::max<size_t>(1, (256u << 20) / px));    // ~256 Mpx: 128 such frames
    uint32_t per = std::min(hi, std::max(lo, (n + 2 * slots - 1) / (2 * slots)));
    per = env_u32("JPGENC_FRAMES_PER_PASS", per);
    return std::max(1u, std::min({per, frames_per_pass_cap(c), n}));
}

int prepare_slot(jpgenc_ctx* root, jpgenc_ctx* l, const Job& job, const uint8_t* prefix, uint32_t prefix_len) {
    if (l != root) {
        std::memcpy(l->qy, root->qy, 64); std::memcpy(l->qc, root->qc, 64);
        std::memcpy(l->dct_a, root->dct_a, sizeof root->dct_a); std::memcpy(l->dct_s, root->dct_s, sizeof root->dct_s);
    }
    int rc = set_geometry(l, job.w, job.h, job.maxval);
    if (rc) return rc;
    if (!l->d_hdr_prefix) JPGENC_CUDA(l, cudaMalloc(reinterpret_cast<void**>(&l->d_hdr_prefix), 256));
    if (l->hdr_prefix_len != prefix_len || std::memcmp(l->hdr_prefix_host, prefix, prefix_len) != 0) {   // new size or quantisers
        JPGENC_CUDA(l, cudaMemcpyAsync(l->d_hdr_prefix, prefix, prefix_len, cudaMemcpyHostToDevice, l->stream));
        JPGENC_CUDA(l, cudaStreamSynchronize(l->stream));      // `prefix` lives on the caller's stack
        std::memcpy(l->hdr_prefix_host, prefix, prefix_len);
        l->hdr_prefix_len = prefix_len;
    }
    if (!l->ev_done) JPGENC_CUDA(l, cudaEventCreateWithFlags(&l->ev_done, cudaEventDisableTiming));
    if (!l->ev_wide) JPGENC_CUDA(l, cudaEventCreateWithFlags(&l->ev_wide, cudaEventDisableTiming));
    if (!l->ev_fwd) JPGENC_CUDA(l, cudaEventCreateWithFlags(&l->ev_fwd, cudaEventDisableTiming));
    if (!l->ev_k4) JPGENC_CUDA(l, cudaEventCreateWithFlags(&l->ev_k4, cudaEventDisableTiming));
    if (!l->ev_copied) JPGENC_CUDA(l, cudaEventCreateWithFlags(&l->ev_copied, cudaEventDisableTiming));
    if (!l->out_stream) JPGENC_CUDA(l, cudaStreamCreateWithFlags(&l->out_stream, cudaStreamNonBlocking));
    l->copy_pending = false;
    return JPGENC_OK;
}

// K3/K4 of a pass: buffers for the limits, finalize (tables -> lookup form, sizes, offsets), K3a, K3b, K4, read-back
int enqueue_entropy(jpgenc_ctx* l, Pass& ps, cudaEvent_t prev_k4 = nullptr) {
    const uint32_t F = ps.F;
    int rc;
    const uint64_t k4_grid = l->raw_limit / kK4TileBytes + F;
    if ((rc = ensure_entropy_buffers(l, l->raw_limit, l->out_limit, k4_grid))) return rc;
    ps.k4_grid = static_cast<uint32_t>(k4_grid);
    if (l->copy_pending) {                                     // the previous pass's files are still leaving d_scan
        JPGENC_CUDA(l, cudaStreamWaitEvent(l->stream, l->ev_copied, 0));
        l->copy_pending = false;
    }
    if ((rc = launch_finalize_tables(l))) return rc;
    // K3/K4 of the passes run one pass after the other (each of them fills the GPU anyway): the passes then FINISH one
    // after the other and the files of pass p travel while pass p + 1 is packed, instead of all slots' files queueing up
    // behind the last kernel (trace of 4 slots in step: 4 x 11 MB = 0.8 ms of copies after the last K4)
    if (prev_k4) JPGENC_CUDA(l, cudaStreamWaitEvent(l->stream, prev_k4, 0));
    g_timeline.mark(l, ps.f0, "finalize");
    if ((rc = launch_entropy(l, ps.k4_grid))) return rc;
    g_timeline.mark(l, ps.f0, "K3+K4");
    JPGENC_CUDA(l, cudaEventRecord(l->ev_k4, l->stream));
    uint8_t* h = static_cast<uint8_t*>(l->h_pinned) + stage_meta_off(F);
    JPGENC_CUDA(l, cudaMemcpyAsync(h, l->d_meta, pass_meta_bytes(F), cudaMemcpyDeviceToHost, l->stream));
    JPGENC_CUDA(l, cudaEventRecord(l->ev_done, l->stream));
    ++ps.attempts;
    return JPGENC_OK;
}

// the whole chain of one pass on slot `l`; `ready` (optional): event after which the pixels are valid
// `after` (optional): the previous pass's "wide kernels done" event (see run_passes)
int enqueue_pass(jpgenc_ctx* root, jpgenc_ctx* l, Pass& ps, const void* const* dev_frames, cudaEvent_t ready, cudaEvent_t after, cudaEvent_t prev_k4) {
    const uint32_t F = ps.F;
    int rc;
    l->nframes = F;
    l->have_coef = l->have_scan = l->have_items = false; l->k2_tiles_done = 0;
    l->file_mode = true;
    bool aligned = true;
    for (uint32_t f = 0; f < F; ++f) aligned = aligned && (reinterpret_cast<uintptr_t>(dev_frames[f]) % 16 == 0);
    l->frames_aligned = aligned;
    l->d_rgb = static_cast<const uint8_t*>(dev_frames[0]);
    l->have_pixels = true;
    if ((rc = ensure_coef(l))) return rc;
    if ((rc = ensure_stats_buffers(l))) return rc;               // also sizes the pinned staging (stage_bytes(F))
    {
        size_t cap = l->frame_ptrs_cap;
        void* p = l->d_frame_ptrs;
        if ((rc = ensure(l, reinterpret_cast<uint8_t**>(&p), &cap, F * sizeof(void*)))) return rc;
        l->d_frame_ptrs = static_cast<const uint8_t**>(p);
        l->frame_ptrs_cap = cap;
    }
    // the pointer array travels through pinned memory (the statistics' staging area is unused in this path), so that the
    // copy is asynchronous and the caller's array need not outlive the call
    std::memcpy(l->h_pinned, dev_frames, F * sizeof(void*));
    JPGENC_CUDA(l, cudaMemcpyAsync(l->d_frame_ptrs, l->h_pinned, F * sizeof(void*), cudaMemcpyHostToDevice, l->stream));
    if (ready) JPGENC_CUDA(l, cudaStreamWaitEvent(l->stream, ready, 0));
    if (after) JPGENC_CUDA(l, cudaStreamWaitEvent(l->stream, after, 0));
    g_timeline.mark(l, ps.f0, "start");
    if ((rc = launch_forward_rows(l, 0, l->mcu_h, true, true))) return rc;                       // K1 + exact refinement
    g_timeline.mark(l, ps.f0, "K1+refine");
    l->have_coef = true;
    JPGENC_CUDA(l, cudaEventRecord(l->ev_fwd, l->stream));
    const size_t nblocks = static_cast<size_t>(l->mcu_w) * l->mcu_h * kBlocksPerMcu, tiles = (nblocks + 383) / 384;
    if ((rc = launch_symbol_stats(l, 0, static_cast<uint32_t>(tiles * F), true))) return rc;     // K2
    l->have_items = true;
    g_timeline.mark(l, ps.f0, "K2");
    JPGENC_CUDA(l, cudaEventRecord(l->ev_wide, l->stream));
    const uint32_t nt = 4 * F;
    if ((rc = ensure(l, reinterpret_cast<uint8_t**>(&l->d_tab_scratch), &l->tab_scratch_cap, nt * table_scratch_bytes()))) return rc;
    if ((rc = ensure(l, &l->d_built_tables, &l->built_tables_cap, nt * (sizeof(jpgenc_huff_table) + sizeof(uint32_t))))) return rc;
    // (JPGENC_DEBUG_SKIP_TABLES=1, development only: keep the tables of the previous call -- valid output only when the same
    // frames are encoded again; shows what the pipeline would do with a free table build)
    static const bool skip_tables = env_u32("JPGENC_DEBUG_SKIP_TABLES", 0) != 0;
    if (!(skip_tables && l->debug_tables_valid >= nt))
        if ((rc = launch_build_tables(l, l->d_stats, static_cast<uint32_t>(kStatsBytes), nt, l->d_tab_scratch, l->d_built_tables,
                                      reinterpret_cast<uint32_t*>(l->d_built_tables + nt)))) return rc;
    l->debug_tables_valid = nt;
    g_timeline.mark(l, ps.f0, "tables");
    // how much raw scan / output the pass may use: 1.5 x the largest frame the batch has produced so far
    // (root->batch_raw_per_frame), before the first pass has finished a guess of 12 bytes per block; a pass that needs more
    // is refused by finalize_tables_kernel and its entropy stage re-run with what it asked for (finish_pass)
    // (JPGENC_RAW_GUESS_PER_BLOCK: tests set it to 0 so that every first pass is refused and re-run)
    const uint64_t per_frame = root->batch_raw_per_frame ? root->batch_raw_per_frame : nblocks * env_u32("JPGENC_RAW_GUESS_PER_BLOCK", 12) + 4096;
    l->raw_limit = F * raw_slot_bytes(per_frame);
    l->out_limit = F * (kMaxHeaderBytes + 2ull) + 2 * l->raw_limit;
    ps.attempts = 0;
    return enqueue_entropy(l, ps, prev_k4);
}

// waits for the pass on slot `l`, re-runs its entropy stage with larger buffers if it was refused, reports sizes and sends
// the files on their way to the caller's memory (the copies are enqueued on the slot's stream; the caller of run_passes
// synchronises all slots at the end)
int finish_pass(jpgenc_ctx* root, jpgenc_ctx* l, Pass& ps, Job& job) {
    const uint32_t F = ps.F;
    for (;;) {
        JPGENC_CUDA(l, cudaEventSynchronize(l->ev_done));
        const PassMeta m = pass_meta_view(static_cast<uint8_t*>(l->h_pinned) + stage_meta_off(F), F);
        const PassHeader& hd = *m.hdr;
        if (hd.nframes != F) return fail(l, JPGENC_ERR_CUDA, "batched pass: the device left no result");
        if (hd.error & kPassMissingSymbol) return fail(l, JPGENC_ERR_ARG, "Huffman table lacks a symbol that occurs in the image");
        if (hd.error & kPassNoSymbols) return fail(l, JPGENC_ERR_ARG, "Huffman table build failed: a table has no symbol");
        if (hd.error & (kPassRawOverflow | kPassOutOverflow)) {
            if (ps.attempts >= 3) return fail(l, JPGENC_ERR_CUDA, "batched pass: scan buffers still too small after two retries");
            // finalize computed the full requirement before refusing: reserve that (plus a margin) and run K3/K4 again --
            // coefficients, symbol items and tables are still on the device
            l->raw_limit = hd.raw_total + hd.raw_total / 8 + 4096;
            l->out_limit = F * (kMaxHeaderBytes + 2ull) + 2 * l->raw_limit;
            const int rc = enqueue_entropy(l, ps);
            if (rc) return rc;
            continue;
        }
        uint64_t max_raw = 0;
        for (uint32_t f = 0; f < F; ++f) {
            if (m.total_bits[f] != m.frame_bits[f]) {
                l->error = "entropy coder wrote " + std::to_string(m.total_bits[f]) + " bits, statistics predicted " + std::to_string(m.frame_bits[f]);
                return JPGENC_ERR_ARG;
            }
            max_raw = std::max<uint64_t>(max_raw, m.raw_bytes[f]);
        }
        // later passes reserve 1.5 x the largest frame seen so far
        root->batch_raw_per_frame = std::max<uint64_t>(root->batch_raw_per_frame, max_raw + max_raw / 2 + 4096);
        const uint64_t total = hd.out_total + m.ff_incl[F - 1];
        // The files go home on the slot's OUTPUT stream (behind ev_done, i.e. behind K4), so that the slot's next pass can
        // start its K1/K2/table build meanwhile; only its K3/K4, which overwrite d_scan, wait for the copy (enqueue_entropy).
        uint64_t base = 0;
        const bool copies = job.packed || job.out_ptrs;
        if (copies) JPGENC_CUDA(l, cudaStreamWaitEvent(l->out_stream, l->ev_done, 0));
        if (job.packed_mode) {                                       // passes finish in order: the files end up in frame order
            base = job.packed_at;
            job.packed_at += total;
            if (job.packed) {
                if (job.packed_at > job.packed_cap) return fail(l, JPGENC_ERR_CAPACITY, "JPEG buffer too small");
                JPGENC_CUDA(l, cudaMemcpyAsync(job.packed + base, l->d_scan, total, cudaMemcpyDeviceToHost, l->out_stream));
            }
        }
        for (uint32_t f = 0; f < F; ++f) {
            const uint64_t before = f ? m.ff_incl[f - 1] : 0;
            const uint64_t off = m.file_base[f] + before;
            const uint64_t size = m.hdr_len[f] + m.raw_bytes[f] + (m.ff_incl[f] - before) + 2;
            job.sizes[ps.f0 + f] = size;
            if (job.offsets) job.offsets[ps.f0 + f] = base + off;
            if (job.out_ptrs) {
                if (job.caps[ps.f0 + f] < size) return fail(l, JPGENC_ERR_CAPACITY, "JPEG buffer too small");
                JPGENC_CUDA(l, cudaMemcpyAsync(job.out_ptrs[ps.f0 + f], l->d_scan + off, size, cudaMemcpyDeviceToHost, l->out_stream));
            }
        }
        if (copies) {
            JPGENC_CUDA(l, cudaEventRecord(l->ev_copied, l->out_stream));
            l->copy_pending = true;
        }
        return JPGENC_OK;
    }
}

// Runs the passes over the slots.  frames_of(p, slot context) -> device pointers of pass p's frames and the event after
// which they are valid; done_with(p) is called when pass p's pixels are no longer needed.
template <class FramesOf, class DoneWith>
int run_passes(jpgenc_ctx* c, Job& job, uint32_t per_pass, uint32_t nslots, FramesOf&& frames_of, DoneWith&& done_with) {
    const uint32_t npasses = (job.n + per_pass - 1) / per_pass;
    // SOI .. SOF0 are the same bytes for every frame of the batch: jpgenc_write_headers with empty tables gives
    // prefix + 4 x (DHT with no symbol: 21 bytes) + SOS (14 bytes)
    uint8_t hdr[512];
    jpgenc_huff_table empty[4];
    std::memset(empty, 0, sizeof empty);
    const size_t hdr_all = jpgenc_write_headers(job.w, job.h, c->qy, c->qc, empty, hdr);
    const uint32_t prefix_len = static_cast<uint32_t>(hdr_all - 4 * 21 - 14);
    int rc;
    while (c->lanes.size() + 1 < nslots) {
        jpgenc_ctx* l = nullptr;
        if ((rc = jpgenc_create(c->device, &l))) return fail(c, rc, jpgenc_last_error(nullptr));
        c->lanes.push_back(l);
    }
    std::vector<jpgenc_ctx*> slot(nslots);
    for (uint32_t k = 0; k < nslots; ++k) {
        slot[k] = k ? c->lanes[k - 1] : c;
        if ((rc = prepare_slot(c, slot[k], job, hdr, prefix_len))) return k ? fail(c, rc, jpgenc_last_error(slot[k])) : rc;
    }
    std::vector<Pass> pass(npasses);
    // (measured: 128 frames 1.26 -> 1.33 ms, 1024 frames 6.9 -> 7.2 ms: the cross-stream waits cost more than the earlier
    // table builds gain; kept as a switch)
    // default: the streams run side by side, except for a call of exactly two passes -- there nothing else will ever run beside
    // the table builds, and letting pass 1's K1 start when pass 0's K1 is through (mode 2) puts pass 0's table build beside pass 1's
    // wide kernels (timeline, 128 frames: 1.057 -> 1.014 ms; 256 and 1024 frames are 2-4 % slower staggered)
    const uint32_t stagger = env_u32("JPGENC_STAGGER", npasses == 2 ? 2u : 0u);
    rc = JPGENC_OK;
    jpgenc_ctx* failed = nullptr;
    const double t0 = trace_on() ? now_us() : 0;
    g_timeline.on = env_u32("JPGENC_TRACE", 0) >= 2;
    for (uint32_t p = 0; p < npasses + nslots && rc == JPGENC_OK; ++p) {
        if (p >= nslots) {                                          // the slot's previous pass: results, files, slot free again
            const uint32_t q = p - nslots;
            jpgenc_ctx* l = slot[q % nslots];
            rc = finish_pass(c, l, pass[q], job);
            if (rc) { failed = l; break; }
            done_with(q);
            if (trace_on()) std::fprintf(stderr, "[jpgenc batch] pass %u (%u frames, slot %u) finished at %.0f us\n", q, pass[q].F, q % nslots, now_us() - t0);
        }
        if (p < npasses) {
            jpgenc_ctx* l = slot[p % nslots];
            pass[p].f0 = p * per_pass;
            pass[p].F = std::min(per_pass, job.n - pass[p].f0);
            const void* const* ptrs = nullptr;
            cudaEvent_t ready = nullptr;
            rc = frames_of(p, l, &ptrs, &ready);
            // Staggering: the wide kernels of pass p (K1, refinement, K2) start when those of pass p - 1 are through.  Left
            // alone, the streams run the same stage of all passes side by side, every table build then starts at the same
            // (late) moment and nothing wide is left to run beside it.
            cudaEvent_t after = nullptr;
            if (stagger && p > 0 && nslots > 1) after = stagger == 2 ? slot[(p - 1) % nslots]->ev_fwd : slot[(p - 1) % nslots]->ev_wide;
            // (only when files travel back: sizes-only calls are 2-5 % faster with the slots' K3/K4 side by side.  Measured,
            // 1024 frames, in order / not: files returned 6.61 / 7.00 ms, sizes only 6.30 / 6.20 ms; JPGENC_ENTROPY_ORDER=0/1 forces)
            static const uint32_t order_env = env_u32("JPGENC_ENTROPY_ORDER", 2);
            const bool in_order = order_env == 2 ? (job.packed != nullptr || job.out_ptrs != nullptr) : order_env != 0;
            cudaEvent_t prev_k4 = (in_order && p > 0 && nslots > 1) ? slot[(p - 1) % nslots]->ev_k4 : nullptr;
            if (rc == JPGENC_OK) rc = enqueue_pass(c, l, pass[p], ptrs, ready, after, prev_k4);
            if (rc) { failed = l; break; }
        }
    }
    // the files' copies are still in flight; also after a failure: nothing may run once the caller's buffers are gone
    for (uint32_t k = 0; k < nslots; ++k) {
        cudaError_t e = cudaStreamSynchronize(slot[k]->stream);
        if (e == cudaSuccess && slot[k]->out_stream) e = cudaStreamSynchronize(slot[k]->out_stream);
        slot[k]->copy_pending = false;
        if (e != cudaSuccess && rc == JPGENC_OK) { rc = JPGENC_ERR_CUDA; c->error = std::string("batched pass: ") + cudaGetErrorString(e); }
    }
    g_timeline.print();
    if (failed && failed != c) c->error = failed->error;
    for (uint32_t k = 1; k < nslots; ++k) {
        c->launches += slot[k]->launches;
        slot[k]->launches = 0;
        leave_batch_state(slot[k]);
    }
    leave_batch_state(c);
    return rc;
}

uint32_t slots_for(uint32_t n, uint32_t per_pass) {
    const uint32_t npasses = (n + per_pass - 1) / per_pass;
    uint32_t s = env_u32("JPGENC_SLOTS", env_u32("JPGENC_LANES", kDefaultSlots));
    return std::max(1u, std::min({s, kMaxSlots, npasses}));
}

int check_job(jpgenc_ctx* c, const Job& job, const void* frames) {
    if (!c || !frames || !job.sizes) return JPGENC_ERR_ARG;
    if (job.out_ptrs && !job.caps) return JPGENC_ERR_ARG;
    if (job.packed_mode && !job.offsets) return JPGENC_ERR_ARG;
    return JPGENC_OK;
}

int encode_device_frames(jpgenc_ctx* c, Job& job, const void* const* dev_frames) {
    // an empty batch is legal and touches nothing: an image bound to the context keeps its geometry
    if (job.n == 0) return (job.w && job.h && job.maxval && job.maxval <= 255) ? JPGENC_OK : fail(c, JPGENC_ERR_ARG, "width/height/maxval out of range");
    JPGENC_CUDA(c, cudaSetDevice(c->device));
    int rc = set_geometry(c, job.w, job.h, job.maxval);
    if (rc) return rc;
    // what earlier calls learnt about the size of such frames' scans stays valid (a later frame that needs more is refused on the
    // device and re-run, finish_pass); a new frame size starts from the guess again
    if (c->batch_geom_w != job.w || c->batch_geom_h != job.h) { c->batch_raw_per_frame = 0; c->batch_geom_w = job.w; c->batch_geom_h = job.h; }
    const uint32_t nslots0 = std::min(kMaxSlots, env_u32("JPGENC_SLOTS", env_u32("JPGENC_LANES", kDefaultSlots)));
    const uint32_t per_pass = pass_frames(c, job.n, std::max(1u, nslots0));
    const uint32_t nslots = slots_for(job.n, per_pass);
    return run_passes(c, job, per_pass, nslots,
                      [&](uint32_t p, jpgenc_ctx*, const void* const** ptrs, cudaEvent_t* ready) {
                          *ptrs = dev_frames + static_cast<size_t>(p) * per_pass;
                          *ready = nullptr;
                          return JPGENC_OK;
                      },
                      [](uint32_t) {});
}

// Frames in host memory (pinned for full PCIe speed): they travel on the copy stream into a ring of pass-sized slices
// of device memory, always a few passes ahead of the kernels; when pass p is finished its slice is re-filled with the
// frames of pass p + ring.
int encode_host_frames(jpgenc_ctx* c, Job& job, const uint8_t* const* frames) {
    // an empty batch is legal and touches nothing: an image bound to the context keeps its geometry
    if (job.n == 0) return (job.w && job.h && job.maxval && job.maxval <= 255) ? JPGENC_OK : fail(c, JPGENC_ERR_ARG, "width/height/maxval out of range");
    JPGENC_CUDA(c, cudaSetDevice(c->device));
    int rc = set_geometry(c, job.w, job.h, job.maxval);
    if (rc) return rc;
    // what earlier calls learnt about the size of such frames' scans stays valid (a later frame that needs more is refused on the
    // device and re-run, finish_pass); a new frame size starts from the guess again
    if (c->batch_geom_w != job.w || c->batch_geom_h != job.h) { c->batch_raw_per_frame = 0; c->batch_geom_w = job.w; c->batch_geom_h = job.h; }
    const uint32_t nslots0 = std::min(kMaxSlots, env_u32("JPGENC_SLOTS", env_u32("JPGENC_LANES", kDefaultSlots)));
    const uint32_t per_pass = pass_frames(c, job.n, std::max(1u, nslots0));
    const uint32_t nslots = slots_for(job.n, per_pass);
    const uint32_t npasses = (job.n + per_pass - 1) / per_pass;
    constexpr uint32_t kRingMax = sizeof(c->ev_band) / sizeof(c->ev_band[0]);
    const uint32_t ring = std::min({nslots + 2, npasses, kRingMax});
    const size_t fbytes = static_cast<size_t>(job.w) * job.h * 3, fstride = (fbytes + 255) & ~static_cast<size_t>(255);
    if ((rc = ensure(c, &c->d_rgb_owned, &c->rgb_cap, static_cast<size_t>(ring) * per_pass * fstride + 16))) return rc;
    auto upload = [&](uint32_t p) -> int {
        const uint32_t f0 = p * per_pass, F = std::min(per_pass, job.n - f0), s = p % ring;
        // frames that follow each other in host memory (a video buffer) travel as one strided copy
        for (uint32_t f = 0; f < F;) {
            uint32_t run = 1;
            while (f + run < F && frames[f0 + f + run] == frames[f0 + f + run - 1] + fbytes) ++run;
            uint8_t* dst = c->d_rgb_owned + (static_cast<size_t>(s) * per_pass + f) * fstride;
            if (run == 1) JPGENC_CUDA(c, cudaMemcpyAsync(dst, frames[f0 + f], fbytes, cudaMemcpyHostToDevice, c->copy_stream));
            else JPGENC_CUDA(c, cudaMemcpy2DAsync(dst, fstride, frames[f0 + f], fbytes, fbytes, run, cudaMemcpyHostToDevice, c->copy_stream));
            f += run;
        }
        JPGENC_CUDA(c, cudaEventRecord(c->ev_band[s], c->copy_stream));
        return JPGENC_OK;
    };
    for (uint32_t p = 0; p < ring; ++p)
        if ((rc = upload(p))) return rc;
    std::vector<std::vector<const void*>> ptrs(nslots);
    int upload_rc = JPGENC_OK;
    rc = run_passes(c, job, per_pass, nslots,
                    [&](uint32_t p, jpgenc_ctx*, const void* const** out, cudaEvent_t* ready) {
                        const uint32_t f0 = p * per_pass, F = std::min(per_pass, job.n - f0), s = p % ring;
                        std::vector<const void*>& v = ptrs[p % nslots];
                        v.resize(F);
                        for (uint32_t f = 0; f < F; ++f) v[f] = c->d_rgb_owned + (static_cast<size_t>(s) * per_pass + f) * fstride;
                        *out = v.data();
                        *ready = c->ev_band[s];
                        return JPGENC_OK;
                    },
                    [&](uint32_t q) {
                        // every kernel of pass q has finished (its event is behind K4): the slice is free again.  The copy
                        // stream must not overtake the kernels that still read it -- it does not: they are done.
                        if (q + ring < npasses && upload_rc == JPGENC_OK) upload_rc = upload(q + ring);
                    });
    if (rc == JPGENC_OK) rc = upload_rc;
    const cudaError_t e = cudaStreamSynchronize(c->copy_stream);
    if (e != cudaSuccess && rc == JPGENC_OK) rc = fail(c, JPGENC_ERR_CUDA, "host-to-device copy of a frame failed");
    c->d_rgb = c->d_rgb_owned;
    return rc;
}

}  // namespace

extern "C" {

int jpgenc_encode_frames_device(jpgenc_ctx* c, uint32_t n, const void* const* dev_frames, uint32_t w, uint32_t h, uint32_t maxval,
                                uint8_t* const* out, const uint64_t* caps, uint64_t* sizes) try {
    Job job;
    job.n = n; job.w = w; job.h = h; job.maxval = maxval;
    job.out_ptrs = out; job.caps = caps; job.sizes = sizes;
    const int rc = check_job(c, job, dev_frames);
    return rc ? rc : encode_device_frames(c, job, dev_frames);
} JPGENC_CATCH(c)

int jpgenc_encode_frames(jpgenc_ctx* c, uint32_t n, const uint8_t* const* frames, uint32_t w, uint32_t h, uint32_t maxval,
                         uint8_t* const* out, const uint64_t* caps, uint64_t* sizes) try {
    Job job;
    job.n = n; job.w = w; job.h = h; job.maxval = maxval;
    job.out_ptrs = out; job.caps = caps; job.sizes = sizes;
    const int rc = check_job(c, job, frames);
    return rc ? rc : encode_host_frames(c, job, frames);
} JPGENC_CATCH(c)

int jpgenc_encode_frames_packed(jpgenc_ctx* c, uint32_t n, const void* const* frames, int frames_on_device, uint32_t w, uint32_t h,
                                uint32_t maxval, uint8_t* out, uint64_t cap, uint64_t* offsets, uint64_t* sizes, uint64_t* total_bytes) try {
    Job job;
    job.n = n; job.w = w; job.h = h; job.maxval = maxval;
    job.packed_mode = true; job.packed = out; job.packed_cap = cap; job.offsets = offsets; job.sizes = sizes;
    int rc = check_job(c, job, frames);
    if (rc) return rc;
    rc = frames_on_device ? encode_device_frames(c, job, frames) : encode_host_frames(c, job, reinterpret_cast<const uint8_t* const*>(frames));
    if (total_bytes) *total_bytes = job.packed_at;
    return rc;
} JPGENC_CATCH(c)

}  // extern "C"
