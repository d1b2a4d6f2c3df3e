// Shared declarations for the sm_100a JPEG encode path (kernels + C-ABI).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <exception>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/jpgenc_b200.h"

namespace jpgenc {

constexpr int kBlocksPerMcu = 6;          // Y00 Y01 Y10 Y11 Cb Cr  (src/Image.cpp:959-967)
constexpr int kCoefPerBlock = 64;
constexpr int kBlockBytes = 128;          // 64 x int16

// zigzag scan: natural index (v*8+u) of the i-th coefficient (include/Coding.hpp:57-81)
#define JPGENC_ZIGZAG_LIST                                                                            \
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, \
    7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, \
    31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63

// Per-table constants of the FP32 fast path, natural order.
//   mul[i] = s_v * s_u / q[i]   (AAN output scales folded with the quantiser; Dct.hpp:36-43, Coding.hpp:92-94)
//   thr[i] = 0.5 - delta_i      a quotient whose distance to the nearest integer exceeds thr is "on a rounding
//                               boundary" as far as FP32 can tell; such blocks are redone in exact FP64
struct QuantConsts {
    float mul[64];
    float thr[64];
};
// the same constants in the pair order of the packed (FP32x2) transform: entry [u*4+q] holds natural indices
// ((2q)*8+u, (2q+1)*8+u)
struct QuantConsts2 {
    float2 mul[32];
    float2 thr[32];
};

// Constants of the exact path: the reference's own doubles (Dct.hpp:21-43) and integer quantisers.
struct ExactConsts {
    double a1, a2, a3, a4, a5;
    double s[8];
    double scale;                         // 255. / maxval  (src/Image.cpp:465)
    uint8_t qy[64];
    uint8_t qc[64];
};

struct ColorConsts {                      // FP32 fast path; sample scale folded in, chroma rows also carry the /4
    float y[3], cb[3], cr[3];
};

struct ForwardParams {
    const uint8_t* rgb;                   // one image ...
    const uint8_t* const* frames;         // ... or (non-null) one device pointer per frame of a batch of equally sized frames
    uint32_t blocks_per_frame;            // frame f owns coefficient blocks [f * blocks_per_frame, (f + 1) * blocks_per_frame)
    int16_t* coef;
    uint32_t* refine_list;                // block ids (mcu*6+k) that need the exact path
    uint32_t* refine_count;
    uint32_t refine_cap;
    uint32_t real_w, real_h;
    uint32_t mcu_w, mcu_h;
    uint32_t mcu_y0;                      // first MCU row of this launch (band-wise launches behind the H2D copies)
    uint32_t prefetch_ahead;              // > 0: every CTA pulls the strip of the CTA this many positions later into L2
    ColorConsts color;
    QuantConsts2 luma;
    QuantConsts2 chroma;
    alignas(64) CUtensorMap coef_map;     // coefficient array as rows of 128 bytes, for the tensor store of full strips
};

__host__ __device__ inline uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }

// device-side Huffman tables, indexed [table][symbol]:
//   entry  (length << 16) | right-aligned code; 0 = absent
//   fast   ((length + cat) << 27) | (code << cat) with cat = symbol & 15: OR the magnitude bits in and the word is the
//          complete (bit count, bits) pair of the symbol; 0 where that does not fit in 27 bits (or the symbol is absent)
struct DeviceTables {
    uint32_t entry[4][256];
    uint32_t fast[4][256];
};

// ---- entropy-stage geometry of a pass (one image = a pass of one frame), device resident -------------------------
// Produced either by the host (single image: the tables are built on the host, so the sizes are known there) or by
// finalize_tables_kernel (batches: tables, sizes and offsets never leave the device); K3a/K3b/K4 read it from device
// memory and take no size from the host.  One block of memory, in this order:
//   PassHeader | raw_off u64[F] | raw_bytes u64[F] | frame_bits u64[F] | file_base u64[F] | k4_tile0 u32[F+1] | hdr_len u32[F]
//   | (8-byte aligned) total_bits u64[F] | ff_incl u64[F]
// The part up to hdr_len is input of K3/K4; total_bits (K3b) and ff_incl (K4) are their results.
struct PassHeader {
    unsigned long long raw_total;    // bytes of `raw` in use: every frame's scan padded to 16 bytes, plus slack
    unsigned long long out_total;    // bytes of the pass's output before byte stuffing: sum of hdr_len + raw_bytes + tail
    unsigned long long raw_sum;      // sum of raw_bytes (an upper bound of the FF bytes K4 can stuff)
    uint32_t k4_tiles;               // K4 tiles of all frames
    uint32_t error;                  // kPass* below; when set, K3a/K3b/K4 return at once
    uint32_t error_frame;
    uint32_t nframes;
    unsigned long long pad[3];
};
static_assert(sizeof(PassHeader) == 64, "PassHeader is copied as 64 bytes");
constexpr uint32_t kPassOk = 0, kPassMissingSymbol = 1, kPassNoSymbols = 2, kPassRawOverflow = 4, kPassOutOverflow = 8;

struct PassMeta {
    PassHeader* hdr;
    unsigned long long *raw_off, *raw_bytes, *frame_bits, *file_base;
    uint32_t *k4_tile0, *hdr_len;
    unsigned long long *total_bits, *ff_incl;
};
__host__ __device__ inline size_t pass_meta_input_bytes(uint32_t F) {
    return (sizeof(PassHeader) + static_cast<size_t>(F) * 32 + (static_cast<size_t>(F) + 1) * 4 + static_cast<size_t>(F) * 4 + 7) & ~static_cast<size_t>(7);
}
__host__ __device__ inline size_t pass_meta_bytes(uint32_t F) { return pass_meta_input_bytes(F) + static_cast<size_t>(F) * 16; }
__host__ __device__ inline PassMeta pass_meta_view(void* base, uint32_t F) {
    uint8_t* b = static_cast<uint8_t*>(base);
    PassMeta m;
    m.hdr = reinterpret_cast<PassHeader*>(b);
    m.raw_off = reinterpret_cast<unsigned long long*>(b + sizeof(PassHeader));
    m.raw_bytes = m.raw_off + F;
    m.frame_bits = m.raw_bytes + F;
    m.file_base = m.frame_bits + F;
    m.k4_tile0 = reinterpret_cast<uint32_t*>(m.file_base + F);
    m.hdr_len = m.k4_tile0 + F + 1;
    m.total_bits = reinterpret_cast<unsigned long long*>(b + pass_meta_input_bytes(F));
    m.ff_incl = m.total_bits + F;
    return m;
}
// a frame's raw (un-stuffed) scan occupies its bytes rounded up to 16, plus slack for the last words K3b touches
__host__ __device__ inline unsigned long long raw_slot_bytes(unsigned long long nbytes) { return ((nbytes + 15) & ~15ull) + 64; }

}  // namespace jpgenc

namespace jpgenc { class TablePool; class HostPool; }

struct jpgenc_ctx {
    int device = 0;
    jpgenc::TablePool* pool = nullptr;    // host worker threads that build the four Huffman tables side by side
    jpgenc::HostPool* host_pool = nullptr;   // parallel table builds of a batch of frames
    bool owns_host_pool = false;             // pipeline lanes share their parent context's pool
    bool parallel_tables = true;          // false inside a batch: there the frames run in parallel instead
    std::vector<jpgenc_ctx*> lanes;       // further pipeline lanes of the batched-frame calls (contexts on the same device with their own
                                          // stream and buffers): while one lane's pass waits for its Huffman tables on the host, the
                                          // other lanes' kernels run
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // host-to-device copies of jpgenc_encode_rgb, overlapped with K1 band by band
    cudaEvent_t ev_band[16] = {};
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_u0 = nullptr, ev_u1 = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;
    // K3/K4 of one image: the time of encode N is read while K2 of encode N + 1 runs (or by jpgenc_get_stats)
    cudaEvent_t ev_e0 = nullptr, ev_e1 = nullptr;
    bool ent_pending = false, stats_pending = false, last_whole = false;
    // Event records between the kernels of an image (jpgenc_set_stage_timing): 0 none -- inside the replayed graphs every
    // record is a node of its own, ~3 us each, 22 us per image with all eight --, 1 around the K1 fast kernel alone (the
    // roofline kernel), 2 around every stage.  *_timed: the level the events of the encode in flight were recorded at.
    int stage_timing = 0, fwd_timed = 0, stats_timed = 0, ent_timed = 0;
    uint32_t k4_grid_hint = 0;            // single image: the grid K4a / K4 are captured with (run_pipeline)
    float last_k1 = 0, last_fwd = 0, last_st = 0;
    std::string error;
    int sm_count = 148;
    uint64_t launches = 0;
    bool capturing = false;               // the stream is being captured into a graph (run_phase): events are recorded as external event nodes
    uint64_t alloc_gen = 0;               // bumped whenever a buffer of the context is (re)allocated or a parameter changes: captured graphs are stale then
    // whole-image encodes of pixels that stay bound: the two GPU phases (K1 .. K2 + publish; tables upload .. K4 + publish) as
    // CUDA graphs, captured when the same configuration is encoded a second time
    struct PhaseGraph {
        cudaGraphExec_t exec = nullptr;
        uint64_t key[5] = {0, 0, 0, 0, 0};
        uint64_t seen[5] = {0, 0, 0, 0, 0};   // key of the previous encode (a graph is captured when it repeats)
        uint32_t launches = 0;                // kernels per replay
        uint32_t publishes = 0;               // mailbox announcements per replay
    } graph_a, graph_b;

    // parameters
    uint8_t qy[64], qc[64];
    double dct_a[5], dct_s[8];

    // image state
    uint32_t nframes = 1;                 // > 1: a batch of equally sized frames goes through every kernel at once
    const uint8_t** d_frame_ptrs = nullptr;   // [nframes] device pointers of the frames (batches only)
    size_t frame_ptrs_cap = 0;
    bool frames_aligned = false;          // every frame pointer is 16-byte aligned (bulk-copy path of K1)
    const uint8_t* d_rgb = nullptr;       // bound or owned
    uint8_t* d_rgb_owned = nullptr;
    size_t rgb_cap = 0;
    uint32_t real_w = 0, real_h = 0, maxval = 255, mcu_w = 0, mcu_h = 0;
    bool have_pixels = false, have_coef = false, have_scan = false, forward_pending = false, upload_pending = false;

    int16_t* d_coef = nullptr;
    size_t coef_cap = 0;
    uint32_t* d_refine_list = nullptr;
    size_t refine_cap = 0;
    uint32_t* d_counters = nullptr;       // kCounterWords words, see kCnt* below
    uint8_t* d_stats = nullptr;           // per frame: hist u32[4][256] + first-occurrence keys u64[4][256]; then a copy of the refine counter
    size_t stats_cap = 0;
    uint8_t* d_meta = nullptr;            // PassMeta block of the frames bound to the context (lives behind d_tables, same allocation)
    uint32_t entropy_runs = 0;            // K3/K4 passes over the current symbol items (reset by K2)
    bool file_mode = false;               // K4 writes complete files (header + scan + EOI per frame) instead of bare scans
    uint8_t* d_hdr_prefix = nullptr;      // SOI .. SOF0 of the frames of a batch (identical for all of them)
    uint32_t hdr_prefix_len = 0;
    uint64_t raw_limit = 0, out_limit = 0;   // what finalize_tables_kernel may use of d_raw / d_scan
    uint64_t batch_raw_per_frame = 0;     // batched calls: raw-scan bytes reserved per frame (1.5 x the largest seen so far; 0 = nothing seen yet)
    uint32_t batch_geom_w = 0, batch_geom_h = 0;   // ... the frame size that figure was learnt on (kept across calls of one size)
    cudaEvent_t ev_done = nullptr;        // batched calls: end of the slot's current pass
    cudaEvent_t ev_wide = nullptr;        // ... its K1/refinement/K2 are through (JPGENC_STAGGER=1: the next pass's wide kernels wait for it)
    cudaEvent_t ev_k4 = nullptr;          // ... its K4 is through (the next pass's K3 waits for it: passes finish one after the other)
    cudaEvent_t ev_fwd = nullptr;         // ... its K1/refinement are through (JPGENC_STAGGER=2: the next pass's K1 waits for it)
    cudaEvent_t ev_copied = nullptr;      // ... its files have left d_scan (recorded on out_stream)
    cudaStream_t out_stream = nullptr;    // device-to-host copies of finished passes, beside the slot's kernels
    bool copy_pending = false;
    uint8_t hdr_prefix_host[256] = {};    // what d_hdr_prefix holds
    uint32_t debug_tables_valid = 0;      // JPGENC_DEBUG_SKIP_TABLES (development): tables in d_built_tables from an earlier call
    uint32_t* d_items = nullptr;          // K2's symbol stream (blockwalk.cuh), consumed by K3
    size_t items_cap = 0;
    uint32_t* d_tile_cnt = nullptr;             // per K2 tile: number of items in its slab
    size_t tile_cnt_cap = 0;
    uint32_t* d_range_bits = nullptr;           // per K2 tile: bits its items encode to (K3a)
    size_t range_bits_cap = 0;
    unsigned long long* d_range_base = nullptr; // bits per group of 8 tiles, then per 256 groups (K3a); K3b derives offsets from them
    size_t range_base_cap = 0;
    bool have_items = false;
    uint32_t k2_tiles_done = 0;           // K2 tiles already processed behind the bands of an upload (jpgenc_encode_rgb)
    std::vector<uint32_t> host_hist;      // K2's histograms as last read back, [nframes][4][256]
    std::vector<uint64_t> frame_bits, frame_out_off, frame_ff;   // per frame after K3/K4: scan bits, byte offset of its stuffed scan in d_scan, stuffed FFs
    bool k2_configured = false;
    // launch_forward_rows(first) clears K1's counters AND what K2 / K3a accumulate into in one kernel; launch_symbol_stats(first)
    // then skips its own clears -- valid while no buffer was reallocated and the frame count is the same
    bool stats_clear_valid = false;
    uint64_t stats_clear_gen = 0;
    uint32_t stats_clear_frames = 0;
    void* d_tab_scratch = nullptr;        // device-side table build (tables_device.cu): work space, one slab per table
    size_t tab_scratch_cap = 0;
    jpgenc_huff_table* d_built_tables = nullptr;   // ... its results [4 * nframes], and one status word per table behind them
    size_t built_tables_cap = 0;
    jpgenc::DeviceTables* d_tables = nullptr;   // [nframes]
    size_t tables_cap = 0;
    unsigned long long* d_lookback = nullptr;  // one look-back word per K4 tile
    size_t lookback_cap = 0;
    uint32_t* d_raw = nullptr;            // un-stuffed scan, 32-bit words, bytes in stream order
    size_t raw_cap = 0;
    uint8_t* d_scan = nullptr;            // K4's output: the stuffed scan, or (file_mode) the complete files of a pass back to back
    size_t scan_cap = 0;
    void* d_flush = nullptr;              // L2 eviction scratch
    size_t flush_bytes = 0;
    void* h_file_pinned = nullptr;        // two band-sized pinned buffers for streamed inputs (jpgenc_encode_ppm_file)
    size_t file_pinned_bytes = 0;
    // mailbox: mapped pinned memory the kernels of ONE image write their results into (the host polls it instead of
    // synchronising the stream); layout: kMail* below
    unsigned long long* h_mailbox = nullptr;   // host view
    unsigned long long* d_mailbox = nullptr;   // device view of the same memory
    uint32_t stat_count[4][256];          // K2's statistics of the current image as last received
    uint64_t stat_first[4][256];
    uint32_t mailbox_seq = 0;             // announced by K2 / K4 of the current encode
    void* h_pinned = nullptr;             // small pinned staging (stats, totals)
    size_t pinned_bytes = 0;

    std::vector<jpgenc_huff_table> pass_tables;   // batched-frame calls: tables of the lane's current pass
    std::vector<const void*> pass_ptrs;           // ... device pointers of its frames
    double trace_convert_us = 0;
    std::vector<uint32_t> hdr_len;                // ... header length of every frame's file

    jpgenc_stats stats{};
    jpgenc_huff_table last_tables[4];     // tables of the last whole-image run (for jpgenc_assemble_last)
    bool have_tables = false;
};

// event on the context's stream; inside a stream capture it becomes an event-record NODE of the graph (a plain record there
// only marks a dependency and leaves an event that cannot be waited for or timed)
inline cudaError_t jpgenc_record(jpgenc_ctx* c, cudaEvent_t ev) {
    return cudaEventRecordWithFlags(ev, c->stream, c->capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
}

// an event of the per-stage timing: recorded only when stage timing is on at `level` or above
inline cudaError_t stage_record(jpgenc_ctx* c, cudaEvent_t ev, int level) {
    return c->stage_timing >= level ? jpgenc_record(c, ev) : cudaSuccess;
}

#define JPGENC_CUDA(ctx, expr)                                                                       \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            (ctx)->error = std::string(#expr) + ": " + cudaGetErrorString(e__);                      \
            return JPGENC_ERR_CUDA;                                                                  \
        }                                                                                            \
    } while (0)

// No exception crosses the C boundary (include/jpgenc_b200.h): every entry point that takes a context is a function-try-block
// closed by this.  What can arrive here is std::bad_alloc / std::length_error from a host container.
#define JPGENC_CATCH(ctx)                                                                            \
    catch (const std::exception& ex__) {                                                             \
        if (ctx) (ctx)->error = std::string("host exception: ") + ex__.what();                      \
        return JPGENC_ERR_NOMEM;                                                                     \
    } catch (...) {                                                                                  \
        if (ctx) (ctx)->error = "host exception";                                                   \
        return JPGENC_ERR_NOMEM;                                                                     \
    }

// launchers implemented in the kernel translation units
namespace jpgenc {
// words of jpgenc_ctx::d_counters
constexpr int kCntRefine = 0;      // entries in the refinement list (K1)
constexpr int kCntRefined = 3;     // list entries already refined (band-wise encodes)
constexpr int kCntFinalize = 4;    // finalize_tables_kernel's CTA ticket (resets itself)
constexpr int kCntK3Redo = 6;     // chunks K3b packed the slow way (zeroed by K3a; read by tests through jpgenc_debug_counter)
constexpr int kCntSeq = 5;         // mailbox sequence number (mailbox_publish_kernel increments it; never cleared)
constexpr int kCounterWords = 16;
// The host mailbox (jpgenc_ctx::h_mailbox, mapped pinned memory) as 64-bit words.  Every word carries its own validity: value
// (42 bits) | tag << 42, tag = mail_tag(sequence number of the publishing kernel) != 0 -- an aligned 8-byte store arrives
// whole, so the host needs no flag behind the data and the kernel no system-wide fence in front of one; the host zeroes what
// it has consumed, so a tag is never met again.
//   [kMailStatsHead]           n present symbols (11 bits) | K1's refinement counter << 11
//   [kMailStatsHead + 1 + 2k]  record k: count (32 bits) | (table * 256 + symbol) << 32
//   [kMailStatsHead + 2 + 2k]            first-occurrence key
//   [kMailTotals], [+1]        scan bits, stuffed FF bytes (K4)
constexpr int kMailStatsHead = 0;
constexpr int kMailTotals = 2064;
constexpr int kMailWords64 = 2080;
constexpr int kMailTagShift = 42;
__host__ __device__ inline unsigned long long mail_tag(uint32_t seq) { return static_cast<unsigned long long>(seq % 0x1FFFFFu + 1u); }
__host__ __device__ inline unsigned long long mail_word(unsigned long long value, uint32_t seq) { return value | (mail_tag(seq) << kMailTagShift); }
int launch_forward(jpgenc_ctx* c);
int launch_forward_rows(jpgenc_ctx* c, uint32_t y0, uint32_t rows, bool first, bool last);
int launch_dct_quant_blocks(jpgenc_ctx* c, const float* in, int16_t* out, uint64_t nblocks, const uint8_t q[64],
                            uint64_t* refined);
int launch_planes_to_mcu(jpgenc_ctx* c, const int32_t* d_qy, const int32_t* d_qcb, const int32_t* d_qcr);
int launch_refine_pending(jpgenc_ctx* c);
int launch_planes_exact(jpgenc_ctx* c, const double* d_planes, bool ycbcr);
// K2 over the tiles [tile0, tile0 + ntiles) of the bound image(s); `first` also clears the statistics
int launch_symbol_stats(jpgenc_ctx* c, uint32_t tile0, uint32_t ntiles, bool first);
// K1's counters, the statistics of all frames bound and K3a's sums in one launch (false: the statistics' buffers do not exist yet)
bool launch_encode_clear(jpgenc_ctx* c);
int launch_entropy(jpgenc_ctx* c, uint32_t k4_grid);
int launch_publish_stats(jpgenc_ctx* c);
// batches: built tables + K2's histograms -> DeviceTables and the PassMeta block, all on the device
int launch_finalize_tables(jpgenc_ctx* c);
#ifndef JPGENC_K4_THREADS
#define JPGENC_K4_THREADS 256
#endif
constexpr uint32_t kK4TileBytes = JPGENC_K4_THREADS * 16;   // input bytes per K4 tile: 16 per thread (entropy.cu)
size_t table_scratch_bytes();
int build_table_arrays_host(const uint32_t count[256], const uint64_t first_pos[256], jpgenc_huff_table* out);
int launch_build_tables(jpgenc_ctx* c, const uint8_t* d_stats, uint32_t stats_stride, uint32_t ntables, void* d_scratch,
                        jpgenc_huff_table* d_out, uint32_t* d_status);
int launch_synth_rgb(jpgenc_ctx* c, uint8_t* d, uint32_t w, uint32_t h, uint32_t seed);
int launch_synth_blocks(jpgenc_ctx* c, float* d, uint64_t nblocks);
int launch_flush(jpgenc_ctx* c);
void fill_quant_consts(const uint8_t q[64], const double s[8], QuantConsts* out, double input_err);
void default_dct_constants(double a[5], double s[8]);
}  // namespace jpgenc
