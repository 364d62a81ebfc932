// Ingest front ends, host side only. The library does no device I/O: whoever owns
// the rtl_tcp socket or the librtlsdr handle pushes the bytes it receives through these two pieces
// and hands the resulting callback blocks to sdrb_bank_process_host.
//
//   sdrb_rtltcp_*  the rtl_tcp client protocol of sdrj.cpp:31-74,125-188: 12-byte "RTL0" dongle
//                  header, then uint8 IQ cut into callback blocks; 5-byte big-endian commands.
//   sdrb_ring_*    the 20-buffer hand-over between the librtlsdr callback thread and the demodulator
//                  thread (jonti/sdr.cpp:100-184, jonti/sdr.h:83-99): a full ring DROPS the new buffer.
//                  Buffers hold the raw bytes (the byte -> float step of sdr.cpp:122-129 happens on
//                  the GPU) and can be pinned, so the H2D copy reads them in place.
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <vector>

#include "plan.hpp"

using sdrb::set_error;

// ------------------------------------------------------------------ rtl_tcp framing
struct sdrb_rtltcp {
    size_t block_bytes = 0;
    bool header_done = false, has_header = false;
    uint32_t tuner_type = 0, gain_count = 0;
    std::vector<uint8_t> head;          // first bytes of the stream until the header question is settled
    std::vector<uint8_t> partial;       // bytes of the block being filled
    std::deque<std::vector<uint8_t>> ready;
    uint64_t bytes_in = 0;
};

extern "C" int sdrb_rtltcp_create(int sample_rate, size_t block_bytes, sdrb_rtltcp **out) {
    if (!out || sample_rate <= 0) { set_error("sdrb_rtltcp_create: bad argument"); return SDRB_E_INVALID; }
    sdrb_rtltcp *f = new (std::nothrow) sdrb_rtltcp();
    if (!f) { set_error("out of memory"); return SDRB_E_NOMEM; }
    // sdrj.cpp:46: tcpFloats.resize((samplerate/4)*2) -- the reference cuts the TCP stream into quarter-second
    // buffers whatever the plan's callback size is; a caller whose plan uses 5 callbacks/s (288 kS/s) passes its own size
    f->block_bytes = block_bytes ? block_bytes : (size_t)(sample_rate / 4) * 2;
    f->partial.reserve(f->block_bytes);
    *out = f;
    return SDRB_OK;
}

extern "C" void sdrb_rtltcp_destroy(sdrb_rtltcp *f) { delete f; }

extern "C" size_t sdrb_rtltcp_block_bytes(const sdrb_rtltcp *f) { return f ? f->block_bytes : 0; }

static void rtltcp_data(sdrb_rtltcp *f, const uint8_t *p, size_t n) {
    while (n) {
        const size_t take = std::min(n, f->block_bytes - f->partial.size());
        f->partial.insert(f->partial.end(), p, p + take);
        p += take; n -= take;
        if (f->partial.size() == f->block_bytes) {
            f->ready.emplace_back(std::move(f->partial));
            f->partial.clear();
            f->partial.reserve(f->block_bytes);
        }
    }
}

// Returns the number of complete blocks waiting. rtl_tcp sends the 12-byte dongle header
// ("RTL0", tuner type, tuner gain count, big endian; sdrj.cpp:134-149) before any sample.
extern "C" int sdrb_rtltcp_feed(sdrb_rtltcp *f, const uint8_t *bytes, size_t n) {
    if (!f || (!bytes && n)) { set_error("sdrb_rtltcp_feed: bad argument"); return SDRB_E_INVALID; }
    f->bytes_in += n;
    if (!f->header_done) {
        f->head.insert(f->head.end(), bytes, bytes + n);
        if (f->head.size() < 4 && memcmp(f->head.data(), "RTL0", f->head.size()) == 0) return (int)f->ready.size();
        if (f->head.size() >= 4 && memcmp(f->head.data(), "RTL0", 4) == 0) {
            if (f->head.size() < 12) return (int)f->ready.size();
            const uint8_t *b = f->head.data();
            f->tuner_type = (uint32_t)b[4] << 24 | (uint32_t)b[5] << 16 | (uint32_t)b[6] << 8 | b[7];
            f->gain_count = (uint32_t)b[8] << 24 | (uint32_t)b[9] << 16 | (uint32_t)b[10] << 8 | b[11];
            f->has_header = true;
            f->header_done = true;
            rtltcp_data(f, b + 12, f->head.size() - 12);
        } else {                        // a source without the header (a recorded stream): everything is samples
            f->header_done = true;
            rtltcp_data(f, f->head.data(), f->head.size());
        }
        f->head.clear();
        f->head.shrink_to_fit();
        return (int)f->ready.size();
    }
    rtltcp_data(f, bytes, n);
    return (int)f->ready.size();
}

extern "C" int sdrb_rtltcp_header(const sdrb_rtltcp *f, uint32_t *tuner_type, uint32_t *gain_count) {
    if (!f || !f->has_header) return 0;
    if (tuner_type) *tuner_type = f->tuner_type;
    if (gain_count) *gain_count = f->gain_count;
    return 1;
}

extern "C" int sdrb_rtltcp_pop(sdrb_rtltcp *f, uint8_t *dst) {
    if (!f || !dst) { set_error("sdrb_rtltcp_pop: bad argument"); return SDRB_E_INVALID; }
    if (f->ready.empty()) return 0;
    memcpy(dst, f->ready.front().data(), f->block_bytes);
    f->ready.pop_front();
    return 1;
}

// sdrj::sendCommand (sdrj.cpp:168-188): command byte, then the 32-bit value most significant byte first.
extern "C" void sdrb_rtltcp_command(uint8_t cmd, uint32_t value, uint8_t out[5]) {
    out[0] = cmd;
    out[1] = (uint8_t)(value >> 24); out[2] = (uint8_t)(value >> 16); out[3] = (uint8_t)(value >> 8); out[4] = (uint8_t)value;
}

// The five commands sdrj::start_tcp_rtl sends after connecting (sdrj.cpp:56-66): AGC off, manual
// tuner gain, gain index, sample rate, centre frequency. 25 bytes.
extern "C" int sdrb_rtltcp_start_sequence(int sample_rate, int frequency, int gain_index, uint8_t out[25]) {
    if (!out) { set_error("sdrb_rtltcp_start_sequence: NULL"); return SDRB_E_INVALID; }
    sdrb_rtltcp_command(SDRB_RTLTCP_SET_AGC_MODE, 0, out);
    sdrb_rtltcp_command(SDRB_RTLTCP_SET_TUNER_GAIN_MODE, 1, out + 5);
    sdrb_rtltcp_command(SDRB_RTLTCP_SET_TUNER_GAIN_INDEX, (uint32_t)gain_index, out + 10);
    sdrb_rtltcp_command(SDRB_RTLTCP_SET_SAMPLE_RATE, (uint32_t)sample_rate, out + 15);
    sdrb_rtltcp_command(SDRB_RTLTCP_SET_FREQ, (uint32_t)frequency, out + 20);
    return 25;
}

// ------------------------------------------------------------------ callback ring
struct sdrb_ring {
    size_t block_bytes = 0;
    int n = 0;
    bool pinned = false;
    std::vector<uint8_t *> buf;
    std::vector<uint32_t> valid;        // buffers_size_valid
    int head = 0, tail = 0, used = 0;   // buffers_head_ptr / buffers_tail_ptr / buffers_used
    bool cancel = false, popped = false;
    uint64_t dropped = 0, pushed = 0;
    std::mutex mut;
    std::condition_variable not_empty;
};

extern "C" void sdrb_ring_destroy(sdrb_ring *r) {
    if (!r) return;
    for (uint8_t *p : r->buf) {
        if (!p) continue;
        if (r->pinned) sdrb_host_free(p); else delete[] p;
    }
    delete r;
}

extern "C" int sdrb_ring_create(size_t block_bytes, int n_buffers, int pinned, sdrb_ring **out) {
    if (!out || block_bytes == 0 || n_buffers < 0) { set_error("sdrb_ring_create: bad argument"); return SDRB_E_INVALID; }
    sdrb_ring *r = new (std::nothrow) sdrb_ring();
    if (!r) { set_error("out of memory"); return SDRB_E_NOMEM; }
    r->block_bytes = block_bytes;
    r->n = n_buffers ? n_buffers : 20;                       // N_BUFFERS, jonti/sdr.h:83
    r->pinned = pinned != 0;
    r->buf.assign((size_t)r->n, nullptr);
    r->valid.assign((size_t)r->n, 0);
    for (int i = 0; i < r->n; i++) {
        r->buf[(size_t)i] = r->pinned ? (uint8_t *)sdrb_host_alloc(block_bytes) : new (std::nothrow) uint8_t[block_bytes];
        if (!r->buf[(size_t)i]) {
            set_error(r->pinned ? "sdrb_ring_create: pinned allocation failed (no CUDA device?)" : "out of memory");
            sdrb_ring_destroy(r);
            return r->pinned ? SDRB_E_CUDA : SDRB_E_NOMEM;
        }
    }
    *out = r;
    return SDRB_OK;
}

// sdr::rtlsdr_callback (jonti/sdr.cpp:100-145): 1 = queued, 0 = ring full, buffer dropped.
extern "C" int sdrb_ring_push(sdrb_ring *r, const uint8_t *bytes, uint32_t len) {
    if (!r || !bytes || len > r->block_bytes) { set_error("sdrb_ring_push: bad argument or buffer longer than block_bytes"); return SDRB_E_INVALID; }
    int slot;
    {
        std::lock_guard<std::mutex> g(r->mut);
        if (r->used >= r->n) { r->dropped++; return 0; }     // "Dropped RTL buffer!!"
        r->head %= r->n;
        slot = r->head;
    }
    memcpy(r->buf[(size_t)slot], bytes, len);                // only the producer touches a free slot
    {
        std::lock_guard<std::mutex> g(r->mut);
        r->valid[(size_t)slot] = len;
        r->used++;
        r->head++;
        r->pushed++;
    }
    r->not_empty.notify_all();
    return 1;
}

// sdr::demod_dispatcher (jonti/sdr.cpp:147-184): wait for a buffer; it stays owned by the consumer
// until sdrb_ring_release (the reference decrements buffers_used after the slot has been handled).
extern "C" int sdrb_ring_pop(sdrb_ring *r, const uint8_t **bytes, uint32_t *len, int timeout_ms) {
    if (!r || !bytes || !len) { set_error("sdrb_ring_pop: bad argument"); return SDRB_E_INVALID; }
    std::unique_lock<std::mutex> g(r->mut);
    if (r->popped) { set_error("sdrb_ring_pop: release the previous buffer first"); return SDRB_E_INVALID; }
    auto ready = [r] { return r->used > 0 || r->cancel; };
    if (timeout_ms < 0) r->not_empty.wait(g, ready);
    else if (!r->not_empty.wait_for(g, std::chrono::milliseconds(timeout_ms), ready)) return 0;
    if (r->cancel) return 0;
    r->tail %= r->n;
    *bytes = r->buf[(size_t)r->tail];
    *len = r->valid[(size_t)r->tail];
    r->popped = true;
    return 1;
}

extern "C" int sdrb_ring_release(sdrb_ring *r) {
    if (!r) { set_error("sdrb_ring_release: NULL"); return SDRB_E_INVALID; }
    std::lock_guard<std::mutex> g(r->mut);
    if (!r->popped) { set_error("sdrb_ring_release: nothing to release"); return SDRB_E_INVALID; }
    r->popped = false;
    r->tail++;
    r->used--;
    return SDRB_OK;
}

extern "C" void sdrb_ring_cancel(sdrb_ring *r) {             // do_demod_dispatcher_cancel
    if (!r) return;
    { std::lock_guard<std::mutex> g(r->mut); r->cancel = true; }
    r->not_empty.notify_all();
}

extern "C" int sdrb_ring_stats(sdrb_ring *r, uint64_t *pushed, uint64_t *dropped, int *used) {
    if (!r) { set_error("sdrb_ring_stats: NULL"); return SDRB_E_INVALID; }
    std::lock_guard<std::mutex> g(r->mut);
    if (pushed) *pushed = r->pushed;
    if (dropped) *dropped = r->dropped;
    if (used) *used = r->used;
    return SDRB_OK;
}
