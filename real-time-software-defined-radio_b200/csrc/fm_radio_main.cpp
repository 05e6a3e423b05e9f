// fm_radio — drop-in for the reference executable (process contract of /root/reference/src/fm_radio.cpp:732-798):
//   argv   none -> mode 0 ; "1" -> mode 1 ; anything else -> message on stderr, exit 1 (:736-764, including "0")
//   stdin  raw 8-bit unsigned interleaved IQ, consumed in 307200-byte blocks (:23,:66)
//   stdout headerless little-endian int16, interleaved L,R, 48 kHz, one write per block (:286-302)
//   stderr the reference's diagnostics: argc, mode line, rf_Fs, and in mode 0 the frame_thread lines (:516,:619-701)
// Extensions (after the mode argument, all optional): --profile binary|intent (default binary = byte-compatible with
// the shipped executable), --blocks N (blocks per GPU call, default 1), --device D, --quiet,
// --audio-rate 44100 (mode 0 only: audio through the x147 /800 polyphase resampler, 2822 samples per block; the 44.1 kHz
// mode of the project the reference never implemented), --rds-info (PI / PS / RadioText from the decoded bits, at exit).
// EOF handling is normalised (Q9): only whole blocks are processed.
//
// The reference's four threads and three bounded queues are replaced by: a reader thread filling a ring of pinned host
// slots, the GPU pipeline behind fmrx_batch_process (copy-in stream / compute streams / copy-out stream), and the main
// thread draining results to stdout/stderr.  Back-pressure is a real bounded ring (while-loops on the condition
// variables), not the reference's `if`-guarded waits.
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "fmrx.h"

namespace {
constexpr int kSlots = 4;  // QUEUE_BLOCKS - 1 of the reference (:22)

struct Ring {
    std::mutex m;
    std::condition_variable cv;
    int filled[kSlots] = {0};  // blocks in the slot, -1 = end of stream
    bool ready[kSlots] = {false};
};

size_t read_fully(uint8_t *dst, size_t n) {
    size_t got = 0;
    while (got < n) {
        size_t r = fread(dst + got, 1, n - got, stdin);
        if (r == 0) break;
        got += r;
    }
    return got;
}
}  // namespace

int main(int argc, char *argv[]) {
    int mode = 0, lib_mode = -1, profile = FMRX_PROFILE_BINARY, blocks = 1, device = 0;
    bool quiet = false, rds_info = false;
    int pos = 1;
    std::cerr << ((argc >= 2 && argv[1][0] != '-') ? 2 : 1) << std::endl;  // the reference prints argc first (:738); options are not counted
    if (argc >= 2 && argv[1][0] != '-') {
        mode = atoi(argv[1]);
        if (mode != 1) {
            std::cerr << "Wrong mode " << mode << std::endl;  // :750
            return 1;
        }
        pos = 2;
    }
    for (; pos < argc; ++pos) {
        std::string a = argv[pos];
        auto next = [&]() -> const char * { return pos + 1 < argc ? argv[++pos] : ""; };
        if (a == "--profile") profile = std::string(next()) == "intent" ? FMRX_PROFILE_INTENT : FMRX_PROFILE_BINARY;
        else if (a == "--blocks") blocks = std::max(1, atoi(next()));
        else if (a == "--device") device = atoi(next());
        else if (a == "--quiet") quiet = true;
        else if (a == "--audio-rate") {  // extension: 44100 selects the x147 /800 audio resampler (library mode 2); only with the 2.4 Msps front end
            const int rate = atoi(next());
            if ((rate != 48000 && rate != 44100) || (rate == 44100 && mode != 0)) { std::cerr << "Usage " << argv[0] << std::endl; return 1; }
            if (rate == 44100) lib_mode = 2;
        }
        else if (a == "--rds-info") rds_info = true;  // extension: decode the RDS groups (PI, PS, RadioText) and report them at the end
        else { std::cerr << "Usage " << argv[0] << std::endl; return 1; }  // :762
    }
    std::cerr << "Operating in mode " << mode << std::endl;            // :741,:756
    std::cerr << "rf_Fs = " << (mode == 1 ? 2500000 : 2400000) << std::endl;  // :61

    fmrx_config cfg{};
    cfg.mode = lib_mode >= 0 ? lib_mode : mode; cfg.profile = profile; cfg.n_streams = 1; cfg.max_blocks = blocks; cfg.device = device;
    fmrx_batch *rx = nullptr;
    if (fmrx_batch_create(&cfg, &rx) != FMRX_OK) {
        std::cerr << "fm_radio: " << fmrx_last_error() << std::endl;
        return 2;
    }
    const int na = fmrx_batch_audio_per_block(rx);
    const size_t slot_bytes = (size_t)blocks * FMRX_BLOCK_BYTES;
    uint8_t *slots[kSlots];
    for (auto &s : slots)
        if (fmrx_pinned_alloc((void **)&s, slot_bytes) != FMRX_OK) { std::cerr << "fm_radio: " << fmrx_last_error() << std::endl; return 2; }
    int16_t *audio = nullptr;
    fmrx_pinned_alloc((void **)&audio, (size_t)blocks * 2 * na * sizeof(int16_t));
    std::vector<fmrx_rds_event> ev((size_t)blocks * FMRX_MAX_EVENTS);
    std::vector<int32_t> nev(blocks);
    std::vector<uint8_t> bits((size_t)blocks * FMRX_MAX_BITS);
    std::vector<int32_t> nbits(blocks);
    fmrx_rds_app *app = nullptr;
    if (rds_info && mode == 0 && fmrx_rds_app_create(1, &app) != FMRX_OK) { std::cerr << "fm_radio: " << fmrx_last_error() << std::endl; return 2; }

    Ring ring;
    std::thread reader([&] {
        for (int i = 0;; i = (i + 1) % kSlots) {
            {
                std::unique_lock<std::mutex> lk(ring.m);
                ring.cv.wait(lk, [&] { return !ring.ready[i]; });
            }
            const size_t got = read_fully(slots[i], slot_bytes);
            const int nb = (int)(got / FMRX_BLOCK_BYTES);  // whole blocks only (Q9)
            {
                std::lock_guard<std::mutex> lk(ring.m);
                ring.filled[i] = nb > 0 ? nb : -1;
                ring.ready[i] = true;
            }
            ring.cv.notify_all();
            if (got < slot_bytes) {
                if (nb > 0) {  // publish the end marker in the next slot
                    const int j = (i + 1) % kSlots;
                    std::unique_lock<std::mutex> lk(ring.m);
                    ring.cv.wait(lk, [&] { return !ring.ready[j]; });
                    ring.filled[j] = -1;
                    ring.ready[j] = true;
                    ring.cv.notify_all();
                }
                return;
            }
        }
    });

    long long block_id = 0;
    int rc = 0;
    std::vector<int32_t> offset(1, 0);
    for (int i = 0;; i = (i + 1) % kSlots) {
        int nb;
        {
            std::unique_lock<std::mutex> lk(ring.m);
            ring.cv.wait(lk, [&] { return ring.ready[i]; });
            nb = ring.filled[i];
        }
        if (nb < 0) break;
        fmrx_outputs out{};
        out.audio = audio;
        if (mode == 0) { out.rds_events = ev.data(); out.rds_n_events = nev.data(); }
        if (app) { out.rds_bits = bits.data(); out.rds_n_bits = nbits.data(); }
        if (fmrx_batch_process(rx, slots[i], nb, &out) != FMRX_OK) {
            std::cerr << "fm_radio: " << fmrx_last_error() << std::endl;
            rc = 2;
            break;
        }
        fwrite(audio, sizeof(int16_t), (size_t)nb * 2 * na, stdout);  // :302
        if (app) fmrx_rds_app_feed(app, bits.data(), nbits.data(), nb, nullptr, 0, nullptr);
        if (mode == 0 && !quiet) {
            if (block_id == 0) fmrx_batch_rds_offsets(rx, offset.data());
            char text[8192];
            for (int b = 0; b < nb; ++b) {
                fmrx_rds_format_block((int)(block_id + b), offset[0], ev.data() + (size_t)b * FMRX_MAX_EVENTS, nev[b], text, sizeof(text));
                std::cerr << text;
            }
        }
        block_id += nb;
        {
            std::lock_guard<std::mutex> lk(ring.m);
            ring.ready[i] = false;
        }
        ring.cv.notify_all();
    }
    fflush(stdout);
    if (rc != 0) { fclose(stdin); }
    reader.join();
    for (auto &s : slots) fmrx_pinned_free(s);
    fmrx_pinned_free(audio);
    fmrx_batch_destroy(rx);
    if (app) {
        fmrx_rds_station st;
        if (fmrx_rds_app_station(app, 0, &st) == FMRX_OK) {
            char line[256];
            snprintf(line, sizeof(line), "RDS: PI %04X PTY %d TP %d PS \"%s\" RT \"%s\" groups %u blocks ok/corrected/bad %u/%u/%u", st.pi < 0 ? 0 : st.pi, st.pty, st.tp, st.ps,
                     st.rt, st.groups, st.blocks_ok, st.blocks_corrected, st.blocks_bad);
            std::cerr << line << std::endl;
        }
        fmrx_rds_app_destroy(app);
    }
    if (rc == 0) std::cerr << "Run: gnuplot -e 'set terminal png size 1024,768' example.gnuplot > ../data/example.png" << std::endl;  // :795
    return rc;
}
