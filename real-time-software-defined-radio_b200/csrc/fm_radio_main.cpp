// fm_radio — drop-in for the reference executable (process contract of /root/reference/src/fm_radio.cpp:732-798):
//   argv   none -> mode 0 ; "1" -> mode 1 ; anything else -> message on stderr, exit 1 (:736-764, including "0")
//   stdin  raw 8-bit unsigned interleaved IQ, consumed in 307200-byte blocks (:23,:66)
//   stdout headerless little-endian int16, interleaved L,R, 48 kHz, one write per block (:286-302)
//   stderr the reference's diagnostics: argc, mode line, rf_Fs, and in mode 0 the frame_thread lines (:516,:619-701)
// Extensions (after the mode argument, all optional): --profile binary|intent (default binary = byte-compatible with
// the shipped executable), --blocks N (blocks per GPU call, default 1), --device D, --quiet,
// --numerics strict|reference|fma (include/fmrx.h; strict, with the stage-by-stage RDS back end, is the default HERE),
// --audio-rate 44100 (mode 0 only: audio through the x147 /800 polyphase resampler, 2822 samples per block; the 44.1 kHz
// mode of the project the reference never implemented), --rds-info (PI / PS / RadioText from the decoded bits, at exit).
// EOF handling is normalised (Q9): only whole blocks are processed.
//
// The reference's four threads and three bounded queues (:86-138, :212-223, :376-387, :414-423) are replaced by the
// library's ingest / egress ring (fmrx_ring_*, csrc/fmrx_ring.cpp): a reader thread acquires a page-locked slot, fills
// it from stdin and commits it (H2D copy + the three-phase pipeline + D2H copy, all asynchronous), the main thread
// takes the finished steps in order and writes stdout / stderr.  Block k+1 is read and copied in while block k is on the
// GPU; a full ring blocks the reader (back-pressure), a slow stdout blocks the main thread and, through the ring, stdin.
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "fmrx.h"

namespace {
constexpr int kSlots = 4;  // QUEUE_BLOCKS - 1 of the reference (:22)

size_t read_fully(uint8_t *dst, size_t n) {
    size_t got = 0;
    while (got < n) {
        const ssize_t r = read(0, dst + got, n - got);  // unbuffered: a block is handed on as soon as its last byte has arrived
        if (r <= 0) break;
        got += (size_t)r;
    }
    return got;
}
}  // namespace

int main(int argc, char *argv[]) {
    int mode = 0, lib_mode = -1, profile = FMRX_PROFILE_BINARY, blocks = 1, device = 0, numerics = FMRX_NUMERICS_STRICT;
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
        else if (a == "--numerics") {  // extension: strict (default here) | reference | fma, see include/fmrx.h
            const std::string v = next();
            if (v == "reference") numerics = FMRX_NUMERICS_REFERENCE;
            else if (v == "strict") numerics = FMRX_NUMERICS_STRICT;
            else if (v == "fma") numerics = FMRX_NUMERICS_FMA;
            else { std::cerr << "Usage " << argv[0] << std::endl; return 1; }
        }
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

    auto die = [](const char *what) {
        // fatal: the reader may sit in read(0) for ever, so the process leaves without joining it
        std::cerr << "fm_radio: " << what << ": " << fmrx_last_error() << std::endl;
        fflush(stdout);
        _exit(2);
    };
    fmrx_config cfg{};
    cfg.mode = lib_mode >= 0 ? lib_mode : mode; cfg.profile = profile; cfg.n_streams = 1; cfg.max_blocks = blocks; cfg.device = device; cfg.numerics = numerics;
    // The executable is the drop-in for ONE receiver, where bit-identity with the reference matters and throughput does not: by default
    // every filter keeps the reference's roundings (STRICT) and the RDS back end runs stage by stage, so audio, decoded bits and the
    // frame_thread lines on stderr are the reference's by construction, whatever the input.  (The batch library defaults to REFERENCE
    // numerics and the symbol-rate back end; there 2 of 1.2 M decoded bits differed from the reference's on noisy input)
    if (numerics == FMRX_NUMERICS_STRICT) cfg.paths = FMRX_PATH_AUDIO | FMRX_PATH_RDS | FMRX_PATH_RDS_STAGES;
    fmrx_batch *rx = nullptr;
    if (fmrx_batch_create(&cfg, &rx) != FMRX_OK) die("fmrx_batch_create");
    const int na = fmrx_batch_audio_per_block(rx);
    fmrx_ring *ring = nullptr;
    if (fmrx_ring_create(rx, kSlots, blocks, &ring) != FMRX_OK) die("fmrx_ring_create");
    fmrx_rds_app *app = nullptr;
    if (rds_info && mode == 0 && fmrx_rds_app_create(1, &app) != FMRX_OK) die("fmrx_rds_app_create");

    std::atomic<int> reader_rc{0};
    const size_t slot_bytes = (size_t)blocks * FMRX_BLOCK_BYTES;
    std::thread reader([&] {
        for (;;) {
            uint8_t *iq = nullptr;
            if (fmrx_ring_acquire(ring, -1, &iq) != FMRX_OK) { reader_rc = 2; break; }
            const size_t got = read_fully(iq, slot_bytes);
            const int nb = (int)(got / FMRX_BLOCK_BYTES);  // whole blocks only (Q9)
            if (nb > 0 && fmrx_ring_commit_blocks(ring, nb) != FMRX_OK) {
                std::cerr << "fm_radio: " << fmrx_last_error() << std::endl;
                reader_rc = 2;
                break;
            }
            if (got < slot_bytes) break;
        }
        fmrx_ring_close(ring);  // end of input (or a failed commit): the main thread drains what was committed and stops
    });

    long long block_id = 0;
    std::vector<int32_t> offset(1, 0);
    for (;;) {
        fmrx_outputs out{};
        const int st = fmrx_ring_next(ring, -1, &out);
        if (st == FMRX_ERR_EOF) break;
        if (st != FMRX_OK) die("fmrx_ring_next");
        const int nb = fmrx_ring_step_blocks(ring);
        if (fwrite(out.audio, sizeof(int16_t), (size_t)nb * 2 * na, stdout) != (size_t)nb * 2 * na) {  // :302; a closed pipe downstream ends the run
            std::cerr << "fm_radio: short write on stdout" << std::endl;
            _exit(2);
        }
        fflush(stdout);  // one write per block, like the reference: a player downstream gets its 64 ms as soon as they exist
        if (app) fmrx_rds_app_feed(app, out.rds_bits, out.rds_n_bits, nb, nullptr, 0, nullptr);
        if (mode == 0 && !quiet) {
            if (block_id == 0) fmrx_batch_rds_offsets(rx, offset.data());
            char text[16384];
            for (int b = 0; b < nb; ++b) {
                fmrx_rds_format_block((int)(block_id + b), offset[0], out.rds_events + (size_t)b * FMRX_MAX_EVENTS, out.rds_n_events[b], text, sizeof(text));
                std::cerr << text;
            }
        }
        block_id += nb;
        if (fmrx_ring_release(ring) != FMRX_OK) die("fmrx_ring_release");
    }
    fflush(stdout);
    reader.join();
    const int rc = reader_rc.load();
    fmrx_ring_destroy(ring);
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
