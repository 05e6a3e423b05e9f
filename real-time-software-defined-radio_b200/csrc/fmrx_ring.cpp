// Ingest / egress ring in front of a batch handle: the many-station counterpart of what
// /root/reference/src/fm_radio.cpp:86-138 does for one station with a five-slot ring, three std::queue's and condition
// variables.  Host code only: it owns page-locked slots and sequences fmrx_batch_submit / fmrx_batch_wait; every copy
// and kernel runs on the handle's CUDA streams.
//
// Differences from the reference's queues, on purpose: waits are `while` loops on the predicate (the reference's are
// `if`s: a spurious wake-up walks on), the producer never computes into a slot it does not own (the reference fills the
// ring slot before taking the lock, :83-86), and a full ring blocks the producer instead of overwriting (real
// back-pressure; with a timeout for callers that would rather drop).
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "fmrx.h"
#include "fmrx_internal.h"

struct fmrx_ring {
    struct Slot {
        uint8_t *iq = nullptr;
        int16_t *audio = nullptr;
        uint8_t *bits = nullptr;
        int32_t *nbits = nullptr, *nev = nullptr;
        fmrx_rds_event *ev = nullptr;
        long long ticket = -1;
        int blocks = 0;  // blocks committed in this slot (<= n_blocks: a short last step at end of input)
    };
    fmrx_batch *b = nullptr;
    int n_blocks = 0;
    std::vector<Slot> slots;
    std::mutex m;
    std::condition_variable freed, committed_cv;
    // step counters: acquired >= committed >= taken >= released; slot of step k is k % n_slots
    long long acquired = 0, committed = 0, taken = 0, released = 0;
    bool closed = false;
};

namespace {
template <class T>
int pinned(T *&p, size_t count) {
    void *v = nullptr;
    const int e = fmrx_pinned_alloc(&v, (count ? count : 1) * sizeof(T));
    p = static_cast<T *>(v);
    return e;
}

template <class Pred>
bool wait_for(std::condition_variable &cv, std::unique_lock<std::mutex> &lk, int timeout_ms, Pred pred) {
    if (timeout_ms < 0) { cv.wait(lk, pred); return true; }
    return cv.wait_for(lk, std::chrono::milliseconds(timeout_ms), pred);
}
}  // namespace

extern "C" {

int fmrx_ring_create(fmrx_batch *b, int n_slots, int n_blocks, fmrx_ring **out) {
    if (!b || !out) return fmrx::fail(FMRX_ERR_ARG, "fmrx_ring_create: null pointer");
    *out = nullptr;
    int S, NB, na, audio_on, rds_on;
    fmrx::batch_shape(b, &S, &NB, &na, &audio_on, &rds_on);
    if (n_slots < 2 || n_slots > 8 || n_blocks <= 0 || n_blocks > NB)  // 8 = the completion events the handle keeps (kTickets)
        return fmrx::fail(FMRX_ERR_ARG, "fmrx_ring_create: 2 <= n_slots <= 8, 1 <= n_blocks <= %d (the handle's max_blocks)", NB);
    fmrx_ring *r = new (std::nothrow) fmrx_ring();
    if (!r) return fmrx::fail(FMRX_ERR_ALLOC, "out of host memory");
    r->b = b; r->n_blocks = n_blocks;
    r->slots.resize(n_slots);
    const size_t sb = (size_t)S * n_blocks;
    for (auto &s : r->slots) {
        int e = pinned(s.iq, sb * FMRX_BLOCK_BYTES);
        if (!e && audio_on) e = pinned(s.audio, sb * 2 * na);
        if (!e && rds_on) e = pinned(s.bits, sb * FMRX_MAX_BITS);
        if (!e && rds_on) e = pinned(s.nbits, sb);
        if (!e && rds_on) e = pinned(s.ev, sb * FMRX_MAX_EVENTS);
        if (!e && rds_on) e = pinned(s.nev, sb);
        if (e) { fmrx_ring_destroy(r); return e; }
    }
    *out = r;
    return FMRX_OK;
}

void fmrx_ring_destroy(fmrx_ring *r) {
    if (!r) return;
    if (r->committed > r->released) fmrx_batch_sync(r->b);  // nothing may still be copying into a slot that is about to be freed
    for (auto &s : r->slots) {
        fmrx_pinned_free(s.iq); fmrx_pinned_free(s.audio); fmrx_pinned_free(s.bits);
        fmrx_pinned_free(s.nbits); fmrx_pinned_free(s.ev); fmrx_pinned_free(s.nev);
    }
    delete r;
}

int fmrx_ring_acquire(fmrx_ring *r, int timeout_ms, uint8_t **iq) {
    if (!r || !iq) return fmrx::fail(FMRX_ERR_ARG, "fmrx_ring_acquire: null pointer");
    std::unique_lock<std::mutex> lk(r->m);
    if (r->closed) return fmrx::fail(FMRX_ERR_STATE, "fmrx_ring_acquire: the ring is closed");
    if (r->acquired != r->committed) return fmrx::fail(FMRX_ERR_STATE, "fmrx_ring_acquire: the previous slot was not committed");
    const long long n = (long long)r->slots.size();
    if (!wait_for(r->freed, lk, timeout_ms, [&] { return r->acquired - r->released < n; }))
        return fmrx::fail(FMRX_ERR_TIMEOUT, "fmrx_ring_acquire: all %lld slots in flight for %d ms", n, timeout_ms);
    *iq = r->slots[r->acquired % n].iq;
    r->acquired += 1;
    return FMRX_OK;
}

int fmrx_ring_commit(fmrx_ring *r) { return r ? fmrx_ring_commit_blocks(r, r->n_blocks) : fmrx::fail(FMRX_ERR_ARG, "null handle"); }

int fmrx_ring_commit_blocks(fmrx_ring *r, int n_blocks) {
    if (!r) return fmrx::fail(FMRX_ERR_ARG, "null handle");
    if (n_blocks <= 0 || n_blocks > r->n_blocks) return fmrx::fail(FMRX_ERR_ARG, "fmrx_ring_commit_blocks: 1 <= n_blocks <= %d", r->n_blocks);
    std::unique_lock<std::mutex> lk(r->m);
    if (r->acquired != r->committed + 1) return fmrx::fail(FMRX_ERR_STATE, "fmrx_ring_commit: no slot acquired");
    fmrx_ring::Slot &s = r->slots[r->committed % (long long)r->slots.size()];
    fmrx_outputs o{};
    o.audio = s.audio; o.rds_bits = s.bits; o.rds_n_bits = s.nbits; o.rds_events = s.ev; o.rds_n_events = s.nev;
    // the submit itself only enqueues (copies and kernels are asynchronous); holding the lock keeps handle calls serial
    if (int e = fmrx_batch_submit(r->b, s.iq, n_blocks, &o, &s.ticket)) { r->acquired -= 1; return e; }
    s.blocks = n_blocks;
    r->committed += 1;
    lk.unlock();
    r->committed_cv.notify_all();
    return FMRX_OK;
}

int fmrx_ring_close(fmrx_ring *r) {
    if (!r) return fmrx::fail(FMRX_ERR_ARG, "null handle");
    {
        std::lock_guard<std::mutex> lk(r->m);
        if (r->acquired != r->committed) r->acquired = r->committed;  // an acquired but never committed slot is dropped
        r->closed = true;
    }
    r->committed_cv.notify_all();
    return FMRX_OK;
}

int fmrx_ring_next(fmrx_ring *r, int timeout_ms, fmrx_outputs *out) {
    if (!r || !out) return fmrx::fail(FMRX_ERR_ARG, "fmrx_ring_next: null pointer");
    long long ticket;
    fmrx_ring::Slot *s;
    {
        std::unique_lock<std::mutex> lk(r->m);
        if (r->taken != r->released) return fmrx::fail(FMRX_ERR_STATE, "fmrx_ring_next: the previous step was not released");
        if (!wait_for(r->committed_cv, lk, timeout_ms, [&] { return r->committed > r->taken || r->closed; }))
            return fmrx::fail(FMRX_ERR_TIMEOUT, "fmrx_ring_next: nothing committed for %d ms", timeout_ms);
        if (r->committed == r->taken) return fmrx::fail(FMRX_ERR_EOF, "fmrx_ring_next: closed and drained");
        s = &r->slots[r->taken % (long long)r->slots.size()];
        ticket = s->ticket;
        r->taken += 1;
    }
    // outside the lock: the producer keeps committing while this thread sleeps on the step's copy-out event
    if (int e = fmrx_batch_wait(r->b, ticket)) {
        std::lock_guard<std::mutex> lk(r->m);
        r->taken -= 1;
        return e;
    }
    std::memset(out, 0, sizeof(*out));
    out->audio = s->audio; out->rds_bits = s->bits; out->rds_n_bits = s->nbits; out->rds_events = s->ev; out->rds_n_events = s->nev;
    return FMRX_OK;
}

int fmrx_ring_release(fmrx_ring *r) {
    if (!r) return fmrx::fail(FMRX_ERR_ARG, "null handle");
    {
        std::lock_guard<std::mutex> lk(r->m);
        if (r->taken != r->released + 1) return fmrx::fail(FMRX_ERR_STATE, "fmrx_ring_release: no step taken");
        r->released += 1;
    }
    r->freed.notify_all();
    return FMRX_OK;
}

int fmrx_ring_step_blocks(fmrx_ring *r) {
    if (!r) return 0;
    std::lock_guard<std::mutex> lk(r->m);
    if (r->taken != r->released + 1) return 0;
    return r->slots[r->released % (long long)r->slots.size()].blocks;
}

int fmrx_ring_in_flight(fmrx_ring *r) {
    if (!r) return 0;
    std::lock_guard<std::mutex> lk(r->m);
    return (int)(r->committed - r->released);
}

}  // extern "C"
