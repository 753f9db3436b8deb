// combiner.h -- turns concurrent single-request calls into batches (flat combining).
//
// The reference's CLI calls RIFE::process from several `proc` threads on one object (src/main.cpp:346-366, `-j l:p:s`).
// Serialising those calls runs one frame pair at a time; the fused path is at its best with a lock-step batch of
// pairs.  With a Combiner the first caller to arrive becomes the leader: it takes its own request plus whatever other
// threads have queued meanwhile (up to max_batch), executes them as ONE batch, publishes the status to their owners
// and hands leadership over.  Nobody waits for a batch to fill: a lone caller runs alone, at once.
// Host-only, no CUDA: tests/emu/test_combiner.cpp exercises it with a fake batch function.
#pragma once
#include <condition_variable>
#include <deque>
#include <mutex>

namespace rife {

template <class Req>
class Combiner {
public:
    // fn(Req** reqs, int n) -> status for all n requests.  Called without the combiner's lock held, by one thread at a time.
    template <class Fn>
    int submit(Req* req, int max_batch, Fn&& fn) {
        Slot me;
        me.req = req;
        std::unique_lock<std::mutex> lk(m_);
        pending_.push_back(&me);
        for (;;) {
            if (me.done) return me.status;
            if (!leader_) break;  // nobody is serving the queue: this thread does
            cv_.wait(lk);
        }
        leader_ = true;
        // one batch: the oldest requests first (always includes requests older than ours, and ours unless the queue
        // ahead of us is longer than max_batch -- then we lead again after handing the results out)
        while (!me.done) {
            Slot* take[kMax];
            Req* reqs[kMax];
            int n = 0;
            const int cap = max_batch < 1 ? 1 : (max_batch > kMax ? kMax : max_batch);
            while (n < cap && !pending_.empty()) {
                take[n] = pending_.front();
                reqs[n] = take[n]->req;
                pending_.pop_front();
                n++;
            }
            lk.unlock();
            int status;
            try {
                status = fn(reqs, n);
            } catch (...) {
                status = kThrown;  // the batch function must not take the queue down with it: everyone gets an error, leadership moves on
            }
            lk.lock();
            for (int i = 0; i < n; i++) { take[i]->status = status; take[i]->done = true; }
            batches_++;
            requests_ += n;
            cv_.notify_all();
        }
        leader_ = false;
        cv_.notify_all();  // a waiter whose request is still queued takes over
        return me.status;
    }
    // statistics (diagnostics / tests)
    unsigned long long batches() { std::lock_guard<std::mutex> lk(m_); return batches_; }
    unsigned long long requests() { std::lock_guard<std::mutex> lk(m_); return requests_; }

    static constexpr int kMax = 32;
    static constexpr int kThrown = -1000;  // status handed to every request of a batch whose function threw

private:
    struct Slot {
        Req* req = nullptr;
        bool done = false;
        int status = 0;
    };
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<Slot*> pending_;
    bool leader_ = false;
    unsigned long long batches_ = 0, requests_ = 0;
};

}  // namespace rife
