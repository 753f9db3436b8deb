// TEST INFRASTRUCTURE: host-only exercise of rife::Combiner (csrc/combiner.h) -- concurrent submitters, a fake batch
// function, every request answered exactly once with its own result, batches formed under contention, no batch above
// the cap, a lone caller served alone.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

#include "combiner.h"

struct Req {
    int in = 0, out = -1, served = 0;
};

int main() {
    rife::Combiner<Req> c;
    std::atomic<int> maxn{0}, calls{0}, inside{0}, overlap{0};
    auto fn = [&](Req** r, int n) {
        if (inside.fetch_add(1) != 0) overlap++;  // two batches at once would break the "one leader" rule
        calls++;
        int m = maxn.load();
        while (n > m && !maxn.compare_exchange_weak(m, n)) {}
        std::this_thread::sleep_for(std::chrono::microseconds(300));
        for (int i = 0; i < n; i++) { r[i]->out = r[i]->in * 2 + 1; r[i]->served++; }
        inside.fetch_sub(1);
        return n > 0 ? 0 : -1;
    };
    // lone caller
    Req one;
    one.in = 20;
    int st = c.submit(&one, 8, fn);
    if (st != 0 || one.out != 41 || c.batches() != 1 || c.requests() != 1) { printf("FAIL lone caller\n"); return 1; }
    const int T = 12, K = 40, CAP = 8;
    std::vector<std::vector<Req>> reqs(T, std::vector<Req>(K));
    std::vector<std::thread> th;
    std::atomic<int> bad{0};
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            for (int k = 0; k < K; k++) {
                Req& r = reqs[t][k];
                r.in = t * 1000 + k;
                if (c.submit(&r, CAP, fn) != 0 || r.out != r.in * 2 + 1 || r.served != 1) bad++;
            }
        });
    for (auto& x : th) x.join();
    const unsigned long long nb = c.batches() - 1, nr = c.requests() - 1;
    printf("threads %d x %d requests: %llu batches, largest %d, fn calls %d\n", T, K, nb, maxn.load(), calls.load() - 1);
    if (bad.load() || overlap.load() || nr != (unsigned long long)T * K || maxn.load() > CAP || nb >= nr || maxn.load() < 2) {
        printf("FAIL bad=%d overlap=%d requests=%llu max=%d\n", bad.load(), overlap.load(), nr, maxn.load());
        return 1;
    }
    // a throwing batch function: every waiter gets the error status, the queue keeps working afterwards
    {
        std::atomic<int> thrown{0};
        auto bad_fn = [&](Req** r, int n) -> int {
            std::this_thread::sleep_for(std::chrono::microseconds(200));
            if (r[0]->in < 0) { thrown++; throw 42; }
            for (int i = 0; i < n; i++) r[i]->out = 7;
            return 0;
        };
        std::vector<Req> rq(6);
        std::vector<int> st2(6, 1);
        std::vector<std::thread> t2;
        for (int i = 0; i < 6; i++) {
            rq[i].in = -1;
            t2.emplace_back([&, i] { st2[i] = c.submit(&rq[i], 8, bad_fn); });
        }
        for (auto& x : t2) x.join();
        for (int i = 0; i < 6; i++)
            if (st2[i] != rife::Combiner<Req>::kThrown) { printf("FAIL throw status %d\n", st2[i]); return 1; }
        Req ok;
        ok.in = 1;
        if (c.submit(&ok, 8, bad_fn) != 0 || ok.out != 7 || thrown.load() < 1) { printf("FAIL queue dead after throw\n"); return 1; }
    }
    printf("COMBINER OK\n");
    return 0;
}
