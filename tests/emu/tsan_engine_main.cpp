// tsan_engine_main.cpp -- TEST INFRASTRUCTURE.  ThreadSanitizer over the engine's host-side threading: the engine's unmodified host
// code (capi / engine / exec / model), the host-memory runtime and kernel launches that do nothing (EMU_SKIP_KERNELS), driven the way
// the reference's CLI drives one RIFE object: several "proc" threads call rife_b200_process on ONE handle (src/main.cpp:346-366)
// -- pageable frames, so every call also goes through the per-thread staging slots and the request combiner -- while another
// thread reads and writes options and error text.  Results are not looked at (no kernels ran); a data race is a TSan report and a
// non-zero exit.  Built and run by tests/test_engine_emu_cpu.py.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "rife_b200.h"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    rife_b200_t* h = nullptr;
    if (rife_b200_create(&h, 0, 0, 0, 0, 1, 0, 1) != 0) return 3;
    if (rife_b200_load(h, argv[1]) != 0) { fprintf(stderr, "load: %s\n", rife_b200_last_error(h)); return 4; }
    const int w = 64, hh = 32, nthreads = 6, calls = 6;
    const size_t nb = (size_t)w * hh * 3;
    std::atomic<int> failures{0};
    std::atomic<bool> stop{false};
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([&, t]() {
            std::vector<unsigned char> a(nb, (unsigned char)(16 * t)), b(nb, (unsigned char)(16 * t + 7)), o(nb);
            for (int i = 0; i < calls; i++) {
                const float ts = i == 2 ? 0.f : 0.5f;  // one copy-through call per thread
                if (rife_b200_process(h, a.data(), b.data(), w, hh, ts, o.data()) != 0) failures++;
                (void)rife_b200_last_error(h);
            }
            // a refused call and its message, concurrently with the others' good calls
            if (rife_b200_process(h, nullptr, b.data(), w, hh, 0.5f, o.data()) == 0) failures++;
            if (!rife_b200_last_error(h) || !*rife_b200_last_error(h)) failures++;
        });
    std::thread opt([&]() {
        int v = 0, k = 0;
        while (!stop.load()) {
            rife_b200_get_option(h, "fast_active", &v);
            rife_b200_get_option(h, "combined_batches", &v);
            rife_b200_set_option(h, "frame_cache", (k++ >> 3) & 1);
            rife_b200_set_option(h, "combine", 1);
            rife_b200_forget_frames(h);
            std::this_thread::yield();
        }
    });
    for (auto& t : th) t.join();
    stop.store(true);
    opt.join();
    int batches = 0, requests = 0;
    rife_b200_get_option(h, "combined_batches", &batches);
    rife_b200_get_option(h, "combined_requests", &requests);
    printf("TSAN-ENGINE failures=%d combined_batches=%d combined_requests=%d\n", failures.load(), batches, requests);
    rife_b200_destroy(h);
    return failures.load() ? 5 : 0;
}
