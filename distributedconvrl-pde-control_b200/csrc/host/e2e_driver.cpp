// Native host driver for the end-to-end measurement: what a compiled-language host (the reference is Julia: tasks
// on real threads, no GIL) does with the C ABI.  One std::thread per environment shard runs the drop-in call
// sequence with HOST buffers, `steps` times:
//     pdeb200_policy_act            (policy(env): actor forward on the device)
//     pdeb200_get(ARR_ACTION_IN)    (the action comes back to the host, src/PDEagent.jl:198)
//     pdeb200_step_host             (env(action): H2D action; kernels; D2H reward, state, done)
// Only include/pdeb200.h is used; this file is built into libpdeb200_host.so next to libpdeb200.so.
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "../../../include/pdeb200.h"

// fused = 1: the same two reference-side calls as ONE C call with one synchronisation and one packed result copy
// (pdeb200_act_step_host); h_packed[k] then receives [reward | done | state] or its [reward | done] prefix (pdeb200_result_layout).
// h_noise != NULL: exploration noise drawn on the host (the reference's randn(policy.rng, ...), PDEagent.jl:201) goes in with
// every step; h_act == NULL: the action stays on the device (device policy + device trajectory).
extern "C" int32_t pdeb200_host_drive2(int32_t n_shards, pdeb200_ctx** ctxs, int32_t steps, void** h_act, void** h_packed, double act_limit,
                                       double* seconds_out, const double** h_noise, double act_noise) {
    if (n_shards < 1 || !ctxs || steps < 0 || !h_packed || !seconds_out) return PDEB200_EINVAL;
    std::atomic<int> ready{0}, failed{0};
    std::atomic<bool> go{false};
    std::vector<std::thread> th;
    auto work = [&](int k) {
        ready.fetch_add(1);
        while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
        // prefetch (act_noise < 0 selects it): step i+1's noise is handed over before the call for step i, so that its upload
        // overlaps step i's kernels (pdeb200_noise_prefetch); the noise of every step still crosses PCIe inside the timed region
        const bool prefetch = h_noise && act_noise < 0.0;
        const double sigma = act_noise < 0.0 ? -act_noise : act_noise;
        if (prefetch && steps > 0 && pdeb200_noise_prefetch(ctxs[k], h_noise[k])) { failed.store(PDEB200_ESTATE); return; }
        for (int i = 0; i < steps; ++i) {
            if (prefetch && i + 1 < steps && pdeb200_noise_prefetch(ctxs[k], h_noise[k])) { failed.store(PDEB200_ESTATE); return; }
            const int32_t rc = pdeb200_act_step_host(ctxs[k], (h_noise && !prefetch) ? h_noise[k] : nullptr, h_noise ? sigma : 0.0, act_limit,
                                                     h_act ? h_act[k] : nullptr, nullptr, h_packed[k], nullptr, nullptr, nullptr);
            if (rc) { failed.store(rc); return; }
        }
    };
    for (int k = 0; k < n_shards; ++k) th.emplace_back(work, k);
    while (ready.load() < n_shards) std::this_thread::yield();
    const auto t0 = std::chrono::steady_clock::now();
    go.store(true, std::memory_order_release);
    for (auto& t : th) t.join();
    *seconds_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return failed.load();
}

extern "C" int32_t pdeb200_host_drive(int32_t n_shards, pdeb200_ctx** ctxs, int32_t steps, void** h_act, const size_t* act_bytes,
                                      void** h_reward, void** h_state, uint8_t** h_done, double act_limit, double* seconds_out) {
    if (n_shards < 1 || !ctxs || steps < 0 || !h_act || !act_bytes || !seconds_out) return PDEB200_EINVAL;
    std::atomic<int> ready{0}, failed{0};
    std::atomic<bool> go{false};
    std::vector<std::thread> th;
    auto work = [&](int k) {
        ready.fetch_add(1);
        while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
        for (int i = 0; i < steps; ++i) {
            int32_t rc = pdeb200_policy_act(ctxs[k], nullptr, 0.0, act_limit);
            if (!rc) rc = pdeb200_get(ctxs[k], PDEB200_ARR_ACTION_IN, h_act[k], act_bytes[k]);
            if (!rc) rc = pdeb200_step_host(ctxs[k], h_act[k], nullptr, h_reward ? h_reward[k] : nullptr, h_state ? h_state[k] : nullptr,
                                            h_done ? h_done[k] : nullptr);
            if (rc) { failed.store(rc); return; }
        }
    };
    for (int k = 0; k < n_shards; ++k) th.emplace_back(work, k);
    while (ready.load() < n_shards) std::this_thread::yield();
    const auto t0 = std::chrono::steady_clock::now();
    go.store(true, std::memory_order_release);
    for (auto& t : th) t.join();
    *seconds_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return failed.load();
}
