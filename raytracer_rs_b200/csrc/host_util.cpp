// host_util.cpp — stand-ins for raytracer_lib/src/stats.rs (`Stats`) and the `timing` crate (`BenchMark`),
// so the native binary's loop (raytracer/src/main.rs:181,213,216) can keep its reporting calls.
#include <algorithm>
#include <charconv>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rt_b200.h"

namespace {

using Clock = std::chrono::steady_clock;

// Rust's `{}` for f32: shortest decimal that round-trips, never scientific notation
std::string rust_f32(float v) {
    if (v != v) return "NaN";
    if (v == __builtin_inff()) return "inf";
    if (v == -__builtin_inff()) return "-inf";
    char buf[128];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

int emit(const std::string& s, char* out, size_t out_len) {
    if (!out || out_len == 0) return RT_ERR_INVALID;
    std::snprintf(out, out_len, "%s", s.c_str());
    return RT_OK;
}

}  // namespace

struct rt_stats {  // stats.rs:1-6
    Clock::time_point last_iteration = Clock::now();
    float fps_sum = 0.f, primrays_per_sec_sum = 0.f;
    uint32_t num_measurements = 0;
};

struct rt_benchmark {  // timing/src/lib.rs:6-9,61-66 ; insertion order kept only to make ties deterministic
    struct Timing {
        std::string name;
        Clock::time_point start;
        std::chrono::nanoseconds total{0};
        size_t samples = 0;
    };
    std::vector<Timing> timings;
    Timing* find(const char* name) {
        for (auto& t : timings)
            if (t.name == name) return &t;
        return nullptr;
    }
};

extern "C" {

rt_stats* rt_stats_new(void) { return new rt_stats(); }
void rt_stats_free(rt_stats* s) { delete s; }

int rt_stats_stats(rt_stats* s, uint32_t num_primary_rays, char* out, size_t out_len) {  // stats.rs:21-32
    if (!s) return RT_ERR_INVALID;
    const auto now = Clock::now();
    const float secs = std::chrono::duration<float>(now - s->last_iteration).count();
    s->last_iteration = now;
    const float fps = 1.0f / secs;
    s->fps_sum += fps;
    const float prps = (float)num_primary_rays / secs;
    s->primrays_per_sec_sum += prps;
    s->num_measurements += 1;
    // `primrays_per_sec as u32` saturates
    const uint32_t as_u32 = prps != prps ? 0u : prps >= 4294967296.0f ? 0xFFFFFFFFu : prps <= 0.f ? 0u : (uint32_t)prps;
    return emit("fps: " + rust_f32(fps) + "  primary rays/s: " + std::to_string(as_u32), out, out_len);
}

int rt_stats_mean_stats(const rt_stats* s, char* out, size_t out_len) {  // stats.rs:34-40
    if (!s) return RT_ERR_INVALID;
    const float n = (float)s->num_measurements;
    return emit("mean fps: " + rust_f32(s->fps_sum / n) + "  mean primary rays/s: " + rust_f32(s->primrays_per_sec_sum / n), out, out_len);
}

rt_benchmark* rt_benchmark_new(void) { return new rt_benchmark(); }
void rt_benchmark_free(rt_benchmark* b) { delete b; }

int rt_benchmark_start(rt_benchmark* b, const char* name) {  // timing/src/lib.rs:18-24
    if (!b || !name) return RT_ERR_INVALID;
    const auto now = Clock::now();
    if (auto* t = b->find(name)) {
        t->start = now;
    } else {
        rt_benchmark::Timing nt;
        nt.name = name;
        nt.start = now;
        b->timings.push_back(nt);
    }
    return RT_OK;
}

int rt_benchmark_stop(rt_benchmark* b, const char* name) {  // timing/src/lib.rs:26-35
    if (!b || !name) return RT_ERR_INVALID;
    auto* t = b->find(name);
    if (!t) return RT_ERR_INVALID;  // the reference panics: "unexpected name in stop()"
    t->total += std::chrono::duration_cast<std::chrono::nanoseconds>(Clock::now() - t->start);
    t->samples += 1;
    return RT_OK;
}

int rt_benchmark_report(const rt_benchmark* b, char* out, size_t out_len) {  // lib.rs:45-58,95-109
    if (!b) return RT_ERR_INVALID;
    std::vector<const rt_benchmark::Timing*> order;
    for (auto& t : b->timings) order.push_back(&t);
    std::stable_sort(order.begin(), order.end(), [](auto* x, auto* y) { return x->total > y->total; });
    std::string s;
    for (auto* t : order) {
        if (t->samples == 0) continue;  // the reference divides by zero samples and panics; skip instead
        const auto total_us = std::chrono::duration_cast<std::chrono::microseconds>(t->total).count();
        const auto mean_ns = t->total.count() / (long long)t->samples;
        const auto mean_us = mean_ns / 1000;
        s += t->name + " total: " + rust_f32((float)total_us / 1000.0f) + "ms, mean: " + rust_f32((float)mean_us / 1000.0f) +
             "ms, samples: " + std::to_string(t->samples) + "\n";
    }
    return emit(s, out, out_len);
}

}  // extern "C"
