// Host-side helpers shared by the CUDA engine (af_engine.cu) and the CPU-only
// debugging twin (tests/host_twin): scenario validation, workspace layout and the
// serial histogram-quantile used to cross-check the device percentile kernel.
#pragma once
#include <math.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "af_core.cuh"

namespace afh {

constexpr int32_t DEFAULT_EVENT_CAPACITY = 2048;
constexpr int32_t DEFAULT_REQUEST_CAPACITY = 16384;
constexpr int32_t DEFAULT_EV_SMEM = 64;
constexpr int32_t DEFAULT_RQ_SMEM = 64;

inline bool validate(const AfScenario& s, std::string& err) {
    char buf[256];
#define AF_CHECK(cond, ...) do { if (!(cond)) { snprintf(buf, sizeof buf, __VA_ARGS__); err = buf; return false; } } while (0)
    AF_CHECK(s.n_edges > 0 && s.n_edges <= (int)afc::AUX_MASK, "n_edges=%d out of range [1,%u]", s.n_edges, afc::AUX_MASK);
    AF_CHECK(s.n_servers > 0 && s.n_servers <= 120, "n_servers=%d out of range [1,120]", s.n_servers);
    AF_CHECK(s.n_endpoints > 0 && s.n_endpoints < 4096, "n_endpoints=%d out of range", s.n_endpoints);
    AF_CHECK(s.n_steps >= 0 && s.n_steps < (1 << 20), "n_steps=%d out of range", s.n_steps);
    AF_CHECK(s.horizon_s >= 1, "horizon_s=%d must be >= 1", s.horizon_s);
    AF_CHECK(s.window_s >= 1, "window_s=%d must be >= 1", s.window_s);
    AF_CHECK(s.sample_period > 0.0, "sample_period must be > 0");
    AF_CHECK(s.users_dist == AF_DIST_POISSON || s.users_dist == AF_DIST_NORMAL, "users_dist=%d unsupported", s.users_dist);
    AF_CHECK(s.gen_edge >= 0 && s.gen_edge < s.n_edges, "gen_edge out of range");
    AF_CHECK(s.client_edge >= 0 && s.client_edge < s.n_edges, "client_edge out of range");
    AF_CHECK(s.edges && s.servers && s.endpoints, "null topology arrays");
    AF_CHECK(s.n_steps == 0 || s.steps, "null steps");
    bool has_lb_target = false;
    for (int i = 0; i < s.n_edges; ++i) {
        const AfEdge& e = s.edges[i];
        AF_CHECK(e.dist >= 0 && e.dist <= 4, "edge %d: dist=%d", i, e.dist);
        AF_CHECK(e.target_kind >= 0 && e.target_kind <= 2, "edge %d: target_kind=%d", i, e.target_kind);
        if (e.target_kind == AF_TARGET_SERVER)
            AF_CHECK(e.target_index >= 0 && e.target_index < s.n_servers, "edge %d: target server %d", i, e.target_index);
        if (e.target_kind == AF_TARGET_LB) has_lb_target = true;
        AF_CHECK(e.dropout >= 0.0 && e.dropout <= 1.0, "edge %d: dropout", i);
    }
    if (has_lb_target) {
        AF_CHECK(s.n_lb_edges > 0 && s.lb_edges, "an edge targets the LB but n_lb_edges == 0");
        AF_CHECK(s.lb_algo == AF_LB_ROUND_ROBIN || s.lb_algo == AF_LB_LEAST_CONNECTIONS, "lb_algo=%d", s.lb_algo);
    }
    for (int i = 0; i < s.n_lb_edges; ++i)
        AF_CHECK(s.lb_edges[i] >= 0 && s.lb_edges[i] < s.n_edges, "lb_edges[%d] out of range", i);
    for (int i = 0; i < s.n_servers; ++i) {
        const AfServer& v = s.servers[i];
        AF_CHECK(v.cpu_cores >= 1 && v.ram_mb >= 1, "server %d: resources", i);
        AF_CHECK(v.out_edge >= 0 && v.out_edge < s.n_edges, "server %d: out_edge", i);
        AF_CHECK(v.n_endpoints >= 1 && v.endpoint_begin >= 0 && v.endpoint_begin + v.n_endpoints <= s.n_endpoints,
                 "server %d: endpoint range", i);
    }
    for (int i = 0; i < s.n_endpoints; ++i) {
        const AfEndpoint& p = s.endpoints[i];
        AF_CHECK(p.n_steps >= 0 && p.n_steps <= 255 && p.step_begin >= 0 && p.step_begin + p.n_steps <= s.n_steps,
                 "endpoint %d: step range", i);
        AF_CHECK(p.total_ram >= 0, "endpoint %d: total_ram", i);
    }
    for (int i = 0; i < s.n_steps; ++i)
        AF_CHECK((s.steps[i].kind == AF_STEP_CPU || s.steps[i].kind == AF_STEP_IO) && s.steps[i].duration >= 0.0,
                 "step %d invalid", i);
    for (int i = 0; i < s.n_spike_marks; ++i) {
        AF_CHECK(s.spike_marks[i].edge >= 0 && s.spike_marks[i].edge < s.n_edges, "spike mark %d: edge", i);
        AF_CHECK(i == 0 || s.spike_marks[i].fire_time >= s.spike_marks[i - 1].fire_time, "spike marks unsorted");
    }
    for (int i = 0; i < s.n_outage_marks; ++i) {
        AF_CHECK(s.outage_marks[i].lb_edge < s.n_edges, "outage mark %d: edge", i);
        AF_CHECK(i == 0 || s.outage_marks[i].fire_time >= s.outage_marks[i - 1].fire_time, "outage marks unsorted");
    }
#undef AF_CHECK
    return true;
}

inline int32_t trace_tick_capacity(const AfScenario& s) {
    double n = (double)s.horizon_s / s.sample_period;
    if (n > 50.0e6) n = 50.0e6;
    return (int32_t)n + 2;
}


inline void make_layout(const AfScenario& s, const AfOptions& o, int32_t n_sweep_cols, afc::Layout& L) {
    L.n_edges = s.n_edges; L.n_servers = s.n_servers; L.n_endpoints = s.n_endpoints; L.n_steps = s.n_steps;
    L.n_lb_edges = s.n_lb_edges; L.lb_algo = s.lb_algo; L.gen_edge = s.gen_edge; L.client_edge = s.client_edge;
    L.n_spike = s.n_spike_marks; L.n_outage = s.n_outage_marks;
    L.users_dist = s.users_dist; L.window_s = s.window_s; L.horizon_s = s.horizon_s; L.metrics_mask = s.metrics_mask;
    L.users_mean = s.users_mean; L.users_sigma = s.users_sigma; L.rate_per_user = s.rate_per_user;
    L.sample_period = s.sample_period;
    int32_t ev = o.event_capacity > 0 ? o.event_capacity : DEFAULT_EVENT_CAPACITY;
    int32_t rq = o.request_capacity > 0 ? o.request_capacity : DEFAULT_REQUEST_CAPACITY;
    if (rq > (int32_t)afc::SLOT_MASK) rq = (int32_t)afc::SLOT_MASK;
    L.ev_total = ev; L.ev_smem = ev < DEFAULT_EV_SMEM ? ev : DEFAULT_EV_SMEM;
    L.rq_total = rq; L.rq_smem = rq < DEFAULT_RQ_SMEM ? rq : DEFAULT_RQ_SMEM;
    L.n_series = 3 * s.n_servers + s.n_edges;
    L.n_sweep_cols = n_sweep_cols;
    L.collect_hist = o.collect_histogram; L.collect_thr = o.collect_throughput;
    L.trace_replicas = o.trace_replicas; L.trace_clock_cap = o.trace_clock_capacity;
    L.trace_tick_cap = trace_tick_capacity(s);
    afc::layout_finalize(L);
}

// lower edge (seconds) of histogram bin `idx`
inline double hist_bin_edge(int32_t idx) {
    uint64_t bits = ((uint64_t)(idx + ((1023 + AF_HIST_MIN_EXP) << AF_HIST_SUB_BITS))) << (52 - AF_HIST_SUB_BITS);
    double x; memcpy(&x, &bits, 8); return x;
}

// value of the order statistic `rank` (0-based) read off the histogram
inline double hist_order_stat(const uint32_t* hist, uint64_t rank) {
    uint64_t cum = 0;
    for (int32_t b = 0; b < AF_HIST_BINS; ++b) {
        uint64_t c = hist[b];
        if (c && rank < cum + c) {
            double lo = hist_bin_edge(b), hi = hist_bin_edge(b + 1);
            double frac = ((double)(rank - cum) + 0.5) / (double)c;
            return lo + frac * (hi - lo);
        }
        cum += c;
    }
    return NAN;
}

// numpy.percentile(..., method="linear") evaluated on the histogram
inline double hist_percentile(const uint32_t* hist, uint64_t n, double q) {
    if (n == 0) return NAN;
    double pos = q / 100.0 * (double)(n - 1);
    uint64_t lo = (uint64_t)floor(pos);
    double frac = pos - (double)lo;
    double a = hist_order_stat(hist, lo);
    if (frac == 0.0 || lo + 1 >= n) return a;
    double b = hist_order_stat(hist, lo + 1);
    return a + frac * (b - a);
}

}  // namespace afh
