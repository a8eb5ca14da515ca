// af_core.cuh -- the per-replica next-event engine (one replica per warp).
//
// This is the B200-native replacement of the reference's hot loop: SimPy's
// Environment.step() (external simpy 4.1.1, called from reference
// src/asyncflow/runtime/simulation_runner.py:369) popping (time, prio, eid)
// tuples and resuming the AsyncFlow actor generators.  Here every actor is a
// branch of one state machine and only events that carry simulated delay are
// queued (6-7 per request instead of SimPy's 25-30); each
// zero-delay SimPy cascade (Initialize / Store put+get / Container put+get /
// process exit) is executed inline by the handler of the timed event that
// starts it, in the order SimPy would have produced:
//
//   handler            reference being replaced
//   -----------------  ----------------------------------------------------------
//   gen_next_gap       samplers/poisson_poisson.py:52-82, gaussian_poisson.py:64-94
//   on_arrival         runtime/actors/rqs_generator.py:97-119
//   edge_send          runtime/actors/edge.py:73-107 (dropout, latency, spike)
//   on_deliver         edge.py:110-116 + client.py:43-71 + load_balancer.py:60-72
//                      + routing/lb_algorithms.py:10-36 + server.py:303-313
//   server_arrive      runtime/actors/server.py:88-149 (endpoint pick, RAM first)
//   run_steps          server.py:197-255 (lazy CPU lock, IO queue)
//   finish_request     server.py:257-276 (release core, RAM, forward)
//   on_spike/on_outage runtime/events/injection.py:167-226
//   take_samples       metrics/collector.py:50-66
//   complete           client.py:62-69 + metrics/analyzer.py:83-125
//
// Ordering rule: events pop by (time, seq); seq is the per-replica push counter,
// the analogue of SimPy's eid.  Ticks of the sampled-metric collector are not
// queued: state is piecewise constant between events, so the samples that fall
// before an event are emitted lazily just before it (they carry a seq too, so
// ties are ordered as SimPy would).
//
// Execution model: ALL 32 lanes of the warp run this code redundantly on the
// same replica (warp-uniform control flow, identical values in every lane).
// Lanes differ only inside the few helpers that say so: the pending-event pool
// is an unsorted, lane-strided array and pop-min is a warp arg-min
// (redux.sync on the time words, then seq); per-entity loops (sampling, final
// write-back, parameter load) are lane-strided.  The same source compiles for
// the host with a warp of ONE lane (tests/host_twin -- a debugging twin used by
// the CPU-only tests; it is NOT reachable from the product API).
#pragma once
#include "af_rng.cuh"
#include "../../include/asyncflow_b200.h"

#if defined(__CUDA_ARCH__)
#define AF_DEVICE_CODE 1
#else
#define AF_DEVICE_CODE 0
#endif

namespace afc {

constexpr uint32_t NIL = 0xFFFFFFFFu;
constexpr uint64_t INF_BITS = 0x7FF0000000000000ull;

// ---- event payload: kind[29:32) | aux[20:29) | slot[0:20) -------------------
enum : uint32_t { K_ARRIVAL = 0, K_DELIVER = 1, K_STEP_END = 2, K_SPIKE = 3, K_OUTAGE = 4 };
constexpr uint32_t SLOT_BITS = 20, AUX_BITS = 9;
constexpr uint32_t SLOT_MASK = (1u << SLOT_BITS) - 1, AUX_MASK = (1u << AUX_BITS) - 1;
AF_HD uint32_t mk_payload(uint32_t kind, uint32_t aux, uint32_t slot) {
    return (kind << 29) | (aux << SLOT_BITS) | slot;
}

// ---- request record pack: hops[0:8) step[8:16) ep[16:28) core[28] io[29] ----
constexpr uint32_t PK_CORE = 1u << 28, PK_IO = 1u << 29;
AF_HD uint32_t pk_hops(uint32_t p) { return p & 0xFFu; }
AF_HD uint32_t pk_step(uint32_t p) { return (p >> 8) & 0xFFu; }
AF_HD uint32_t pk_ep(uint32_t p) { return (p >> 16) & 0xFFFu; }

struct ReqRec { double t0; uint32_t rid; uint32_t pack; };          // 16 B

// ---- per-warp private tables (shared memory on the device) -------------------
struct EdgeS {            // 48 B
    double mean, sigma, dropout, spike;
    uint32_t meta;        // dist[0:3) | target_kind[3:5) | target_index[5:)
    uint32_t conn, sent, dropped;
};
struct ServerS {          // 48 B
    int32_t cpu_free, ram_free, ready_q, io_q, ram_in_use;
    uint32_t ramq_head, ramq_tail, cpuq_head, cpuq_tail;
    uint32_t out_edge, ep_begin, n_ep;
};
struct EndpointS { uint32_t step_begin, n_steps, total_ram, pad; }; // 16 B
struct StepS { double dur; uint32_t kind, pad; };                   // 16 B
struct SpikeS { double fire, delta; uint32_t edge, pad; };           // 24 B
struct OutageS { double fire; int32_t lb_edge, down; };              // 16 B

// Everything the kernel needs to know about sizes; built on the host.
struct Layout {
    int32_t n_edges, n_servers, n_endpoints, n_steps, n_lb_edges, lb_algo;
    int32_t gen_edge, client_edge, n_spike, n_outage;
    int32_t users_dist, window_s, horizon_s;
    uint32_t metrics_mask;
    double users_mean, users_sigma, rate_per_user, sample_period;
    int32_t ev_smem, ev_total;     // pending-event pool: slots in smem / in total
    int32_t rq_smem, rq_total;     // request table: slots in smem / in total
    int32_t n_series;              // 3*n_servers + n_edges
    int32_t n_sweep_cols;
    int32_t collect_hist, collect_thr;
    int32_t trace_replicas, trace_clock_cap, trace_tick_cap;
    // byte offsets inside the per-warp workspace
    int32_t off_ev_time, off_ev_key, off_rq_rec, off_rq_next, off_edge, off_server,
            off_endpoint, off_step, off_lb, off_spike, off_outage, off_samp_sum, off_samp_max;
    int32_t warp_bytes;
};

AF_HD int32_t align_up(int32_t x, int32_t a) { return (x + a - 1) / a * a; }

inline void layout_finalize(Layout& L) {
    int32_t o = 0;
    L.off_ev_time = o;  o += 8 * L.ev_smem;
    L.off_ev_key = o;   o += 8 * L.ev_smem;
    L.off_rq_rec = o;   o += 16 * L.rq_smem;
    L.off_edge = o;     o += 48 * L.n_edges;
    L.off_server = o;   o += 48 * L.n_servers;
    L.off_endpoint = o; o += 16 * L.n_endpoints;
    L.off_step = o;     o += 16 * L.n_steps;
    L.off_spike = o;    o += 24 * L.n_spike;
    L.off_outage = o;   o += 16 * L.n_outage;
    L.off_samp_sum = o; o += 8 * L.n_series;
    L.off_rq_next = o;  o += 4 * L.rq_smem;
    L.off_samp_max = o; o += 4 * L.n_series;
    L.off_lb = o;       o += 4 * L.n_lb_edges;
    L.warp_bytes = align_up(o, 16);
}

// Device-global pointers (read-only scenario, sweep rows, spill tiers, outputs).
struct Globals {
    const AfEdge* edges; const AfServer* servers; const AfEndpoint* endpoints;
    const AfStep* steps; const int32_t* lb_edges; const AfSpikeMark* spikes;
    const AfOutageMark* outages;
    const AfSweepColumn* sweep_cols; const double* sweep_vals; uint64_t sweep_first, sweep_rows;
    // spill tiers, one region per resident warp slot
    double* spill_ev_time; uint64_t* spill_ev_key; ReqRec* spill_rq_rec; uint32_t* spill_rq_next;
    // outputs, indexed by local replica
    AfReplicaStats* stats; uint32_t* edge_sent; uint32_t* edge_dropped;
    uint32_t* hist; uint32_t* thr; uint64_t* samp_sum; uint32_t* samp_max;
    double* trace_clocks; uint32_t* trace_series; uint32_t* trace_counts;
    unsigned long long* work_counter;
    uint64_t seed, replica_begin, n_replicas;
};

// ---- warp primitives (a warp of ONE lane on the host) ------------------------
#if AF_DEVICE_CODE
constexpr int WARP = 32;
#define AF_FULL 0xFFFFFFFFu
__device__ __forceinline__ int lane_id() { return (int)(threadIdx.x & 31u); }
__device__ __forceinline__ uint32_t w_min(uint32_t v) { return __reduce_min_sync(AF_FULL, v); }
__device__ __forceinline__ uint32_t w_ballot(bool p) { return __ballot_sync(AF_FULL, p); }
__device__ __forceinline__ uint32_t w_shfl(uint32_t v, int src) { return __shfl_sync(AF_FULL, v, src); }
__device__ __forceinline__ void w_sync() { __syncwarp(); }
__device__ __forceinline__ void red_add_u32(uint32_t* p, uint32_t v) { atomicAdd(p, v); }
#else
constexpr int WARP = 1;
inline int lane_id() { return 0; }
inline uint32_t w_min(uint32_t v) { return v; }
inline uint32_t w_ballot(bool p) { return p ? 1u : 0u; }
inline uint32_t w_shfl(uint32_t v, int) { return v; }
inline void w_sync() {}
inline void red_add_u32(uint32_t* p, uint32_t v) { *p += v; }
#endif
AF_HD int first_lane(uint32_t ballot) {
#if AF_DEVICE_CODE
    return __ffs((int)ballot) - 1;
#else
    (void)ballot; return 0;
#endif
}

// ---- one replica --------------------------------------------------------------
struct Replica {
    const Layout& L;
    const Globals& G;
    // workspace views
    double* ev_time; uint64_t* ev_key; ReqRec* rq_rec; uint32_t* rq_next;
    EdgeS* edge; ServerS* server; EndpointS* endpoint; StepS* step; uint32_t* lb;
    SpikeS* spike; OutageS* outage; uint64_t* samp_sum; uint32_t* samp_max;
    double* sp_ev_time; uint64_t* sp_ev_key; ReqRec* sp_rq_rec; uint32_t* sp_rq_next;
    // identity
    uint64_t replica; uint64_t local; int lane;
    // clock + queue
    double now; double horizon; uint32_t seq;
    int32_t ev_hw, ev_live, ev_last_free, ev_hole; uint32_t peak_ev;
    // request table
    uint32_t rq_free, rq_hw, rq_live, peak_rq;
    // generator (two clocks: the sampler's virtual one and the simulation's)
    double g_vnow, g_window_end, g_lam; uint32_t g_pos; uint32_t generated; bool g_done;
    bool need_arrival; uint32_t arm_seq;
    // parameters that may be swept
    double users_mean, users_sigma, rate_per_user;
    // load balancer
    int32_t lb_n;
    // timelines
    int32_t spike_cur, outage_cur;
    // sampler
    double tick_time; uint32_t tick_seq; uint32_t n_ticks;
    // results
    uint32_t completed, flags; uint64_t n_events;
    double lat_sum, lat_sumsq, lat_min, lat_max;
    bool traced;

    AF_HD Replica(const Layout& l, const Globals& g) : L(l), G(g) {}

    // -------------------------------------------------------------- storage
    AF_HD ReqRec rq_load(uint32_t s) const {
        return (int32_t)s < L.rq_smem ? rq_rec[s] : sp_rq_rec[s - L.rq_smem];
    }
    AF_HD void rq_store(uint32_t s, const ReqRec& r) {
        if ((int32_t)s < L.rq_smem) rq_rec[s] = r; else sp_rq_rec[s - L.rq_smem] = r;
    }
    AF_HD void rq_set_pack(uint32_t s, uint32_t pack) {
        if ((int32_t)s < L.rq_smem) rq_rec[s].pack = pack; else sp_rq_rec[s - L.rq_smem].pack = pack;
    }
    AF_HD uint32_t nx_load(uint32_t s) const {
        return (int32_t)s < L.rq_smem ? rq_next[s] : sp_rq_next[s - L.rq_smem];
    }
    AF_HD void nx_store(uint32_t s, uint32_t v) {
        if ((int32_t)s < L.rq_smem) rq_next[s] = v; else sp_rq_next[s - L.rq_smem] = v;
    }
    AF_HD uint64_t evt_load(int32_t k) const {
        return afr::d2u(k < L.ev_smem ? ev_time[k] : sp_ev_time[k - L.ev_smem]);
    }
    AF_HD uint64_t evk_load(int32_t k) const {
        return k < L.ev_smem ? ev_key[k] : sp_ev_key[k - L.ev_smem];
    }
    AF_HD void ev_store(int32_t k, uint64_t tbits, uint64_t key) {
        if (k < L.ev_smem) { ev_time[k] = afr::u2d(tbits); ev_key[k] = key; }
        else { sp_ev_time[k - L.ev_smem] = afr::u2d(tbits); sp_ev_key[k - L.ev_smem] = key; }
    }
    AF_HD void ev_mark_free(int32_t k) {
        if (k < L.ev_smem) ev_time[k] = afr::u2d(INF_BITS); else sp_ev_time[k - L.ev_smem] = afr::u2d(INF_BITS);
    }

    // -------------------------------------------------------------- request slots
    AF_HD uint32_t rq_alloc() {
        uint32_t s;
        if (rq_free != NIL) { s = rq_free; rq_free = nx_load(s); }
        else if ((int32_t)rq_hw < L.rq_total) { s = rq_hw++; }
        else { flags |= AF_FLAG_REQUEST_OVERFLOW; return NIL; }
        ++rq_live;
        if (rq_live > peak_rq) peak_rq = rq_live;
        return s;
    }
    AF_HD void rq_release(uint32_t s) { nx_store(s, rq_free); rq_free = s; --rq_live; }

    AF_HD void fifo_push(uint32_t& head, uint32_t& tail, uint32_t s) {
        nx_store(s, NIL);
        if (tail == NIL) head = s; else nx_store(tail, s);
        tail = s;
    }
    AF_HD uint32_t fifo_pop(uint32_t& head, uint32_t& tail) {
        uint32_t s = head;
        head = nx_load(s);
        if (head == NIL) tail = NIL;
        return s;
    }

    // -------------------------------------------------------------- event pool
    // An unsorted, lane-strided array: slot k belongs to lane k % 32.  push is O(1) -- it
    // reuses the slot freed by the last pop, else a hole the last pop's scan noticed, else
    // appends; pop is a warp arg-min over (time, seq).  Every lane executes push (same slot,
    // same values), only the owner lane ever reads the slot back.
    AF_HD void push_seq(double t, uint32_t payload, uint32_t s) {
        if (!(t < horizon)) return;   // env.run(until=T): events at >= T never fire
        int32_t slot;
        if (ev_last_free >= 0) { slot = ev_last_free; ev_last_free = -1; }
        else if (ev_hole >= 0) { slot = ev_hole; ev_hole = -1; }
        else {
            if (ev_hw >= L.ev_total) { flags |= AF_FLAG_EVENT_OVERFLOW; return; }
            slot = ev_hw++;
        }
        ev_store(slot, afr::d2u(t), ((uint64_t)s << 32) | payload);
        ++ev_live;
        if ((uint32_t)ev_live > peak_ev) peak_ev = (uint32_t)ev_live;
    }
    AF_HD void push(double t, uint32_t payload) { push_seq(t, payload, seq++); }

    // pop the (time, seq)-minimum; false when the pool is empty
    AF_HD bool pop(double& t, uint32_t& payload, uint32_t& ev_seq) {
        if (ev_live == 0) return false;
        uint64_t bt = ~0ull, bk = ~0ull; int32_t bi = -1; int32_t hole = 0x7FFFFFFF;
        for (int32_t k = lane; k < ev_hw; k += WARP) {
            uint64_t tb = evt_load(k);
            if (tb == INF_BITS) { if (k < hole) hole = k; continue; }
            uint64_t kk = evk_load(k);
            if (tb < bt || (tb == bt && kk < bk)) { bt = tb; bk = kk; bi = k; }
        }
#if AF_DEVICE_CODE
        uint32_t hi = (uint32_t)(bt >> 32), lo = (uint32_t)bt;
        uint32_t mhi = w_min(hi);
        bool cand = (hi == mhi) && bi >= 0;
        uint32_t mlo = w_min(cand ? lo : 0xFFFFFFFFu);
        cand = cand && lo == mlo;
        uint32_t b = w_ballot(cand);
        if (__popc(b) > 1) {          // equal times: the earlier push wins (SimPy eid order)
            uint32_t sq = (uint32_t)(bk >> 32);
            uint32_t msq = w_min(cand ? sq : 0xFFFFFFFFu);
            cand = cand && sq == msq;
            b = w_ballot(cand);
        }
        int owner = __ffs((int)b) - 1;
        uint32_t k_hi = w_shfl((uint32_t)(bk >> 32), owner), k_lo = w_shfl((uint32_t)bk, owner);
        int32_t slot = (int32_t)w_shfl((uint32_t)bi, owner);
        bt = ((uint64_t)mhi << 32) | mlo;
        bk = ((uint64_t)k_hi << 32) | k_lo;
        hole = (int32_t)w_min((uint32_t)hole);
#else
        int32_t slot = bi;
#endif
        ev_mark_free(slot);
        --ev_live;
        if (slot == ev_hw - 1) { --ev_hw; ev_last_free = -1; }
        else ev_last_free = slot;
        ev_hole = hole < ev_hw ? hole : -1;
        t = afr::u2d(bt);
        payload = (uint32_t)bk;
        ev_seq = (uint32_t)(bk >> 32);
        return true;
    }

    // -------------------------------------------------------------- generator
    // samplers/poisson_poisson.py:52-82 / gaussian_poisson.py:64-94.  Returns false
    // when the sampler is exhausted; otherwise the next yielded gap.
    AF_HD bool gen_next_gap(double& gap) {
        const double T = horizon;
        for (;;) {
            if (!(g_vnow < T)) return false;
            if (g_vnow >= g_window_end) {
                g_window_end = g_vnow + (double)L.window_s;
                afr::GenDraw d = afr::gen_users(G.seed, replica, g_pos, L.users_dist, users_mean, users_sigma);
                g_pos = d.pos;
                g_lam = d.value * rate_per_user;
            }
            if (g_lam <= 0.0) { g_vnow = g_window_end; continue; }
            afr::Src s = afr::make_gen(G.seed, replica, g_pos);
            double u = s.next53();
            g_pos = s.pos;
            if (u < 1e-15) u = 1e-15;               // max(u, 1e-15)
            double dt = afr::af_div(-afr::af_log(1.0 - u), g_lam);
            if (g_vnow + dt > T) return false;
            if (g_vnow + dt >= g_window_end) { g_vnow = g_window_end; continue; }
            g_vnow += dt;
            gap = dt;
            return true;
        }
    }

    // -------------------------------------------------------------- edges
    // EdgeRuntime._deliver up to the timeout (edge.py:73-107)
    AF_HD void edge_send(uint32_t slot, uint32_t e, const ReqRec& r) {
        EdgeS& E = edge[e];
        uint32_t s = seq++;                          // SimPy schedules the timeout here
        afr::EdgeDraw d = afr::edge_draw(G.seed, replica, r.rid, pk_hops(r.pack), (int)(E.meta & 7u),
                                         E.mean, E.sigma, E.dropout);
        E.sent += 1;
        if (d.u < E.dropout) {                      // the request vanishes (edge.py:79-86)
            E.dropped += 1;
            rq_release(slot);
            return;
        }
        E.conn += 1;
        double effective = d.transit + E.spike;    // spike read at SEND time (edge.py:94-106)
        push_seq(now + effective, mk_payload(K_DELIVER, e, slot), s);
    }

    // -------------------------------------------------------------- server
    AF_HD void grant_cpu_waiter(uint32_t sidx) {
        // Container.put(1) processed -> head of the CPU get-queue resumes (server.py:222-231)
        ServerS& S = server[sidx];
        if (S.cpuq_head == NIL) return;
        uint32_t w = fifo_pop(S.cpuq_head, S.cpuq_tail);
        S.cpu_free -= 1;
        S.ready_q -= 1;
        ReqRec r = rq_load(w);
        r.pack |= PK_CORE;
        rq_set_pack(w, r.pack);
        const EndpointS& ep = endpoint[pk_ep(r.pack)];
        push(now + step[ep.step_begin + pk_step(r.pack)].dur, mk_payload(K_STEP_END, sidx, w));
    }

    // the `for step in selected_endpoint.steps` loop from the current step (server.py:197-255).
    // Returns true when no step is left (the caller then runs finish_request).
    AF_HD bool run_steps(uint32_t slot, uint32_t sidx, ReqRec& r) {
        ServerS& S = server[sidx];
        const EndpointS ep = endpoint[pk_ep(r.pack)];
        uint32_t st = pk_step(r.pack);
        if (st >= ep.n_steps) return true;
        const StepS sp = step[ep.step_begin + st];
        if (sp.kind == AF_STEP_CPU) {
            if (r.pack & PK_IO) { r.pack &= ~PK_IO; S.io_q -= 1; }
            if (!(r.pack & PK_CORE)) {
                if (S.cpu_free > 0) { S.cpu_free -= 1; r.pack |= PK_CORE; }
                else {                              // cpu_req not triggered -> ready queue
                    S.ready_q += 1;
                    rq_set_pack(slot, r.pack);
                    fifo_push(S.cpuq_head, S.cpuq_tail, slot);
                    return false;
                }
            }
            rq_set_pack(slot, r.pack);
            push(now + sp.dur, mk_payload(K_STEP_END, sidx, slot));
        } else {
            bool release = (r.pack & PK_CORE) != 0;
            if (release) { r.pack &= ~PK_CORE; S.cpu_free += 1; }
            if (!(r.pack & PK_IO)) { r.pack |= PK_IO; S.io_q += 1; }
            rq_set_pack(slot, r.pack);
            // SimPy order: the releasing request schedules its IO timeout before the
            // woken waiter schedules its CPU timeout (see DESIGN.md "tie rule")
            push(now + sp.dur, mk_payload(K_STEP_END, sidx, slot));
            if (release) grant_cpu_waiter(sidx);
        }
        return false;
    }

    // server.py:257-276
    AF_HD void finish_request(uint32_t slot, uint32_t sidx, ReqRec r) {
        ServerS& S = server[sidx];
        const uint32_t total_ram = endpoint[pk_ep(r.pack)].total_ram;
        // SimPy order of the pushes that follow a release (DESIGN.md "tie rule"):
        //   core + RAM : woken CPU waiter, then this request's edge, then RAM waiters
        //   core only  : this request's edge (Initialize is URGENT), then the CPU waiter
        bool had_core = (r.pack & PK_CORE) != 0;
        if (had_core) { r.pack &= ~PK_CORE; S.cpu_free += 1; }
        if (had_core && total_ram) grant_cpu_waiter(sidx);
        if (r.pack & PK_IO) { r.pack &= ~PK_IO; S.io_q -= 1; }
        if (total_ram) {
            S.ram_in_use -= (int32_t)total_ram;
            S.ram_free += (int32_t)total_ram;
        }
        rq_set_pack(slot, r.pack);
        edge_send(slot, S.out_edge, r);
        if (had_core && !total_ram) grant_cpu_waiter(sidx);
        if (total_ram) {
            // Container FIFO with head-of-line blocking
            while (S.ramq_head != NIL) {
                uint32_t w = S.ramq_head;
                ReqRec wr = rq_load(w);
                uint32_t need = endpoint[pk_ep(wr.pack)].total_ram;
                if ((int32_t)need > S.ram_free) break;
                fifo_pop(S.ramq_head, S.ramq_tail);
                S.ram_free -= (int32_t)need;
                S.ram_in_use += (int32_t)need;
                if (run_steps(w, sidx, wr)) {
                    // an endpoint made of RAM steps only: it gives the memory straight back
                    S.ram_in_use -= (int32_t)need;
                    S.ram_free += (int32_t)need;
                    edge_send(w, S.out_edge, wr);
                }
            }
        }
    }

    // ServerRuntime._dispatcher + head of _handle_request (server.py:88-149, 303-313)
    AF_HD void server_arrive(uint32_t slot, uint32_t sidx, ReqRec r) {
        ServerS& S = server[sidx];
        r.pack += 1;                                 // record_hop(SERVER)
        uint32_t epi = 0;
        if (S.n_ep > 1) {
            afr::Src s = afr::make_request(G.seed, replica, afr::P_SERVER, r.rid, pk_hops(r.pack));
            s.load(0);
            epi = (uint32_t)(((uint64_t)s.w.x * S.n_ep) >> 32);
        }
        uint32_t ep_global = S.ep_begin + epi;
        r.pack = (r.pack & 0xFFu) | (ep_global << 16); // step 0, flags clear
        rq_set_pack(slot, r.pack);
        uint32_t total_ram = endpoint[ep_global].total_ram;
        if (total_ram) {
            if (S.ramq_head == NIL && (int32_t)total_ram <= S.ram_free) {
                S.ram_free -= (int32_t)total_ram;
                S.ram_in_use += (int32_t)total_ram;
            } else {
                fifo_push(S.ramq_head, S.ramq_tail, slot);
                return;
            }
        }
        if (run_steps(slot, sidx, r)) finish_request(slot, sidx, r);
    }

    // -------------------------------------------------------------- client: completion
    AF_HD void complete(uint32_t slot, const ReqRec& r) {
        double lat = now - r.t0;                     // finish - start (analyzer.py:86-89)
        completed += 1;
        lat_sum += lat;
        lat_sumsq += lat * lat;
        if (lat < lat_min) lat_min = lat;
        if (lat > lat_max) lat_max = lat;
        if (lane == 0) {
            if (L.collect_hist) {
                int32_t idx = (int32_t)(afr::d2u(lat) >> (52 - AF_HIST_SUB_BITS))
                              - ((1023 + AF_HIST_MIN_EXP) << AF_HIST_SUB_BITS);
                idx = idx < 0 ? 0 : (idx >= AF_HIST_BINS ? AF_HIST_BINS - 1 : idx);
                red_add_u32(&G.hist[local * AF_HIST_BINS + (uint32_t)idx], 1u);
            }
            if (L.collect_thr) {
                // bucket k counts (k, k+1] (analyzer.py:108-125)
                int32_t b = (int32_t)ceil(now) - 1;
                b = b < 0 ? 0 : b;
                if (b < L.horizon_s) red_add_u32(&G.thr[local * (uint64_t)L.horizon_s + (uint32_t)b], 1u);
            }
            if (traced) {
                uint32_t i = completed - 1;
                if ((int32_t)i < L.trace_clock_cap) {
                    double* p = G.trace_clocks + (local * (uint64_t)L.trace_clock_cap + i) * 2;
                    p[0] = r.t0; p[1] = now;
                }
            }
        }
        if (traced && (int32_t)(completed - 1) >= L.trace_clock_cap) flags |= AF_FLAG_TRACE_TRUNCATED;
        rq_release(slot);
    }

    // -------------------------------------------------------------- deliveries
    AF_HD void on_deliver(uint32_t slot, uint32_t e) {
        EdgeS& E = edge[e];
        E.conn -= 1;
        ReqRec r = rq_load(slot);
        r.pack += 1;                                 // record_hop(edge)
        uint32_t tk = (E.meta >> 3) & 3u;
        if (tk == AF_TARGET_CLIENT) {
            r.pack += 1;                             // record_hop(client)
            if (pk_hops(r.pack) > 3) { complete(slot, r); return; }     // client.py:62
            rq_set_pack(slot, r.pack);
            edge_send(slot, (uint32_t)L.client_edge, r);
        } else if (tk == AF_TARGET_LB) {
            r.pack += 1;                             // record_hop(LB)
            rq_set_pack(slot, r.pack);
            uint32_t pick;
            if (L.lb_algo == AF_LB_ROUND_ROBIN) {    // lb_algorithms.py:22-36
                pick = lb[0];
                for (int32_t i = 1; i < lb_n; ++i) lb[i - 1] = lb[i];
                lb[lb_n - 1] = pick;
            } else {                                 // least_connections, :10-20 (first min wins)
                pick = lb[0];
                uint32_t best = edge[pick].conn;
                for (int32_t i = 1; i < lb_n; ++i) {
                    uint32_t c = edge[lb[i]].conn;
                    if (c < best) { best = c; pick = lb[i]; }
                }
            }
            edge_send(slot, pick, r);
        } else {
            server_arrive(slot, E.meta >> 5, r);
        }
    }

    // -------------------------------------------------------------- arrivals
    AF_HD void on_arrival() {
        generated += 1;
        uint32_t slot = rq_alloc();
        // the generator asks the sampler for the next gap right after transport(): its timeout is
        // scheduled BEFORE the edge's delivery timeout (rqs_generator.py:103-119).  The seq is
        // reserved here; the gap itself is drawn at the single arm_generator() site in run().
        arm_seq = seq++;
        need_arrival = true;
        if (slot == NIL) return;
        ReqRec r; r.t0 = now; r.rid = generated; r.pack = 1;  // record_hop(generator)
        rq_store(slot, r);
        edge_send(slot, (uint32_t)L.gen_edge, r);
    }

    AF_HD void arm_generator() {
        need_arrival = false;
        double gap;
        if (!g_done && gen_next_gap(gap)) push_seq(now + gap, mk_payload(K_ARRIVAL, 0, 0), arm_seq);
        else g_done = true;
    }

    // -------------------------------------------------------------- injection
    AF_HD void on_spike() {       // injection.py:167-198: all marks of this instant, then re-arm
        double t = spike[spike_cur].fire;
        while (spike_cur < L.n_spike && spike[spike_cur].fire == t) {
            const SpikeS m = spike[spike_cur];
            edge[m.edge].spike = edge[m.edge].spike + m.delta;
            ++spike_cur;
        }
        if (spike_cur < L.n_spike) push(spike[spike_cur].fire, mk_payload(K_SPIKE, 0, 0));
    }
    AF_HD void on_outage() {      // injection.py:201-226
        double t = outage[outage_cur].fire;
        while (outage_cur < L.n_outage && outage[outage_cur].fire == t) {
            const OutageS m = outage[outage_cur];
            ++outage_cur;
            if (m.lb_edge < 0) continue;
            int32_t at = -1;
            for (int32_t i = 0; i < lb_n; ++i) if (lb[i] == (uint32_t)m.lb_edge) { at = i; break; }
            if (at >= 0) {                           // pop (DOWN) or move_to_end (UP)
                for (int32_t i = at + 1; i < lb_n; ++i) lb[i - 1] = lb[i];
                --lb_n;
            }
            if (!m.down) lb[lb_n++] = (uint32_t)m.lb_edge;
        }
        if (outage_cur < L.n_outage) push(outage[outage_cur].fire, mk_payload(K_OUTAGE, 0, 0));
    }

    // -------------------------------------------------------------- sampled metrics
    AF_HD uint32_t series_value(int32_t j) const {
        if (j < 3 * L.n_servers) {
            const ServerS& S = server[j / 3];
            int m = j % 3;
            return (uint32_t)(m == 0 ? S.ready_q : (m == 1 ? S.io_q : S.ram_in_use));
        }
        return edge[j - 3 * L.n_servers].conn;
    }
    AF_HD bool series_enabled(int32_t j) const {
        if (j < 3 * L.n_servers) return (L.metrics_mask & 7u) == 7u;   // collector.py:60-63
        return (L.metrics_mask & AF_METRIC_EDGE_CONN) != 0;
    }
    // emit every collector tick ordered before (t, ev_seq)  (collector.py:50-66)
    AF_HD void take_samples(double t, uint32_t ev_seq) {
        while (tick_time < t || (tick_time == t && tick_seq < ev_seq)) {
            if (!(tick_time < horizon)) return;
            for (int32_t j = lane; j < L.n_series; j += WARP) {
                if (!series_enabled(j)) continue;
                uint32_t v = series_value(j);
                samp_sum[j] += v;
                if (v > samp_max[j]) samp_max[j] = v;
                if (traced && (int32_t)n_ticks < L.trace_tick_cap)
                    G.trace_series[(local * (uint64_t)L.n_series + (uint32_t)j) * (uint64_t)L.trace_tick_cap + n_ticks] = v;
            }
            n_ticks += 1;
            tick_seq = seq++;                         // the collector re-arms its timeout here
            tick_time = tick_time + L.sample_period;
        }
    }

    // -------------------------------------------------------------- set-up
    AF_HD void bind(unsigned char* ws, uint64_t warp_slot) {
        ev_time = (double*)(ws + L.off_ev_time);
        ev_key = (uint64_t*)(ws + L.off_ev_key);
        rq_rec = (ReqRec*)(ws + L.off_rq_rec);
        rq_next = (uint32_t*)(ws + L.off_rq_next);
        edge = (EdgeS*)(ws + L.off_edge);
        server = (ServerS*)(ws + L.off_server);
        endpoint = (EndpointS*)(ws + L.off_endpoint);
        step = (StepS*)(ws + L.off_step);
        lb = (uint32_t*)(ws + L.off_lb);
        spike = (SpikeS*)(ws + L.off_spike);
        outage = (OutageS*)(ws + L.off_outage);
        samp_sum = (uint64_t*)(ws + L.off_samp_sum);
        samp_max = (uint32_t*)(ws + L.off_samp_max);
        uint64_t ev_sp = (uint64_t)(L.ev_total - L.ev_smem), rq_sp = (uint64_t)(L.rq_total - L.rq_smem);
        sp_ev_time = G.spill_ev_time + warp_slot * ev_sp;
        sp_ev_key = G.spill_ev_key + warp_slot * ev_sp;
        sp_rq_rec = G.spill_rq_rec + warp_slot * rq_sp;
        sp_rq_next = G.spill_rq_next + warp_slot * rq_sp;
    }

    AF_HD void load_params() {
        for (int32_t i = lane; i < L.n_edges; i += WARP) {
            const AfEdge a = G.edges[i];
            EdgeS e;
            e.mean = a.mean; e.sigma = a.sigma; e.dropout = a.dropout; e.spike = 0.0;
            e.meta = (uint32_t)a.dist | ((uint32_t)a.target_kind << 3) | ((uint32_t)a.target_index << 5);
            e.conn = 0; e.sent = 0; e.dropped = 0;
            edge[i] = e;
        }
        for (int32_t i = lane; i < L.n_servers; i += WARP) {
            const AfServer a = G.servers[i];
            ServerS s;
            s.cpu_free = a.cpu_cores; s.ram_free = a.ram_mb; s.ready_q = 0; s.io_q = 0; s.ram_in_use = 0;
            s.ramq_head = s.ramq_tail = s.cpuq_head = s.cpuq_tail = NIL;
            s.out_edge = (uint32_t)a.out_edge; s.ep_begin = (uint32_t)a.endpoint_begin; s.n_ep = (uint32_t)a.n_endpoints;
            server[i] = s;
        }
        for (int32_t i = lane; i < L.n_endpoints; i += WARP) {
            const AfEndpoint a = G.endpoints[i];
            EndpointS e; e.step_begin = (uint32_t)a.step_begin; e.n_steps = (uint32_t)a.n_steps;
            e.total_ram = (uint32_t)a.total_ram; e.pad = 0;
            endpoint[i] = e;
        }
        for (int32_t i = lane; i < L.n_steps; i += WARP) {
            const AfStep a = G.steps[i];
            StepS s; s.dur = a.duration; s.kind = (uint32_t)a.kind; s.pad = 0;
            step[i] = s;
        }
        for (int32_t i = lane; i < L.n_lb_edges; i += WARP) lb[i] = (uint32_t)G.lb_edges[i];
        for (int32_t i = lane; i < L.n_spike; i += WARP) {
            const AfSpikeMark a = G.spikes[i];
            SpikeS s; s.fire = a.fire_time; s.delta = a.delta; s.edge = (uint32_t)a.edge; s.pad = 0;
            spike[i] = s;
        }
        for (int32_t i = lane; i < L.n_outage; i += WARP) {
            const AfOutageMark a = G.outages[i];
            OutageS o; o.fire = a.fire_time; o.lb_edge = a.lb_edge; o.down = a.down;
            outage[i] = o;
        }
        for (int32_t i = lane; i < L.n_series; i += WARP) { samp_sum[i] = 0; samp_max[i] = 0; }
        for (int32_t i = lane; i < L.ev_smem; i += WARP) ev_time[i] = afr::u2d(INF_BITS);
        w_sync();
        users_mean = L.users_mean; users_sigma = L.users_sigma; rate_per_user = L.rate_per_user;
        // sweep overrides of this replica (uniform: every lane applies every column)
        if (L.n_sweep_cols > 0 && replica >= G.sweep_first && replica - G.sweep_first < G.sweep_rows) {
            const double* row = G.sweep_vals + (replica - G.sweep_first) * (uint64_t)L.n_sweep_cols;
            for (int32_t c = 0; c < L.n_sweep_cols; ++c) {
                const AfSweepColumn col = G.sweep_cols[c];
                double v = row[c];
                switch (col.field) {
                case AF_FIELD_USERS_MEAN: users_mean = v; break;
                case AF_FIELD_USERS_SIGMA: users_sigma = v; break;
                case AF_FIELD_RATE_PER_USER: rate_per_user = v; break;
                case AF_FIELD_EDGE_MEAN: edge[col.index].mean = v; break;
                case AF_FIELD_EDGE_SIGMA: edge[col.index].sigma = v; break;
                case AF_FIELD_EDGE_DROPOUT: edge[col.index].dropout = v; break;
                case AF_FIELD_SERVER_CPU_CORES: server[col.index].cpu_free = (int32_t)v; break;
                case AF_FIELD_SERVER_RAM_MB: server[col.index].ram_free = (int32_t)v; break;
                case AF_FIELD_STEP_DURATION: step[col.index].dur = v; break;
                case AF_FIELD_ENDPOINT_RAM: endpoint[col.index].total_ram = (uint32_t)v; break;
                case AF_FIELD_SPIKE_DELTA:
                    spike[col.index].delta = spike[col.index].delta < 0.0 ? -v : v; break;
                default: break;
                }
            }
            w_sync();
        }
    }

    // -------------------------------------------------------------- the replica
    AF_HD void run(uint64_t local_index) {
        local = local_index;
        replica = G.replica_begin + local_index;
        lane = lane_id();
        load_params();
        now = 0.0; horizon = (double)L.horizon_s; seq = 0;
        ev_hw = 0; ev_live = 0; ev_last_free = -1; ev_hole = -1; peak_ev = 0;
        rq_free = NIL; rq_hw = 0; rq_live = 0; peak_rq = 0;
        g_vnow = 0.0; g_window_end = 0.0; g_lam = 0.0; g_pos = 0; generated = 0; g_done = false;
        lb_n = L.n_lb_edges;
        spike_cur = 0; outage_cur = 0;
        n_ticks = 0; completed = 0; flags = 0; n_events = 0;
        lat_sum = 0.0; lat_sumsq = 0.0; lat_min = afr::u2d(INF_BITS); lat_max = 0.0;
        traced = (int64_t)local_index < (int64_t)L.trace_replicas;

        // start order of the reference (simulation_runner.py:339-342, 301-336):
        // spike timeline, outage timeline, generator, ..., collector
        if (L.n_spike > 0) {
            if (spike[0].fire == 0.0) on_spike(); else push(spike[0].fire, mk_payload(K_SPIKE, 0, 0));
        }
        if (L.n_outage > 0) {
            if (outage[0].fire == 0.0) on_outage(); else push(outage[0].fire, mk_payload(K_OUTAGE, 0, 0));
        }
        arm_seq = seq++; need_arrival = true;
        tick_seq = seq++;
        tick_time = 0.0 + L.sample_period;

        double t; uint32_t payload, ev_seq;
        for (;;) {
            if (need_arrival) arm_generator();
            if (!pop(t, payload, ev_seq)) break;
            take_samples(t, ev_seq);
            now = t;
            ++n_events;
            uint32_t kind = payload >> 29, aux = (payload >> SLOT_BITS) & AUX_MASK, slot = payload & SLOT_MASK;
            if (kind == K_DELIVER) on_deliver(slot, aux);
            else if (kind == K_STEP_END) {
                ReqRec r = rq_load(slot);
                r.pack += 1u << 8;                   // next step
                if (run_steps(slot, aux, r)) finish_request(slot, aux, r);
            }
            else if (kind == K_ARRIVAL) on_arrival();
            else if (kind == K_SPIKE) on_spike();
            else on_outage();
            if (flags & (AF_FLAG_EVENT_OVERFLOW | AF_FLAG_REQUEST_OVERFLOW)) break;
        }
        take_samples(horizon, 0u);                    // ticks strictly before the horizon
        w_sync();
        write_back();
    }

    AF_HD void write_back() {
        for (int32_t i = lane; i < L.n_edges; i += WARP) {
            G.edge_sent[local * (uint64_t)L.n_edges + (uint32_t)i] = edge[i].sent;
            G.edge_dropped[local * (uint64_t)L.n_edges + (uint32_t)i] = edge[i].dropped;
        }
        for (int32_t j = lane; j < L.n_series; j += WARP) {
            G.samp_sum[local * (uint64_t)L.n_series + (uint32_t)j] = samp_sum[j];
            G.samp_max[local * (uint64_t)L.n_series + (uint32_t)j] = samp_max[j];
        }
        if (lane == 0) {
            AfReplicaStats st;
            st.n_events = n_events; st.generated = generated; st.completed = completed;
            st.flags = flags; st.n_ticks = n_ticks; st.peak_events = peak_ev; st.peak_requests = peak_rq;
            st.lat_sum = lat_sum; st.lat_sumsq = lat_sumsq;
            st.lat_min = completed ? lat_min : 0.0; st.lat_max = lat_max;
            st.p50 = st.p95 = st.p99 = afr::u2d(0x7FF8000000000000ull);
            G.stats[local] = st;
            if (traced) { G.trace_counts[local * 2] = completed; G.trace_counts[local * 2 + 1] = n_ticks; }
        }
    }
};

}  // namespace afc
