// rr_gateway.cu — the per-request path in native code: concurrent admission -> K1 -> engine hand-off, no Python hop.
//
// The reference's concurrency is N blocked client threads into a single-worker gateway process
// (reference src/demo_load_balancing.py:195-203, bin/start-gateway.sh:54).  Here every client thread calls
// rr_gateway_submit (non-blocking, any thread): the request goes into an admission queue; ONE dispatcher thread
// drains the queue, turns everything that is pending — ADMITs of new requests, DONE / FAIL reports of finished ones,
// re-ADMITs of requests whose backend failed (the fallback walk) — into one ordered event trace, stages it in pinned
// memory, and processes it with ONE launch of the K1 router kernel (rr_router.cu: events and decisions live in the
// router's HBM-resident rings); decisions come back in one D2H copy and admitted requests go straight into the queue
// of the engine that serves the picked deployment (engine_submit_tagged), whose worker thread reports completion
// through a hook.  Concurrent admissions are therefore COALESCED: under a burst of N callers the path costs
// ~N / batch launches instead of 2 N, and the serialised-trace semantics of K1 are unchanged (the trace is the
// arrival order at the queue; it can be recorded and replayed through the oracle, tests/test_gateway_gpu.py).
//
// Replaces: litellm.Router's async request handling behind the proxy (reference pyproject.toml:8,
// config/config.yaml:100-108) for the call shape of reference src/demo_fallback.py:143-147.
#include "rr_kernels.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <new>
#include <string.h>
#include <thread>
#include <unordered_map>
#include <vector>

#define RR_API extern "C" __attribute__((visibility("default")))

namespace {

enum ReqState { Q_ADMIT = 0, RUNNING = 1, FINISHED = 2 };

struct GwReq {
    uint64_t id = 0;
    int group = 0, max_new = 0;
    std::vector<int32_t> prompt;
    int chain_start = 0, attempts = 0;
    ReqState state = Q_ADMIT;
    int status = RR_OK;
    int deployment = -1, served_group = -1, chain_pos = 0, replica = -1;
    rr_engine* eng = nullptr;
    uint64_t eng_ticket = 0;
    std::vector<int32_t> tokens;
    double t_submit = 0, t_admit = 0, t_first = 0, t_done = 0, t_eng_submit = 0;
    bool cancelled = false, cancel_as_failure = false, abandoned = false;
};

struct Pending {
    int type;          // RR_EV_*
    GwReq* req;        // ADMIT
    int target;        // DONE / FAIL: deployment; BURST: group
    int tokens;        // DONE: completion tokens; BURST: burst size
};

}  // namespace

struct rr_gateway {
    rr_router* router = nullptr;
    std::vector<rr_engine*> engines;        // index = position in the create() arrays
    std::unordered_map<int, int> by_replica;   // replica id -> index into engines
    int n_deps = 0, n_groups = 0;
    int min_ctx = 0, min_pf = 0, min_vocab = 0;
    rr_gateway_opts opts;

    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<Pending> pending;
    std::unordered_map<uint64_t, GwReq*> table;
    uint64_t next_id = 1;
    std::thread dispatcher;
    bool stop = false;
    bool busy = false;                     // the dispatcher is between taking a batch and publishing its decisions
    std::atomic<int64_t> manual_now_ms{0};
    std::chrono::steady_clock::time_point t0;

    // trace recording (ring of the last `record_trace` events)
    std::vector<rr_event> tr_ev;
    std::vector<rr_decision> tr_dec;
    size_t tr_count = 0;

    rr_gateway_stats st;
};

static double gw_now_s(rr_gateway* g) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - g->t0).count();
}
static int64_t gw_clock_ms(rr_gateway* g) {
    if (g->opts.manual_clock) return g->manual_now_ms.load(std::memory_order_relaxed);
    return std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
}

// caller holds g->mu
static void gw_finish(rr_gateway* g, GwReq* r, int status) {
    r->status = status;
    r->state = FINISHED;
    r->t_done = gw_now_s(g);
    if (status == RR_OK) g->st.completed++;
    else if (status == RR_RATE_LIMITED) g->st.rate_limited++;
    else g->st.failed++;
    if (r->abandoned) {
        g->table.erase(r->id);
        delete r;
    }
}

// Engine worker thread (or the submitting thread for an injected failure): a request of ours finished on `eng`.
static void gw_on_done(void* ctx, uint64_t tag, uint64_t eng_ticket, int status, int n_generated) {
    rr_gateway* g = (rr_gateway*)ctx;
    if (tag == 0) return;                                    // not submitted through this gateway
    GwReq* r = nullptr;
    rr_engine* eng = nullptr;
    int max_new = 0;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        auto it = g->table.find(tag);
        if (it == g->table.end()) return;
        r = it->second;
        eng = r->eng;
        max_new = r->max_new;
    }
    // consume the engine's record (tokens + timestamps); a cancelled request has none left
    rr_completion comp;
    memset(&comp, 0, sizeof(comp));
    std::vector<int32_t> toks((size_t)(max_new > 0 ? max_new : 1));
    int wrc = RR_INVALID_ARGUMENT;
    if (eng && status != RR_CANCELLED) wrc = rr_engine_wait(eng, eng_ticket, 0.001, &comp, toks.data(), max_new);
    {
        std::lock_guard<std::mutex> lk(g->mu);
        auto it = g->table.find(tag);
        if (it == g->table.end()) return;
        r = it->second;
        const int dep = r->deployment;
        if (wrc == RR_OK || wrc == RR_BACKEND_FAILED) {
            const int n = comp.n_generated < max_new ? comp.n_generated : max_new;
            r->tokens.assign(toks.begin(), toks.begin() + (n > 0 ? n : 0));
            r->t_first = r->t_eng_submit + (comp.t_first_token_s - comp.t_submit_s);
        }
        if (status == RR_OK) {
            g->pending.push_back(Pending{RR_EV_DONE, nullptr, dep, (int)r->tokens.size()});
            gw_finish(g, r, RR_OK);
        } else if (status == RR_CANCELLED) {
            // the owner gave up: a client-side timeout counts as a backend failure (cooldown), a disconnect does not
            g->pending.push_back(Pending{r->cancel_as_failure ? RR_EV_FAIL : RR_EV_DONE, nullptr, dep, n_generated});
            gw_finish(g, r, r->cancel_as_failure ? RR_TIMEOUT : RR_CANCELLED);
        } else {
            // backend failure: report it, then walk on down the fallback chain with the same request
            g->pending.push_back(Pending{RR_EV_FAIL, nullptr, dep, 0});
            if (r->cancelled) {
                gw_finish(g, r, r->cancel_as_failure ? RR_TIMEOUT : RR_CANCELLED);
            } else {
                r->chain_start = r->chain_pos + 1;
                r->state = Q_ADMIT;
                r->eng = nullptr;
                r->tokens.clear();
                g->st.failed_over++;
                g->pending.push_back(Pending{RR_EV_ADMIT, r, 0, 0});
            }
        }
    }
    g->cv_work.notify_one();
    g->cv_done.notify_all();
}

static void gw_dispatch(rr_gateway* g) {
    std::vector<Pending> batch;
    std::vector<rr_event> ev;
    std::vector<rr_decision> dec;
    std::vector<GwReq*> to_submit;
    const int cap = g->opts.max_batch_events > 0 ? g->opts.max_batch_events : 4096;
    for (;;) {
        batch.clear();
        {
            std::unique_lock<std::mutex> lk(g->mu);
            g->cv_work.wait(lk, [&] { return g->stop || !g->pending.empty(); });
            if (g->stop && g->pending.empty()) return;
            while (!g->pending.empty() && (int)batch.size() < cap) {
                const Pending p = g->pending.front();
                g->pending.pop_front();
                if (p.type == RR_EV_ADMIT && p.req->cancelled) {     // given up before admission: no event, no debit
                    gw_finish(g, p.req, p.req->cancel_as_failure ? RR_TIMEOUT : RR_CANCELLED);
                    continue;
                }
                batch.push_back(p);
            }
            if (batch.empty()) { g->cv_done.notify_all(); continue; }
            g->busy = true;
        }
        const int64_t now = gw_clock_ms(g);
        ev.clear();
        for (const Pending& p : batch) {
            rr_event e;
            e.type = p.type; e.now_ms = now; e.chain_start = 0;
            if (p.type == RR_EV_ADMIT) {
                // (fields of a queued request are only touched by this thread until the decision is published)
                e.target = p.req->group; e.tokens = (int32_t)p.req->prompt.size(); e.chain_start = p.req->chain_start;
            } else {
                e.target = p.target; e.tokens = p.tokens;
            }
            ev.push_back(e);
        }
        dec.assign(ev.size(), rr_decision{RR_INTERNAL, -1, -1, 0});
        const int rc = rr_router_process(g->router, ev.data(), (int)ev.size(), dec.data());
        to_submit.clear();
        {
            std::lock_guard<std::mutex> lk(g->mu);
            g->st.launches++;
            g->st.events += ev.size();
            if ((uint64_t)ev.size() > g->st.max_batch) g->st.max_batch = ev.size();
            if (g->opts.record_trace > 0) {
                for (size_t i = 0; i < ev.size(); ++i) {
                    const size_t slot = g->tr_count % (size_t)g->opts.record_trace;
                    g->tr_ev[slot] = ev[i]; g->tr_dec[slot] = dec[i];
                    g->tr_count++;
                }
            }
            const double t = gw_now_s(g);
            for (size_t i = 0; i < batch.size(); ++i) {
                if (batch[i].type != RR_EV_ADMIT) continue;
                GwReq* r = batch[i].req;
                r->attempts++;
                r->t_admit = t;
                g->st.admit_wait_s += t - r->t_submit;
                const rr_decision& d = dec[i];
                if (rc != RR_OK) { gw_finish(g, r, RR_INTERNAL); continue; }
                r->deployment = d.deployment; r->served_group = d.served_group; r->chain_pos = d.chain_pos;
                if (d.status == RR_OK) {
                    g->st.admitted++;
                    r->replica = rr::router_dep_replica(g->router, d.deployment);
                    to_submit.push_back(r);
                } else if (d.status == RR_RATE_LIMITED && r->attempts > 1) {
                    gw_finish(g, r, RR_BACKEND_FAILED);      // a backend failed and nothing is left on the chain
                } else {
                    gw_finish(g, r, d.status);
                }
            }
        }
        for (GwReq* r : to_submit) {
            auto it = g->by_replica.find(r->replica);
            rr_engine* eng = it == g->by_replica.end() ? nullptr : g->engines[it->second];
            const uint64_t id = r->id;           // `r` may be finished and freed by the done hook once it is submitted
            bool cancelled;
            {
                std::lock_guard<std::mutex> lk(g->mu);
                cancelled = r->cancelled;
                if (!cancelled) { r->eng = eng; r->eng_ticket = 0; r->t_eng_submit = gw_now_s(g); }
            }
            uint64_t tk = 0;
            int src = RR_INTERNAL;
            if (!cancelled && eng) src = rr::engine_submit_tagged(eng, r->prompt.data(), (int)r->prompt.size(), r->max_new, id, &tk);
            bool cancel_now = false;
            std::unique_lock<std::mutex> lk(g->mu);
            if (src == RR_OK) {
                // (the done hook may already have run and finished / re-queued the request: only a still-admitted
                // request becomes RUNNING)
                auto f = g->table.find(id);
                if (f != g->table.end() && f->second->state == Q_ADMIT && f->second->eng == eng && f->second->eng_ticket == 0) {
                    f->second->eng_ticket = tk;
                    f->second->state = RUNNING;
                    cancel_now = f->second->cancelled;       // rr_gateway_cancel raced with the submit: drop the row now
                }
            } else {
                // never reached the backend: give the in-flight slot back (DONE with 0 tokens), fail the request
                g->pending.push_back(Pending{RR_EV_DONE, nullptr, r->deployment, 0});
                gw_finish(g, r, cancelled ? (r->cancel_as_failure ? RR_TIMEOUT : RR_CANCELLED)
                                          : (eng ? src : RR_BACKEND_FAILED));
            }
            lk.unlock();
            if (cancel_now) rr_engine_cancel(eng, tk);
        }
        {
            std::lock_guard<std::mutex> lk(g->mu);
            g->busy = false;
        }
        g->cv_done.notify_all();
    }
}

// ---------------------------------------------------------------------------------------------------- C-ABI
RR_API int rr_gateway_create(rr_router* router, rr_engine* const* engines, const int32_t* replica_ids, int n_engines,
                             const rr_gateway_opts* opts, rr_gateway** out) {
    if (!router || !engines || !replica_ids || n_engines < 1 || !out) return RR_INVALID_ARGUMENT;
    rr_gateway* g = new (std::nothrow) rr_gateway();
    if (!g) return RR_INTERNAL;
    g->router = router;
    memset(&g->opts, 0, sizeof(g->opts));
    if (opts) g->opts = *opts;
    memset(&g->st, 0, sizeof(g->st));
    rr::router_shape(router, &g->n_deps, &g->n_groups);
    g->min_ctx = g->min_pf = g->min_vocab = 0x7fffffff;
    for (int i = 0; i < n_engines; ++i) {
        if (!engines[i]) { delete g; return RR_INVALID_ARGUMENT; }
        g->engines.push_back(engines[i]);
        g->by_replica[replica_ids[i]] = i;
        int c = 0, p = 0, v = 0;
        rr::engine_limits(engines[i], &c, &p, &v);
        if (c < g->min_ctx) g->min_ctx = c;
        if (p < g->min_pf) g->min_pf = p;
        if (v < g->min_vocab) g->min_vocab = v;
    }
    if (g->opts.record_trace > 0) {
        g->tr_ev.resize((size_t)g->opts.record_trace);
        g->tr_dec.resize((size_t)g->opts.record_trace);
    }
    g->t0 = std::chrono::steady_clock::now();
    for (rr_engine* e : g->engines) rr::engine_set_done_hook(e, gw_on_done, g);
    g->dispatcher = std::thread(gw_dispatch, g);
    *out = g;
    return RR_OK;
}

RR_API void rr_gateway_destroy(rr_gateway* g) {
    if (!g) return;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->stop = true;
    }
    g->cv_work.notify_all();
    if (g->dispatcher.joinable()) g->dispatcher.join();
    for (rr_engine* e : g->engines) rr::engine_set_done_hook(e, nullptr, nullptr);
    for (auto& kv : g->table) delete kv.second;
    delete g;
}

RR_API int rr_gateway_set_now(rr_gateway* g, int64_t now_ms) {
    if (!g) return RR_INVALID_ARGUMENT;
    g->manual_now_ms.store(now_ms, std::memory_order_relaxed);
    return RR_OK;
}

static int gw_check(rr_gateway* g, int group, const int32_t* ids, int n_prompt, int max_new) {
    if (group < 0 || group >= g->n_groups) return RR_NO_GROUP;
    if (!ids || n_prompt < 1 || max_new < 1) return RR_INVALID_ARGUMENT;
    if (n_prompt + max_new > g->min_ctx || n_prompt > g->min_pf) return RR_INVALID_ARGUMENT;   // -> HTTP 400, before any debit
    for (int i = 0; i < n_prompt; ++i)
        if (ids[i] < 0 || ids[i] >= g->min_vocab) return RR_INVALID_ARGUMENT;
    return RR_OK;
}

RR_API int rr_gateway_submit_batch(rr_gateway* g, int group, const int32_t* prompt_ids, const int32_t* prompt_start, int n,
                                   int max_new_tokens, int declare_burst, uint64_t* tickets) {
    if (!g || !prompt_ids || !prompt_start || n < 1 || !tickets) return RR_INVALID_ARGUMENT;
    for (int i = 0; i < n; ++i) {
        const int rc = gw_check(g, group, prompt_ids + prompt_start[i], prompt_start[i + 1] - prompt_start[i], max_new_tokens);
        if (rc != RR_OK) return rc;
    }
    std::vector<GwReq*> reqs((size_t)n);
    for (int i = 0; i < n; ++i) {
        GwReq* r = new (std::nothrow) GwReq();
        if (!r) { for (int j = 0; j < i; ++j) delete reqs[j]; return RR_INTERNAL; }
        r->group = group; r->max_new = max_new_tokens;
        r->prompt.assign(prompt_ids + prompt_start[i], prompt_ids + prompt_start[i + 1]);
        reqs[i] = r;
    }
    {
        std::lock_guard<std::mutex> lk(g->mu);
        const double t = gw_now_s(g);
        if (declare_burst) g->pending.push_back(Pending{RR_EV_BURST, nullptr, group, n});
        for (int i = 0; i < n; ++i) {                         // one contiguous block of the trace, in request order
            GwReq* r = reqs[i];
            r->id = g->next_id++;
            r->t_submit = t;
            g->table[r->id] = r;
            g->pending.push_back(Pending{RR_EV_ADMIT, r, 0, 0});
            tickets[i] = r->id;
        }
        g->st.submitted += (uint64_t)n;
    }
    g->cv_work.notify_one();
    return RR_OK;
}

RR_API int rr_gateway_submit(rr_gateway* g, int group, const int32_t* prompt_ids, int n_prompt, int max_new_tokens,
                             uint64_t* ticket) {
    const int32_t start[2] = {0, n_prompt};
    return rr_gateway_submit_batch(g, group, prompt_ids, start, 1, max_new_tokens, 0, ticket);
}

static void gw_fill(const GwReq* r, rr_gateway_result* out) {
    out->ticket = r->id; out->status = r->state == FINISHED ? r->status : -1;
    out->deployment = r->deployment; out->served_group = r->served_group; out->chain_pos = r->chain_pos;
    out->replica = r->replica; out->n_prompt = (int32_t)r->prompt.size(); out->n_generated = (int32_t)r->tokens.size();
    out->attempts = r->attempts;
    out->t_submit_s = r->t_submit; out->t_admit_s = r->t_admit; out->t_first_token_s = r->t_first; out->t_done_s = r->t_done;
}

RR_API int rr_gateway_wait(rr_gateway* g, uint64_t ticket, double timeout_s, rr_gateway_result* out, int32_t* tokens_out,
                           int max_tokens_out) {
    if (!g || !out) return RR_INVALID_ARGUMENT;
    std::unique_lock<std::mutex> lk(g->mu);
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::duration<double>(timeout_s > 0 ? timeout_s : 1e9);
    GwReq* r = nullptr;
    for (;;) {
        auto it = g->table.find(ticket);
        if (it == g->table.end()) return RR_INVALID_ARGUMENT;
        r = it->second;
        if (r->state == FINISHED) break;
        if (g->cv_done.wait_until(lk, deadline) == std::cv_status::timeout) {
            it = g->table.find(ticket);
            if (it == g->table.end()) return RR_INVALID_ARGUMENT;
            r = it->second;
            if (r->state == FINISHED) break;
            gw_fill(r, out);
            out->status = RR_TIMEOUT;
            return RR_TIMEOUT;                                // still in flight: wait again or rr_gateway_cancel
        }
    }
    gw_fill(r, out);
    if (tokens_out) {
        const int n = (int)r->tokens.size() < max_tokens_out ? (int)r->tokens.size() : max_tokens_out;
        if (n > 0) memcpy(tokens_out, r->tokens.data(), sizeof(int32_t) * (size_t)n);
    }
    const int status = r->status;
    g->table.erase(ticket);
    delete r;
    return status;
}

RR_API int rr_gateway_peek(rr_gateway* g, uint64_t ticket, int have, double timeout_s, int32_t* tokens_out, int max_tokens_out,
                           int32_t* n_generated, int32_t* done, rr_gateway_result* out) {
    if (!g || !n_generated || !done) return RR_INVALID_ARGUMENT;
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::duration<double>(timeout_s > 0 ? timeout_s : 1e9);
    for (;;) {
        rr_engine* eng = nullptr;
        uint64_t tk = 0;
        {
            std::unique_lock<std::mutex> lk(g->mu);
            for (;;) {
                auto it = g->table.find(ticket);
                if (it == g->table.end()) return RR_INVALID_ARGUMENT;
                GwReq* r = it->second;
                if (out) gw_fill(r, out);
                if (r->state == FINISHED) {
                    const int n = (int)r->tokens.size();
                    *n_generated = n; *done = 1;
                    if (tokens_out && n > 0) memcpy(tokens_out, r->tokens.data(), sizeof(int32_t) * (size_t)(n < max_tokens_out ? n : max_tokens_out));
                    return RR_OK;
                }
                if (r->state == RUNNING) { eng = r->eng; tk = r->eng_ticket; break; }
                if (g->cv_done.wait_until(lk, deadline) == std::cv_status::timeout) { *n_generated = 0; *done = 0; return RR_OK; }
            }
        }
        const double left = std::chrono::duration<double>(deadline - std::chrono::steady_clock::now()).count();
        int32_t n = 0, d = 0;
        const int rc = rr_engine_peek(eng, tk, have, left > 0.001 ? left : 0.001, tokens_out, max_tokens_out, &n, &d, nullptr);
        if (rc == RR_OK && !d) { *n_generated = n; *done = 0; return RR_OK; }
        // finished (the hook is consuming / has consumed the engine's record) or failed over: look at our own record again
        if (std::chrono::steady_clock::now() >= deadline) { *n_generated = rc == RR_OK ? n : 0; *done = 0; return RR_OK; }
        std::unique_lock<std::mutex> lk(g->mu);
        g->cv_done.wait_for(lk, std::chrono::milliseconds(1));
    }
}

RR_API int rr_gateway_cancel(rr_gateway* g, uint64_t ticket, int count_as_failure) {
    if (!g) return RR_INVALID_ARGUMENT;
    rr_engine* eng = nullptr;
    uint64_t tk = 0;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        auto it = g->table.find(ticket);
        if (it == g->table.end()) return RR_INVALID_ARGUMENT;
        GwReq* r = it->second;
        if (r->state == FINISHED) { g->table.erase(it); delete r; return RR_OK; }
        r->cancelled = true; r->cancel_as_failure = count_as_failure != 0; r->abandoned = true;
        if (r->state == RUNNING) { eng = r->eng; tk = r->eng_ticket; }
        // Q_ADMIT: the dispatcher / done hook sees `cancelled` when it gets to the request
    }
    if (eng) rr_engine_cancel(eng, tk);          // the done hook reports RR_CANCELLED and posts the DONE / FAIL event
    return RR_OK;
}

// Block until every event queued so far (DONE / FAIL reports of finished requests included) has been through K1.
RR_API int rr_gateway_quiesce(rr_gateway* g, double timeout_s) {
    if (!g) return RR_INVALID_ARGUMENT;
    std::unique_lock<std::mutex> lk(g->mu);
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::duration<double>(timeout_s > 0 ? timeout_s : 1e9);
    while (!g->pending.empty() || g->busy) {
        g->cv_work.notify_one();
        if (g->cv_done.wait_until(lk, deadline) == std::cv_status::timeout) return RR_TIMEOUT;
    }
    return RR_OK;
}

RR_API int rr_gateway_get_stats(rr_gateway* g, rr_gateway_stats* out) {
    if (!g || !out) return RR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(g->mu);
    *out = g->st;
    out->in_flight = (uint64_t)g->table.size();
    return RR_OK;
}

RR_API int rr_gateway_trace(rr_gateway* g, rr_event* events, rr_decision* decisions, int capacity, int* n) {
    if (!g || !n) return RR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(g->mu);
    const size_t cap = (size_t)(g->opts.record_trace > 0 ? g->opts.record_trace : 0);
    const size_t have = g->tr_count < cap ? g->tr_count : cap;
    *n = (int)have;
    if (!events || !decisions || (size_t)capacity < have) return have ? RR_INVALID_ARGUMENT : RR_OK;
    const size_t first = g->tr_count - have;
    for (size_t i = 0; i < have; ++i) {
        events[i] = g->tr_ev[(first + i) % cap];
        decisions[i] = g->tr_dec[(first + i) % cap];
    }
    return RR_OK;
}
