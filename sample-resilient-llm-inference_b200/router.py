"""Host-side mirror of the reference's router interface.

`Router(model_list=..., routing_strategy=..., enable_pre_call_checks=..., allowed_fails=...,
cooldown_time=..., fallbacks=...)` takes the same arguments the reference's gateway builds
`litellm.Router` from (reference config/config.yaml:35-108, launched by bin/start-gateway.sh:54);
`Router.completion(model=, messages=, timeout=)` / `client.chat.completions.create(...)` is the call
the demos make (reference src/demo_load_balancing.py:106-110, src/demo_fallback.py:143-147,
src/demo_quota_isolation.py:52-56), the result exposes `.model` the way they read it (:116), and a
rate-limited request raises `RateLimitError` (HTTP 429; src/demo_quota_isolation.py:80).

All routing state and decisions live in the CUDA library (K1, rr_router.cu): this file builds event
records, calls `rr_router_process`, and moves token ids to and from the per-GPU engines.  There is no
Python/CPU implementation of the routing logic on this path.
"""
from __future__ import annotations

import ctypes as C
import threading
import time
import uuid
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .config import RouterConfig, build_config, load_config

EV_ADMIT, EV_DONE, EV_FAIL, EV_BURST = 0, 1, 2, 3


# ---------------------------------------------------------------- errors (OpenAI-SDK names)
class APIError(Exception):
    status_code = 500

    def __init__(self, message: str, status_code: Optional[int] = None):
        super().__init__(message)
        self.message = message
        if status_code is not None:
            self.status_code = status_code


class RateLimitError(APIError):
    status_code = 429


class BadRequestError(APIError):
    status_code = 400


class APITimeoutError(APIError):
    status_code = 408


class ServiceUnavailableError(APIError):
    status_code = 503


# ---------------------------------------------------------------- response objects
@dataclass
class Message:
    role: str
    content: str


@dataclass
class Choice:
    index: int
    message: Message
    finish_reason: str = "length"


@dataclass
class Usage:
    prompt_tokens: int
    completion_tokens: int
    total_tokens: int


@dataclass
class ModelResponse:
    id: str
    model: str
    choices: List[Choice]
    usage: Usage
    created: int = 0
    object: str = "chat.completion"
    # extras (underscore = not part of the OpenAI shape)
    _token_ids: List[int] = field(default_factory=list)
    _deployment: int = -1
    _model_group: str = ""
    _fell_back: bool = False
    _ttft_s: float = 0.0
    _latency_s: float = 0.0

    def model_dump(self) -> dict:
        return {"id": self.id, "object": self.object, "created": self.created, "model": self.model,
                "choices": [{"index": c.index, "finish_reason": c.finish_reason,
                             "message": {"role": c.message.role, "content": c.message.content}}
                            for c in self.choices],
                "usage": {"prompt_tokens": self.usage.prompt_tokens,
                          "completion_tokens": self.usage.completion_tokens,
                          "total_tokens": self.usage.total_tokens}}


# ---------------------------------------------------------------- tokenizer (byte level, CUDA kernel K2 in the library)
def tokenize_batch(texts: Sequence[str], vocab: int):
    """Token ids and counts of a batch of messages in ONE launch of the tokenizer kernel (csrc/rr_tokenizer.cu).
    -> (list of int32 arrays, counts int32[n])."""
    raw = [t.encode("utf-8") for t in texts]
    n = len(raw)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(b) for b in raw], out=off[1:])
    blob = b"".join(raw)
    counts = np.zeros(n, dtype=np.int32)
    ids = np.zeros(int(off[-1]) + n, dtype=np.int32)
    ids_off = np.zeros(n + 1, dtype=np.int64)
    _lib.check(_lib.lib.rr_tokenize_batch(blob, off.ctypes.data_as(C.POINTER(C.c_int64)), n, vocab,
                                          counts.ctypes.data_as(C.POINTER(C.c_int32)), ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                          len(ids), ids_off.ctypes.data_as(C.POINTER(C.c_int64))), "rr_tokenize_batch")
    return [ids[ids_off[i]: ids_off[i + 1]] for i in range(n)], counts


def tokenize(text: str, vocab: int) -> np.ndarray:
    return tokenize_batch([text], vocab)[0][0]


def count_tokens(text: str) -> int:
    return int(tokenize_batch([text], 259)[1][0])


def tokenize_host(text: str, vocab: int) -> np.ndarray:
    """Host form of the same tokenizer (rr_tokenize): what the device kernel is checked against."""
    b = text.encode("utf-8")
    ids = np.zeros(len(b) + 1, dtype=np.int32)
    n = C.c_int32()
    _lib.check(_lib.lib.rr_tokenize(b, len(b), vocab, ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                    len(ids), C.byref(n)), "rr_tokenize")
    return ids[: n.value]


def detokenize(ids: Sequence[int]) -> str:
    return bytes(((int(t) - 3) % 256) for t in ids if int(t) >= 3).decode("utf-8", errors="replace")


def messages_to_text(messages: Sequence[dict]) -> str:
    return "\n".join(f"{m.get('role', 'user')}: {m.get('content', '')}" for m in messages)


# ---------------------------------------------------------------- backends
class StubBackend:
    """Mock-completion backend (BASELINE.json configs[0]: plumbing only): a fixed reply of
    `reply_tokens` token ids, optional seeded failures."""

    def __init__(self, vocab: int = 128256, reply_tokens: int = 8, fail_every: int = 0):
        self.vocab = vocab
        self.reply = reply_tokens
        self.fail_every = fail_every
        self._n = 0
        self._lock = threading.Lock()

    def submit(self, prompt_ids, max_new):
        with self._lock:
            self._n += 1
            failed = self.fail_every > 0 and self._n % self.fail_every == 0
        t = time.perf_counter()
        return (failed, min(max_new, self.reply), t)

    def wait(self, handle, timeout):
        failed, n, t = handle
        now = time.perf_counter()
        return (7 if failed else 0), ([] if failed else [3 + (i % 200) for i in range(n)]), now - t, now - t

    def peek(self, handle, have, timeout=0.05):
        failed, n, t = handle
        return ([] if failed else [3 + (i % 200) for i in range(n)]), True, time.perf_counter() - t


class EngineBackend:
    def __init__(self, engine):
        self.engine = engine
        self.vocab = engine.spec.vocab

    def peek(self, handle, have, timeout=0.05):
        return self.engine.peek(handle, have, timeout=timeout)

    def submit(self, prompt_ids, max_new):
        return self.engine.submit(prompt_ids, max_new)

    def wait(self, handle, timeout):
        rec = self.engine.wait(handle, timeout=timeout or 0.0)
        return rec.status, rec.tokens, rec.ttft, rec.latency


# ---------------------------------------------------------------- the router
class Router:
    def __init__(self, model_list: Optional[List[dict]] = None, *, routing_strategy: str = "simple-shuffle",
                 enable_pre_call_checks: bool = False, allowed_fails: Optional[int] = None,
                 cooldown_time: Optional[float] = None, fallbacks: Optional[List[dict]] = None,
                 config: Optional[RouterConfig] = None, backends: Optional[Dict[int, Any]] = None,
                 seed: int = 0, device: int = 0, clock: Optional[Callable[[], float]] = None,
                 default_max_tokens: int = 128):
        if config is None:
            rs = {"routing_strategy": routing_strategy, "enable_pre_call_checks": enable_pre_call_checks,
                  "fallbacks": fallbacks or []}
            if allowed_fails is not None:
                rs["allowed_fails"] = allowed_fails
            if cooldown_time is not None:
                rs["cooldown_time"] = cooldown_time
            config = build_config(model_list or [], rs)
        self.cfg = config
        self.backends: Dict[int, Any] = dict(backends or {})
        self._custom_clock = clock is not None
        self.clock = clock or time.time
        self._gw = None                 # rr_gateway handle (native per-request path), created on first use
        self._gw_sig = None
        self._gw_lock = threading.Lock()
        self.record_trace = 0           # > 0: the gateway keeps the last N (event, decision) pairs (gateway_trace())
        self.default_max_tokens = default_max_tokens
        deps = config.deployments
        if not deps:
            raise ValueError("model_list is empty")
        arr = (_lib.DeploymentDesc * len(deps))()
        for i, d in enumerate(deps):
            arr[i] = _lib.DeploymentDesc(d.group, d.rpm, d.tpm, d.weight, d.gpu, 0)
        ng = len(config.groups)
        offs, flat = [0], []
        for g in range(ng):
            flat += config.fallbacks.get(g, [])
            offs.append(len(flat))
        st = _lib.RouterSettings(config.strategy_id, int(config.enable_pre_call_checks), config.allowed_fails,
                                 int(round(config.cooldown_time * 1000)), 0)
        self._h = C.c_void_p()
        _lib.check(_lib.lib.rr_router_create(arr, len(deps), ng, (C.c_int32 * len(offs))(*offs),
                                             (C.c_int32 * max(1, len(flat)))(*flat), C.byref(st), seed, device,
                                             C.byref(self._h)), "rr_router_create")

    @classmethod
    def from_config(cls, path: str, **kw) -> "Router":
        return cls(config=load_config(path), **kw)

    def close(self):
        if getattr(self, "_gw", None):
            _lib.lib.rr_gateway_destroy(self._gw)
            self._gw = None
        if getattr(self, "_h", None):
            _lib.lib.rr_router_destroy(self._h)
            self._h = None

    # ---- native per-request path (rr_gateway.cu) ------------------------------------------------------
    def _gateway(self):
        """rr_gateway over this router and its engines, or None when a backend is not a native engine (StubBackend:
        BASELINE config #1, plumbing only -- the host loop below serves those)."""
        if not self.backends or not all(isinstance(b, EngineBackend) for b in self.backends.values()):
            return None
        sig = tuple(sorted((k, id(b.engine)) for k, b in self.backends.items()))
        with self._gw_lock:
            if self._gw is not None and self._gw_sig == sig:
                return self._gw
            if self._gw is not None:
                _lib.lib.rr_gateway_destroy(self._gw)
                self._gw = None
            keys = sorted(self.backends)
            engs = (C.c_void_p * len(keys))(*[self.backends[k].engine._h for k in keys])
            reps = (C.c_int32 * len(keys))(*keys)
            opts = _lib.GatewayOpts(1 if self._custom_clock else 0, int(self.record_trace), 0)
            h = C.c_void_p()
            _lib.check(_lib.lib.rr_gateway_create(self._h, engs, reps, len(keys), C.byref(opts), C.byref(h)),
                       "rr_gateway_create")
            self._gw, self._gw_sig = h, sig
            return h

    def gateway_stats(self) -> dict:
        gw = self._gateway()
        if gw is None:
            return {}
        s = _lib.GatewayStats()
        _lib.check(_lib.lib.rr_gateway_get_stats(gw, C.byref(s)), "rr_gateway_get_stats")
        return {k: getattr(s, k) for k, _ in s._fields_}

    def gateway_trace(self) -> List[Tuple[Tuple[int, int, int, int, int], Tuple[int, int, int, int]]]:
        """[(event, decision)] recorded by the gateway (record_trace > 0), oldest first: the serialised trace K1 processed."""
        gw = self._gateway()
        n = C.c_int32()
        if gw is None or _lib.lib.rr_gateway_trace(gw, None, None, 0, C.byref(n)) not in (0, 4) or n.value == 0:
            return []
        ev, dec = (_lib.Event * n.value)(), (_lib.Decision * n.value)()
        _lib.check(_lib.lib.rr_gateway_trace(gw, ev, dec, n.value, C.byref(n)), "rr_gateway_trace")
        return [((e.type, e.target, e.tokens, e.chain_start, e.now_ms), (d.status, d.deployment, d.served_group, d.chain_pos))
                for e, d in zip(ev[: n.value], dec[: n.value])]

    def _gw_response(self, res, toks, n_prompt, t_start, detok=True) -> ModelResponse:
        d = self.cfg.deployments[res.deployment]
        toks = [int(t) for t in toks[: res.n_generated]]
        return ModelResponse(
            id="chatcmpl-" + uuid.uuid4().hex[:24], model=d.response_model,
            choices=[Choice(0, Message("assistant", detokenize(toks) if detok else ""))],
            usage=Usage(n_prompt, len(toks), n_prompt + len(toks)), created=int(time.time()),
            _token_ids=toks, _deployment=res.deployment, _model_group=self.cfg.groups[res.served_group],
            _fell_back=res.chain_pos > 0, _ttft_s=res.t_first_token_s - res.t_submit_s,
            _latency_s=time.perf_counter() - t_start)

    def _gw_error(self, status: int, model: str, res=None, timeout=None) -> APIError:
        if status == 1:
            return RateLimitError(f"No deployments available for selected model, passed model={model} (rate limited)")
        if status == 2:
            return BadRequestError(f"Invalid model name passed in model={model}")
        if status == 4:
            return BadRequestError("prompt + max_tokens exceed the context window of the deployment, or a token id is "
                                   "outside its vocabulary")
        if status == 6:
            return APITimeoutError(f"Request timed out after {timeout}s")
        if status == 7:
            dep = res.deployment if res is not None else -1
            return APIError(f"backend failure on deployment {dep} and no fallback available for model={model}")
        return APIError(f"{_lib.lib.rr_strerror(status).decode()} (model={model})")

    def _gw_submit(self, gw, g: int, prompt_ids, max_new: int, model: str) -> int:
        ids = np.ascontiguousarray(np.asarray(prompt_ids, dtype=np.int32))
        if self._custom_clock:
            _lib.lib.rr_gateway_set_now(gw, self.now_ms())
        t = C.c_uint64()
        rc = _lib.lib.rr_gateway_submit(gw, g, ids.ctypes.data_as(C.POINTER(C.c_int32)), len(ids), max_new, C.byref(t))
        if rc != 0:
            raise self._gw_error(rc, model)
        return t.value

    def _completion_gw(self, gw, model, g, prompt_ids, max_new, timeout):
        t_start = time.perf_counter()
        ticket = self._gw_submit(gw, g, prompt_ids, max_new, model)
        res = _lib.GatewayResult()
        toks = np.zeros(max_new, dtype=np.int32)
        rc = _lib.lib.rr_gateway_wait(gw, ticket, float(timeout or 0.0), C.byref(res),
                                      toks.ctypes.data_as(C.POINTER(C.c_int32)), max_new)
        if rc == 6:                      # client-side timeout: give the row back, count a failure against the deployment
            _lib.lib.rr_gateway_cancel(gw, ticket, 1)
            raise self._gw_error(6, model, res, timeout)
        if rc != 0:
            raise self._gw_error(rc, model, res, timeout)
        return self._gw_response(res, toks, len(prompt_ids), t_start)

    def _stream_gw(self, gw, model, g, prompt_ids, max_new, timeout):
        ticket = self._gw_submit(gw, g, prompt_ids, max_new, model)
        deadline = None if not timeout else time.perf_counter() + timeout
        res = _lib.GatewayResult()
        toks = np.zeros(max_new, dtype=np.int32)
        n, done = C.c_int32(), C.c_int32()
        sent, finished = 0, False
        try:
            while True:
                _lib.check(_lib.lib.rr_gateway_peek(gw, ticket, sent, 0.05, toks.ctypes.data_as(C.POINTER(C.c_int32)), max_new,
                                                    C.byref(n), C.byref(done), C.byref(res)), "rr_gateway_peek")
                if done.value:
                    rc = _lib.lib.rr_gateway_wait(gw, ticket, 1.0, C.byref(res), toks.ctypes.data_as(C.POINTER(C.c_int32)), max_new)
                    finished = True
                    if rc != 0:
                        raise self._gw_error(rc, model, res, timeout)
                    label = self.cfg.deployments[res.deployment].response_model
                    ttft = res.t_first_token_s - res.t_submit_s
                    if res.n_generated > sent:
                        yield label, [int(t) for t in toks[sent: res.n_generated]], False, ttft
                    yield label, [], True, ttft
                    return
                if n.value > sent and res.deployment >= 0:
                    label = self.cfg.deployments[res.deployment].response_model
                    yield label, [int(t) for t in toks[sent: n.value]], False, 0.0
                    sent = n.value
                if deadline is not None and time.perf_counter() > deadline:
                    _lib.lib.rr_gateway_cancel(gw, ticket, 1)
                    finished = True
                    raise self._gw_error(6, model, res, timeout)
        finally:
            if not finished:             # the consumer went away (SSE client disconnected): free the row, no failure counted
                _lib.lib.rr_gateway_cancel(gw, ticket, 0)

    def _batch_gw(self, gw, model, g, prompts, max_new, timeout):
        t0 = time.perf_counter()
        n = len(prompts)
        ids = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.int32) for p in prompts]))
        start = np.zeros(n + 1, dtype=np.int32)
        np.cumsum([len(p) for p in prompts], out=start[1:])
        if self._custom_clock:
            _lib.lib.rr_gateway_set_now(gw, self.now_ms())
        tk = (C.c_uint64 * n)()
        rc = _lib.lib.rr_gateway_submit_batch(gw, g, ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                              start.ctypes.data_as(C.POINTER(C.c_int32)), n, max_new,
                                              1 if self.cfg.routing_strategy == "split" else 0, tk)
        if rc != 0:
            raise self._gw_error(rc, model)
        out: List[Any] = [None] * n
        toks = np.zeros(max_new, dtype=np.int32)
        for i in range(n):
            res = _lib.GatewayResult()
            left = None if not timeout else max(0.001, timeout - (time.perf_counter() - t0))
            rc = _lib.lib.rr_gateway_wait(gw, tk[i], float(left or 0.0), C.byref(res),
                                          toks.ctypes.data_as(C.POINTER(C.c_int32)), max_new)
            if rc == 6:
                _lib.lib.rr_gateway_cancel(gw, tk[i], 1)
            out[i] = self._gw_response(res, toks, len(prompts[i]), t0, detok=False) if rc == 0 \
                else self._gw_error(rc, model, res, timeout)
        return out

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- K1 calls -------------------------------------------------------------------------------
    def now_ms(self) -> int:
        return int(self.clock() * 1000)

    def process(self, events: Sequence[Tuple[int, int, int, int, int]]) -> List[Tuple[int, int, int, int]]:
        """events: (type, target, tokens, chain_start, now_ms) -> (status, deployment, group, chain_pos)."""
        n = len(events)
        if n == 0:
            return []
        self._quiesce()
        ev = (_lib.Event * n)()
        for i, e in enumerate(events):
            ev[i] = _lib.Event(*e)
        out = (_lib.Decision * n)()
        _lib.check(_lib.lib.rr_router_process(self._h, ev, n, out), "rr_router_process")
        return [(d.status, d.deployment, d.served_group, d.chain_pos) for d in out]

    def _quiesce(self):
        """Let the gateway's dispatcher post what is still queued (DONE / FAIL reports of finished requests)."""
        if self._gw is not None:
            _lib.lib.rr_gateway_quiesce(self._gw, 5.0)

    def snapshot(self) -> List[dict]:
        self._quiesce()
        n = len(self.cfg.deployments)
        s = (_lib.DeploymentState * n)()
        _lib.check(_lib.lib.rr_router_snapshot(self._h, s), "rr_router_snapshot")
        return [{k: getattr(x, k) for k, _ in x._fields_} for x in s]

    # ---- request path ---------------------------------------------------------------------------
    def _backend_for(self, dep_index: int):
        d = self.cfg.deployments[dep_index]
        b = self.backends.get(d.gpu)
        if b is None:
            raise ServiceUnavailableError(f"no backend attached for replica {d.gpu} ({d.model})")
        return b

    def _vocab(self) -> int:
        """Token ids must be valid on every replica a request may fall back to: the smallest vocabulary attached."""
        return min((b.vocab for b in self.backends.values()), default=128256)

    def completion(self, model: str, messages: Optional[Sequence[dict]] = None, timeout: Optional[float] = None,
                   max_tokens: Optional[int] = None, prompt_ids: Optional[Sequence[int]] = None,
                   **_ignored) -> ModelResponse:
        g = self.cfg.group_index(model)
        if g < 0:
            raise BadRequestError(f"Invalid model name passed in model={model}")
        if prompt_ids is None:
            prompt_ids = tokenize(messages_to_text(messages or []), self._vocab())
        n_prompt = len(prompt_ids)
        max_new = max_tokens or self.default_max_tokens
        gw = self._gateway()
        if gw is not None:
            return self._completion_gw(gw, model, g, prompt_ids, max_new, timeout)
        # ---- host loop (StubBackend): same admission events, one K1 launch per event
        t_start = time.perf_counter()
        chain_start = 0
        last_err: Optional[APIError] = None
        while True:
            (status, dep, sg, pos), = self.process([(EV_ADMIT, g, n_prompt, chain_start, self.now_ms())])
            if status == 1:
                raise last_err or RateLimitError(
                    f"No deployments available for selected model, passed model={model} (rate limited)")
            if status != 0:
                raise BadRequestError(f"Invalid model name passed in model={model}")
            remaining = None if timeout is None else max(0.001, timeout - (time.perf_counter() - t_start))
            try:
                backend = self._backend_for(dep)
                st, toks, ttft, lat = backend.wait(backend.submit(prompt_ids, max_new), remaining)
            except Exception as exc:     # the request was admitted: never leave its in-flight slot behind
                self.process([(EV_FAIL, dep, 0, 0, self.now_ms())])
                raise exc if isinstance(exc, APIError) else APIError(f"backend error on deployment {dep}: {exc}")
            if st == 0:
                self.process([(EV_DONE, dep, len(toks), 0, self.now_ms())])
                d = self.cfg.deployments[dep]
                return ModelResponse(
                    id="chatcmpl-" + uuid.uuid4().hex[:24], model=d.response_model,
                    choices=[Choice(0, Message("assistant", detokenize(toks)))],
                    usage=Usage(n_prompt, len(toks), n_prompt + len(toks)), created=int(time.time()),
                    _token_ids=list(toks), _deployment=dep, _model_group=self.cfg.groups[sg],
                    _fell_back=pos > 0, _ttft_s=ttft, _latency_s=time.perf_counter() - t_start)
            # backend failed / timed out: report, then walk on down the fallback chain
            self.process([(EV_FAIL, dep, 0, 0, self.now_ms())])
            if st == 6:
                raise APITimeoutError(f"Request timed out after {timeout}s")
            last_err = APIError(f"backend failure on deployment {dep} ({self.cfg.deployments[dep].model})")
            chain_start = pos + 1

    def completion_stream(self, model: str, messages: Optional[Sequence[dict]] = None, timeout: Optional[float] = None,
                          max_tokens: Optional[int] = None, prompt_ids: Optional[Sequence[int]] = None):
        """Streaming variant (SSE in server.py): admission exactly like completion(); yields
        (model_label, new_token_ids, done, ttft_s) as tokens arrive.  A backend failure before the first token walks the
        fallback chain; rate limiting raises RateLimitError before anything is yielded."""
        g = self.cfg.group_index(model)
        if g < 0:
            raise BadRequestError(f"Invalid model name passed in model={model}")
        if prompt_ids is None:
            prompt_ids = tokenize(messages_to_text(messages or []), self._vocab())
        max_new = max_tokens or self.default_max_tokens
        gw = self._gateway()
        if gw is not None:
            yield from self._stream_gw(gw, model, g, prompt_ids, max_new, timeout)
            return
        deadline = None if not timeout else time.perf_counter() + timeout
        chain_start = 0
        while True:
            (status, dep, sg, pos), = self.process([(EV_ADMIT, g, len(prompt_ids), chain_start, self.now_ms())])
            if status == 1:
                raise RateLimitError(f"No deployments available for selected model, passed model={model} (rate limited)")
            if status != 0:
                raise BadRequestError(f"Invalid model name passed in model={model}")
            sent, failed, settled = 0, False, False
            try:
                backend = self._backend_for(dep)
                label = self.cfg.deployments[dep].response_model
                h = backend.submit(prompt_ids, max_new)
                while True:
                    toks, done, ttft = backend.peek(h, sent)
                    if len(toks) > sent:
                        yield label, toks[sent:], False, ttft
                        sent = len(toks)
                    if done:
                        st, toks, ttft, _lat = backend.wait(h, timeout)
                        failed = st != 0
                        break
                    if deadline is not None and time.perf_counter() > deadline:
                        raise APITimeoutError(f"Request timed out after {timeout}s")
                if not failed:
                    self.process([(EV_DONE, dep, sent, 0, self.now_ms())])
                    settled = True
                    yield label, [], True, ttft
                    return
            finally:
                if not settled:          # failure, timeout, or the consumer closed the generator: release the slot
                    self.process([(EV_FAIL if failed or sent == 0 else EV_DONE, dep, sent, 0, self.now_ms())])
            if sent > 0:
                raise APIError(f"backend failure on deployment {dep} after {sent} tokens")
            chain_start = pos + 1

    def completion_batch(self, model: str, prompts: Sequence[Sequence[int]], max_tokens: int,
                         timeout: Optional[float] = None) -> List[Any]:
        """Closed burst of requests to one model group (what the demos do with N threads,
        reference src/demo_load_balancing.py:195-203): ONE admission launch over the whole trace,
        dispatch to the per-GPU engines, one DONE launch.  -> ModelResponse or exception per request."""
        g = self.cfg.group_index(model)
        if g < 0:
            raise BadRequestError(f"Invalid model name passed in model={model}")
        gw = self._gateway()
        if gw is not None:
            return self._batch_gw(gw, model, g, prompts, max_tokens, timeout)
        now = self.now_ms()
        head = [(EV_BURST, g, len(prompts), 0, now)] if self.cfg.routing_strategy == "split" else []
        dec = self.process(head + [(EV_ADMIT, g, len(p), 0, now) for p in prompts])[len(head):]
        t0 = time.perf_counter()
        handles: List[Any] = [None] * len(prompts)
        out: List[Any] = [None] * len(prompts)
        for i, (status, dep, sg, pos) in enumerate(dec):
            if status != 0:
                out[i] = RateLimitError(f"No deployments available for model={model} (rate limited)")
                continue
            try:
                b = self._backend_for(dep)
                handles[i] = (b, b.submit(prompts[i], max_tokens))
            except Exception as exc:     # admitted but never started: release the slot, keep serving the others
                self.process([(EV_FAIL, dep, 0, 0, self.now_ms())])
                out[i] = exc if isinstance(exc, APIError) else APIError(f"backend error on deployment {dep}: {exc}")
        post = []
        retry = []
        for i, h in enumerate(handles):
            if h is None:
                continue
            b, hd = h
            status, dep, sg, pos = dec[i]
            st, toks, ttft, lat = b.wait(hd, timeout)
            if st == 0:
                post.append((EV_DONE, dep, len(toks), 0, self.now_ms()))
                d = self.cfg.deployments[dep]
                out[i] = ModelResponse(
                    id="chatcmpl-" + uuid.uuid4().hex[:24], model=d.response_model,
                    choices=[Choice(0, Message("assistant", ""))],
                    usage=Usage(len(prompts[i]), len(toks), len(prompts[i]) + len(toks)), created=int(time.time()),
                    _token_ids=list(toks), _deployment=dep, _model_group=self.cfg.groups[sg], _fell_back=pos > 0,
                    _ttft_s=ttft, _latency_s=time.perf_counter() - t0)
            else:
                post.append((EV_FAIL, dep, 0, 0, self.now_ms()))
                retry.append((i, pos + 1))
        self.process(post)
        for i, chain_start in retry:          # failed requests walk the fallback chain one by one
            try:
                out[i] = self._retry_from(model, g, prompts[i], max_tokens, chain_start, timeout)
            except APIError as e:
                out[i] = e
        return out

    def _retry_from(self, model, g, prompt_ids, max_new, chain_start, timeout):
        t_start = time.perf_counter()
        while True:
            (status, dep, sg, pos), = self.process([(EV_ADMIT, g, len(prompt_ids), chain_start, self.now_ms())])
            if status != 0:
                raise APIError(f"backend failure and no fallback available for model={model}")
            b = self._backend_for(dep)
            st, toks, ttft, lat = b.wait(b.submit(prompt_ids, max_new), timeout)
            if st == 0:
                self.process([(EV_DONE, dep, len(toks), 0, self.now_ms())])
                d = self.cfg.deployments[dep]
                return ModelResponse(
                    id="chatcmpl-" + uuid.uuid4().hex[:24], model=d.response_model,
                    choices=[Choice(0, Message("assistant", detokenize(toks)))],
                    usage=Usage(len(prompt_ids), len(toks), len(prompt_ids) + len(toks)), created=int(time.time()),
                    _token_ids=list(toks), _deployment=dep, _model_group=self.cfg.groups[sg], _fell_back=pos > 0,
                    _ttft_s=ttft, _latency_s=time.perf_counter() - t_start)
            self.process([(EV_FAIL, dep, 0, 0, self.now_ms())])
            chain_start = pos + 1


# ---------------------------------------------------------------- OpenAI-SDK-shaped facade
class _Completions:
    def __init__(self, router: Router):
        self._r = router

    def create(self, model: str, messages: Sequence[dict], timeout: Optional[float] = None,
               max_tokens: Optional[int] = None, **kw) -> ModelResponse:
        return self._r.completion(model=model, messages=messages, timeout=timeout, max_tokens=max_tokens, **kw)


class _Chat:
    def __init__(self, router: Router):
        self.completions = _Completions(router)


class OpenAI:
    """In-process stand-in for `openai.OpenAI(api_key=..., base_url=http://0.0.0.0:<port>)`
    (reference src/demo_load_balancing.py:24): same `client.chat.completions.create(...)` call."""

    def __init__(self, router: Router, api_key: str = "demo-key", base_url: Optional[str] = None):
        self.api_key = api_key
        self.base_url = base_url
        self.chat = _Chat(router)
