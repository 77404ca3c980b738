"""OpenAI-compatible HTTP front of the router.

Restores the reference's real client boundary: the demos talk to `http://0.0.0.0:<litellm.port>` with the
OpenAI SDK and POST `<base_url>/chat/completions` (reference src/demo_load_balancing.py:24,106-110; the
gateway itself is `litellm --config ./config/config.yaml --port $PORT --num_workers 1`,
reference bin/start-gateway.sh:54).  Rate-limited requests answer HTTP 429 so the SDK raises
`openai.RateLimitError` (reference src/demo_quota_isolation.py:80).

    python -m rr_b200_server --config config/config.yaml            # real replicas (needs the GPUs named in the config)
    python -m rr_b200_server --config config/config.yaml --stub     # mock-completion backends (plumbing only)
"""
import argparse
import time
from typing import Dict, Optional

from .config import RouterConfig, load_config
from .models import resolve_spec
from .router import (APIError, APITimeoutError, BadRequestError, EngineBackend, RateLimitError, Router,
                     StubBackend)


def _error_body(e: APIError, typ: str) -> dict:
    return {"error": {"message": e.message, "type": typ, "param": None, "code": str(e.status_code)}}


def create_app(router: Router):
    from fastapi import FastAPI, Request
    from fastapi.responses import JSONResponse
    from starlette.concurrency import run_in_threadpool

    app = FastAPI(title="rr_b200 gateway")
    t_start = time.time()

    @app.on_event("startup")
    async def _raise_thread_limit():
        # every in-flight request parks one worker thread in rr_gateway_wait (GIL released); the default limit of 40 would
        # cap the concurrency below one replica's 64 decode rows
        import anyio.to_thread
        anyio.to_thread.current_default_thread_limiter().total_tokens = 2048

    async def stream_completion(model, messages, body):
        """Server-sent events in the OpenAI chunk format; the first chunk leaves when the first token exists (TTFT)."""
        import json as _json
        import uuid
        from fastapi.responses import StreamingResponse
        from .router import detokenize
        gen = router.completion_stream(model=model, messages=messages, timeout=body.get("timeout"),
                                       max_tokens=body.get("max_tokens"))
        try:
            first = await run_in_threadpool(next, gen)            # admission errors surface before the 200
        except RateLimitError as e:
            return JSONResponse(_error_body(e, "rate_limit_error"), 429, headers={"retry-after": "1"})
        except BadRequestError as e:
            return JSONResponse(_error_body(e, "invalid_request_error"), 400)
        except APIError as e:
            return JSONResponse(_error_body(e, "api_error"), e.status_code)
        cid = "chatcmpl-" + uuid.uuid4().hex[:24]

        def chunk(label, toks, done):
            delta = {} if done else {"role": "assistant", "content": detokenize(toks)}
            return "data: " + _json.dumps({"id": cid, "object": "chat.completion.chunk", "created": int(time.time()),
                                           "model": label, "choices": [{"index": 0, "delta": delta,
                                                                        "finish_reason": "length" if done else None}]}) + "\n\n"

        def events():
            try:
                label, toks, done, _ = first
                yield chunk(label, toks, done)
                if not done:
                    for label, toks, done, _ in gen:
                        yield chunk(label, toks, done)
                yield "data: [DONE]\n\n"
            finally:
                gen.close()        # client gone before the last chunk: the router cancels the request and frees its decode row

        return StreamingResponse(events(), media_type="text/event-stream")

    async def chat_completions(request: Request):
        try:
            body = await request.json()
        except Exception:
            return JSONResponse(_error_body(BadRequestError("invalid JSON body"), "invalid_request_error"), 400)
        model, messages = body.get("model"), body.get("messages")
        if not isinstance(model, str) or not isinstance(messages, list):
            return JSONResponse(_error_body(BadRequestError("`model` and `messages` are required"),
                                            "invalid_request_error"), 400)
        if body.get("stream"):
            return await stream_completion(model, messages, body)
        try:
            resp = await run_in_threadpool(router.completion, model=model, messages=messages,
                                           timeout=body.get("timeout"), max_tokens=body.get("max_tokens"))
            return JSONResponse(resp.model_dump())
        except RateLimitError as e:
            return JSONResponse(_error_body(e, "rate_limit_error"), 429, headers={"retry-after": "1"})
        except BadRequestError as e:
            return JSONResponse(_error_body(e, "invalid_request_error"), 400)
        except APITimeoutError as e:
            return JSONResponse(_error_body(e, "timeout"), 408)
        except APIError as e:
            return JSONResponse(_error_body(e, "api_error"), e.status_code)

    # the SDK appends /chat/completions to base_url; both spellings are served
    app.add_api_route("/chat/completions", chat_completions, methods=["POST"])
    app.add_api_route("/v1/chat/completions", chat_completions, methods=["POST"])

    @app.get("/health")
    async def health():
        return {"status": "ok", "uptime_s": time.time() - t_start}

    async def models():
        return {"object": "list", "data": [{"id": g, "object": "model", "owned_by": "rr_b200"}
                                           for g in router.cfg.groups]}

    app.add_api_route("/v1/models", models, methods=["GET"])
    app.add_api_route("/models", models, methods=["GET"])          # the SDK appends /models to a base_url without /v1

    @app.get("/router/state")
    async def state():
        snap = await run_in_threadpool(router.snapshot)
        return {"deployments": [dict(model_name=d.model_name, model=d.response_model, gpu=d.gpu, rpm=d.rpm,
                                     tpm=d.tpm, **s) for d, s in zip(router.cfg.deployments, snap)]}

    return app


def build_backends(cfg: RouterConfig, stub: bool, spec_override: Optional[str] = None, max_batch: int = 64,
                   ctx_max: int = 1024) -> Dict[int, object]:
    """One backend per replica (GPU) named in the config; deployments on the same GPU share it."""
    backends: Dict[int, object] = {}
    for d in cfg.deployments:
        if d.gpu in backends:
            continue
        if stub:
            backends[d.gpu] = StubBackend()
            continue
        from .engine import Engine
        spec = resolve_spec(spec_override or d.model)
        backends[d.gpu] = EngineBackend(Engine.synthetic(spec, seed=0, device=d.gpu, max_batch=max_batch,
                                                         ctx_max=ctx_max))
    return backends


def main(argv=None):
    ap = argparse.ArgumentParser(description="rr_b200 OpenAI-compatible gateway")
    ap.add_argument("--config", default="./config/config.yaml")
    ap.add_argument("--port", type=int, default=None)
    ap.add_argument("--host", default="0.0.0.0")
    ap.add_argument("--stub", action="store_true", help="mock-completion backends instead of model replicas")
    ap.add_argument("--spec", default=None, help="override the model spec of every deployment (e.g. tiny)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--max-tokens", type=int, default=64)
    a = ap.parse_args(argv)
    import uvicorn
    cfg = load_config(a.config)
    router = Router(config=cfg, backends=build_backends(cfg, a.stub, a.spec), seed=a.seed,
                    default_max_tokens=a.max_tokens)
    print(f"rr_b200 gateway: {len(cfg.deployments)} deployments in {len(cfg.groups)} model groups, "
          f"routing_strategy={cfg.routing_strategy}, port {a.port or cfg.port}")
    uvicorn.run(create_app(router), host=a.host, port=a.port or cfg.port, workers=1, log_level="warning")


if __name__ == "__main__":
    main()
