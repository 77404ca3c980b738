"""CPU: HTTP surface of the gateway (status codes, error bodies, both URL spellings) with a fake router object —
the real router needs the GPU and is covered by tests/test_router_host_gpu.py."""
from starlette.testclient import TestClient

from rr_b200.router import (APIError, BadRequestError, Choice, Message, ModelResponse, RateLimitError, Usage,
                            detokenize, messages_to_text)
from rr_b200.server import create_app


class _Cfg:
    groups = ["g", "limited"]
    deployments = []


class _FakeRouter:
    cfg = _Cfg()

    def completion(self, model, messages, timeout=None, max_tokens=None):
        if model == "limited":
            raise RateLimitError("No deployments available")
        if model == "boom":
            raise APIError("backend failure")
        if model not in self.cfg.groups:
            raise BadRequestError(f"Invalid model name passed in model={model}")
        return ModelResponse("chatcmpl-1", "llama-3-8b@x", [Choice(0, Message("assistant", "hi"))], Usage(3, 1, 4))

    def snapshot(self):
        return []


def test_routes_and_status_codes():
    c = TestClient(create_app(_FakeRouter()))
    body = {"model": "g", "messages": [{"role": "user", "content": "x"}], "timeout": 30}
    for path in ("/chat/completions", "/v1/chat/completions"):
        r = c.post(path, json=body)
        assert r.status_code == 200 and r.json()["model"] == "llama-3-8b@x"
        assert r.json()["choices"][0]["message"] == {"role": "assistant", "content": "hi"}
        assert r.json()["usage"]["total_tokens"] == 4
    r = c.post("/chat/completions", json={"model": "limited", "messages": []})
    assert r.status_code == 429 and r.json()["error"]["type"] == "rate_limit_error"
    assert c.post("/chat/completions", json={"model": "nope", "messages": []}).status_code == 400
    assert c.post("/chat/completions", json={"model": "boom", "messages": []}).status_code == 500
    assert c.post("/chat/completions", json={"messages": []}).status_code == 400
    assert c.post("/chat/completions", content=b"{not json").status_code == 400
    assert c.get("/health").json()["status"] == "ok"
    assert [m["id"] for m in c.get("/v1/models").json()["data"]] == ["g", "limited"]


def test_text_helpers():
    assert messages_to_text([{"role": "user", "content": "hi"}, {"role": "assistant", "content": "yo"}]) == "user: hi\nassistant: yo"
    assert detokenize([1, 3 + ord("o"), 3 + ord("k")]) == "ok"


def test_sse_stream_with_fake_router():
    class R(_FakeRouter):
        def completion_stream(self, model, messages, timeout=None, max_tokens=None):
            if model == "limited":
                raise RateLimitError("No deployments available")
            yield "llama-3-8b@x", [3 + ord("h")], False, 0.01
            yield "llama-3-8b@x", [3 + ord("i")], False, 0.01
            yield "llama-3-8b@x", [], True, 0.01
    c = TestClient(create_app(R()))
    r = c.post("/chat/completions", json={"model": "g", "messages": [{"role": "user", "content": "x"}], "stream": True})
    assert r.status_code == 200 and r.headers["content-type"].startswith("text/event-stream")
    lines = [ln[6:] for ln in r.text.splitlines() if ln.startswith("data: ")]
    assert lines[-1] == "[DONE]"
    import json
    chunks = [json.loads(x) for x in lines[:-1]]
    assert "".join(ch["choices"][0]["delta"].get("content", "") for ch in chunks) == "hi"
    assert chunks[-1]["choices"][0]["finish_reason"] == "length" and chunks[0]["model"] == "llama-3-8b@x"
    assert c.post("/chat/completions", json={"model": "limited", "messages": [], "stream": True}).status_code == 429
