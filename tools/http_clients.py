"""Client side of bench.py's HTTP leg, run as a SEPARATE process (its threads must not share the gateway's GIL): n
OpenAI-SDK clients, one request each, issued together -- the reference's client boundary
(reference src/demo_load_balancing.py:24,106-110).  Prints one JSON object."""
import json
import sys
import threading
import time

import openai


def main():
    port, model, n, max_new, n_chars, stream = int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6] == "1"
    start_at = float(sys.argv[7]) if len(sys.argv) > 7 else 0.0          # wall-clock time at which every client process fires
    first_id = int(sys.argv[8]) if len(sys.argv) > 8 else 0
    texts = ["".join(chr(97 + ((first_id + i) * 7 + j * 13) % 26) for j in range(n_chars)) for i in range(n)]
    lat, ttft, errs = [None] * n, [None] * n, []
    barrier = threading.Barrier(n + 1)

    def work(i):
        client = openai.OpenAI(api_key="demo-key", base_url=f"http://127.0.0.1:{port}", max_retries=0)
        try:
            client.models.list()         # warms the SDK (lazy imports, ~0.5 s on a first call) and opens the connection
        except Exception:                # noqa: BLE001
            pass
        barrier.wait()
        t0 = time.perf_counter()
        try:
            if stream:
                first = None
                for _ch in client.chat.completions.create(model=model, messages=[{"role": "user", "content": texts[i]}],
                                                          max_tokens=max_new, stream=True, timeout=600):
                    if first is None:
                        first = time.perf_counter() - t0
                ttft[i] = first
            else:
                r = client.chat.completions.create(model=model, messages=[{"role": "user", "content": texts[i]}],
                                                   max_tokens=max_new, timeout=600)
                assert r.usage.completion_tokens == max_new
            lat[i] = time.perf_counter() - t0
        except Exception as e:                                      # noqa: BLE001
            errs.append(repr(e))
    ths = [threading.Thread(target=work, args=(i,)) for i in range(n)]
    for t in ths:
        t.start()
    d = start_at - time.time()
    if d > 0:
        time.sleep(d)
    barrier.wait()
    t0 = time.perf_counter()
    for t in ths:
        t.join()
    wall = time.perf_counter() - t0
    print(json.dumps({"wall_s": wall, "t_end": time.time(), "lat": [x for x in lat if x is not None], "ttft": [x for x in ttft if x is not None],
                      "errors": errs[:3]}))


if __name__ == "__main__":
    main()
