"""HTTP front end with request batching (the reference's server.py, srv:1-38).

Same wire contract as the reference: `POST /v1/completions` with `{"prompt": str | [str, ...], "max_tokens": int}` answers
`{"model": "phi-3-vision", "responses": [str, ...]}`; any other path is a 404 (srv:9-30). The reference serves one request at a
time through a single-threaded HTTPServer and calls `generate(prompts, preload=preload)` once per request. Here connections
are accepted concurrently and a single worker drains a queue: every request that arrives while the GPU is busy (or within
`window_ms` of the first waiting one) joins the next batched decode, up to `max_batch` prompts, grouped by `max_tokens` and by
the side of the LongRoPE switch they fall on (rows of one batch must agree on both, api.generate_batch). One batched call
costs about the same as a single prompt (the decode step is a weight stream), so N concurrent clients cost ~1 generate call.

    python -m phi3_b200.server --port 8000 [--max-batch 16] [--window-ms 2]
"""
import argparse
import json
import queue
import threading
import time
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer


class _Job:
    __slots__ = ('prompts', 'max_tokens', 'done', 'responses', 'error')

    def __init__(self, prompts, max_tokens):
        self.prompts, self.max_tokens = prompts, max_tokens
        self.done = threading.Event()
        self.responses, self.error = None, None


class Batcher:
    """Coalesces queued jobs into batched `generate_fn(prompts, max_tokens) -> [str]` calls on one worker thread."""

    def __init__(self, generate_fn, max_batch=16, window_ms=2.0, group_key=None):
        self.generate_fn, self.max_batch, self.window = generate_fn, max_batch, window_ms / 1e3
        self.group_key = group_key or (lambda job: job.max_tokens)
        self.q = queue.Queue()
        self.calls = 0                                       # batched calls issued (tests / metrics)
        self._stop = False
        self.worker = threading.Thread(target=self._run, daemon=True)
        self.worker.start()

    def submit(self, prompts, max_tokens):
        job = _Job(list(prompts), int(max_tokens))
        self.q.put(job)
        job.done.wait()
        if job.error is not None:
            raise job.error
        return job.responses

    def close(self):
        self._stop = True
        self.q.put(None)
        self.worker.join(timeout=5)

    def _take(self):
        """first job (blocking) + whatever else shows up inside the window"""
        first = self.q.get()
        if first is None:
            return None
        jobs, n = [first], len(first.prompts)
        deadline = time.monotonic() + self.window
        while n < self.max_batch:
            try:
                nxt = self.q.get(timeout=max(0.0, deadline - time.monotonic()))
            except queue.Empty:
                break
            if nxt is None:
                self._stop = True
                break
            jobs.append(nxt)
            n += len(nxt.prompts)
        return jobs

    def _run(self):
        while not self._stop:
            jobs = self._take()
            if jobs is None:
                return
            groups = {}
            for j in jobs:
                groups.setdefault(self.group_key(j), []).append(j)
            for members in groups.values():
                # a group may exceed max_batch when one request alone brings many prompts: split on job boundaries
                chunk, size = [], 0
                for j in members + [None]:
                    if j is None or (chunk and size + len(j.prompts) > self.max_batch):
                        self._serve(chunk)
                        chunk, size = [], 0
                    if j is not None:
                        chunk.append(j)
                        size += len(j.prompts)

    def _serve(self, jobs):
        if not jobs:
            return
        prompts = [p for j in jobs for p in j.prompts]
        try:
            out = self.generate_fn(prompts, jobs[0].max_tokens)
            self.calls += 1
            if isinstance(out, str):
                out = [out]
            if len(out) != len(prompts):
                raise RuntimeError(f'generate returned {len(out)} responses for {len(prompts)} prompts')
            i = 0
            for j in jobs:
                j.responses = list(out[i:i + len(j.prompts)])
                i += len(j.prompts)
        except Exception as e:                               # every waiting client gets the failure, the worker lives on
            for j in jobs:
                j.error = e
        for j in jobs:
            j.done.set()


def make_handler(batcher, model_name='phi-3-vision'):
    class Handler(BaseHTTPRequestHandler):
        def do_POST(self):                                   # srv:8-30
            if self.path != '/v1/completions':
                self.send_error(404, 'Not Found')
                return
            try:
                req = json.loads(self.rfile.read(int(self.headers['Content-Length'])).decode('utf-8'))
                prompts = req.get('prompt', '')
                if isinstance(prompts, str):
                    prompts = [prompts]
                responses = batcher.submit(prompts, req.get('max_tokens', 512))
            except (ValueError, KeyError, TypeError) as e:
                self.send_error(400, str(e))
                return
            except Exception as e:
                self.send_error(500, str(e))
                return
            body = json.dumps({'model': model_name, 'responses': responses}).encode('utf-8')
            self.send_response(200)
            self.send_header('Content-Type', 'application/json')
            self.send_header('Content-Length', str(len(body)))
            self.end_headers()
            self.wfile.write(body)

        def log_message(self, *a):                           # quiet by default
            pass
    return Handler


def model_generate_fn(preload):
    """generate_fn over the B200 path: text prompts -> one batched generate_batch call (per-prompt results equal B=1 calls)"""
    from . import api

    def fn(prompts, max_tokens):
        return api.generate_batch(prompts, preload=preload, max_tokens=max_tokens, verbose=False)
    return fn


def rope_side_key(preload):
    """group key: (max_tokens, which side of the LongRoPE switch the request's prompts fall on)"""
    from . import api
    model, processor = preload
    orig = model.cfg.original_max_position_embeddings

    def key(job):
        texts = [api._apply_chat_template(p, None, False, True)[0] for p in job.prompts]
        lens = api._prompt_lengths(processor, texts, [None] * len(texts))
        return (job.max_tokens, (max(lens) + job.max_tokens) > orig)
    return key


def serve(generate_fn, port=8000, max_batch=16, window_ms=2.0, host='', group_key=None):
    batcher = Batcher(generate_fn, max_batch=max_batch, window_ms=window_ms, group_key=group_key)
    httpd = ThreadingHTTPServer((host, port), make_handler(batcher))
    httpd.batcher = batcher
    return httpd


def run(port=8000, max_batch=16, window_ms=2.0, **load_kwargs):
    from . import api
    preload = api.load(**load_kwargs)                        # srv:5: the model is loaded once, at start-up
    httpd = serve(model_generate_fn(preload), port, max_batch, window_ms, group_key=rope_side_key(preload))
    print(f'Starting server on port {port}')
    try:
        httpd.serve_forever()
    finally:
        httpd.batcher.close()


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--port', type=int, default=8000)
    ap.add_argument('--max-batch', type=int, default=16)
    ap.add_argument('--window-ms', type=float, default=2.0)
    ap.add_argument('--quantize-model', action='store_true')
    ap.add_argument('--quantize-cache', action='store_true')
    a = ap.parse_args()
    run(a.port, a.max_batch, a.window_ms, quantize_model=a.quantize_model, quantize_cache=a.quantize_cache)
