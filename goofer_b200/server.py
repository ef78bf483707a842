"""HTTP front-end that coalesces concurrent resampler requests into GPU batches.

Drop-in for the reference's server mode (/root/reference/SillySampler.py:1187-1224): POST to port 8572 with the
resampler arguments joined by spaces in the body (the last 11 tokens are arguments 3-13, the first two `*.wav`
matches are the paths); GET answers 200 (liveness); success -> 200, failure -> 500 with the traceback text.

The reference renders each POST in its own thread, one note at a time.  Here the handler threads only enqueue:
a single batcher thread drains the queue every `window_ms` (or as soon as `max_batch` notes wait), renders the
notes as ONE batch through the C ABI (cli.render_notes), writes the wav files and wakes the handlers.  OpenUtau
fires its per-note requests concurrently, so a phrase becomes a few GPU batches instead of hundreds of calls.
"""
from __future__ import annotations

import logging
import queue
import re
import threading
import traceback
from http.server import BaseHTTPRequestHandler, HTTPServer
from socketserver import ThreadingMixIn
from typing import Callable, List, Optional, Sequence

from . import cli, host

PORT = 8572


def split_arguments(body: str) -> List[str]:
    """SillySampler.py:1187-1194."""
    parts = body.split(" ")
    other = parts[-11:]
    paths = re.findall(r"([^\s]+\.wav)", " ".join(parts[:-11]))
    if len(paths) < 2:
        raise ValueError("Missing .wav file paths in POST string")
    return [paths[0], paths[1]] + other


class _Job:
    __slots__ = ("args", "done", "error")

    def __init__(self, args):
        self.args, self.done, self.error = args, threading.Event(), None


def _render_and_write(arg_lists: Sequence[Sequence[str]]) -> None:
    # 16-bit PCM encoded on the device (GooferBatch.out_pcm16); GOOFER_DEVICES shards the batch over several GPUs
    outs = cli.render_notes(arg_lists, pcm16=True, devices=cli.default_devices())
    for args, out in zip(arg_lists, outs):
        sr = cli._features(cli.feature_path(args[0])).sr     # cached: the file was read when the batch was assembled
        cli.write_wav_pcm16(args[1], out, sr)


class Batcher(threading.Thread):
    """Collects jobs for up to `window_ms`, renders them together.  If the batch fails as a whole (one bad note
    fails goofer_plan_batch) the jobs are retried one by one so that only the bad request gets the 500."""

    def __init__(self, render: Callable[[Sequence[Sequence[str]]], None] = _render_and_write, window_ms: float = 8.0,
                 max_batch: int = 2048):
        super().__init__(daemon=True)
        self.render, self.window, self.max_batch = render, window_ms / 1000.0, max_batch
        self.q: "queue.Queue[Optional[_Job]]" = queue.Queue()
        self.batches: List[int] = []                      # sizes of the batches rendered so far (for tests / metrics)

    def submit(self, args: Sequence[str]) -> _Job:
        job = _Job(list(args))
        self.q.put(job)
        return job

    def stop(self):
        self.q.put(None)

    def run(self):
        while True:
            job = self.q.get()
            if job is None:
                return
            jobs = [job]
            deadline = threading.Event()
            timer = threading.Timer(self.window, deadline.set)
            timer.start()
            while len(jobs) < self.max_batch and not deadline.is_set():
                try:
                    nxt = self.q.get(timeout=self.window / 4 + 1e-4)
                except queue.Empty:
                    continue
                if nxt is None:
                    self.q.put(None)
                    break
                jobs.append(nxt)
            timer.cancel()
            self._render(jobs)

    def _render(self, jobs: List[_Job]):
        self.batches.append(len(jobs))
        try:
            self.render([j.args for j in jobs])
        except Exception:
            if len(jobs) == 1:
                jobs[0].error = traceback.format_exc()
            else:
                for j in jobs:                              # isolate the failing request(s)
                    try:
                        self.render([j.args])
                    except Exception:
                        j.error = traceback.format_exc()
        for j in jobs:
            j.done.set()


class ThreadedHTTPServer(ThreadingMixIn, HTTPServer):
    daemon_threads = True


def make_handler(batcher: Batcher):
    class RequestHandler(BaseHTTPRequestHandler):
        def log_message(self, *a):                          # keep the console to the resampler's own messages
            pass

        def do_GET(self):
            self.send_response(200)
            self.end_headers()

        def do_POST(self):
            n = int(self.headers.get("Content-Length", "0"))
            body = self.rfile.read(n).decode("utf-8")
            try:
                job = batcher.submit(split_arguments(body))
                job.done.wait()
                err = job.error
            except Exception:
                err = traceback.format_exc()
            if err:
                self.send_response(500)
                self.send_header("Content-type", "text/plain")
                self.end_headers()
                self.wfile.write(f"An error occurred.\n{err}".encode("utf-8"))
                return
            self.send_response(200)
            self.end_headers()
    return RequestHandler


def run(port: int = PORT, window_ms: float = 8.0, render=_render_and_write):
    batcher = Batcher(render, window_ms)
    batcher.start()
    httpd = ThreadedHTTPServer(("", port), make_handler(batcher))
    logging.info(f"Starting HTTP server on port {port}...")
    try:
        httpd.serve_forever()
    finally:
        batcher.stop()


if __name__ == "__main__":
    logging.basicConfig(format="%(message)s", level=logging.INFO)
    run()
