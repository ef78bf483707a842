"""CPU: the coalescing HTTP front-end (SillySampler.py:1187-1224 protocol) with a stub renderer."""
import threading
import time
import urllib.error
import urllib.request
from http.server import HTTPServer

import pytest

from goofer_b200 import server


def test_split_arguments_matches_the_reference_rule():
    body = "C:/voice bank/a i.wav C:/cache/out 1.wav C4 100 g-10 0 1000 0 0 100 0 !120 AA"
    a = server.split_arguments(body)
    assert a[1] == "1.wav" or a[1].endswith(".wav")          # same regex as the reference: no spaces inside a match
    assert a[2:] == ["C4", "100", "g-10", "0", "1000", "0", "0", "100", "0", "!120", "AA"]
    with pytest.raises(ValueError):
        server.split_arguments("only_one.wav C4 100 g0 0 1000 0 0 100 0 !120 AA")


def test_concurrent_posts_are_rendered_as_batches():
    rendered = []

    def fake_render(arg_lists):
        if any("BAD" in a[2] for a in arg_lists):
            raise ValueError("Bad note 'BAD'")
        time.sleep(0.02)
        rendered.append([a[1] for a in arg_lists])

    batcher = server.Batcher(fake_render, window_ms=60.0)
    batcher.start()
    httpd = server.ThreadedHTTPServer(("127.0.0.1", 0), server.make_handler(batcher))
    port = httpd.server_address[1]
    t = threading.Thread(target=httpd.serve_forever, daemon=True)
    t.start()
    try:
        assert urllib.request.urlopen(f"http://127.0.0.1:{port}/").status == 200        # GET = liveness
        codes = {}

        def post(i, pitch):
            body = f"in{i}.wav out{i}.wav {pitch} 100 g0 0 1000 0 0 100 0 !120 AA".encode()
            try:
                codes[i] = urllib.request.urlopen(urllib.request.Request(f"http://127.0.0.1:{port}/", data=body)).status
            except urllib.error.HTTPError as e:
                codes[i] = e.code
                codes[f"body{i}"] = e.read().decode()

        th = [threading.Thread(target=post, args=(i, "BAD" if i == 5 else "C4")) for i in range(12)]
        for x in th:
            x.start()
        for x in th:
            x.join(timeout=30)
        assert all(codes[i] == 200 for i in range(12) if i != 5)
        assert codes[5] == 500 and "Bad note" in codes["body5"]                        # only the bad request fails
        assert sum(len(b) for b in rendered) == 11
        assert max(batcher.batches) > 1                                                 # requests were coalesced
    finally:
        httpd.shutdown()
        batcher.stop()
