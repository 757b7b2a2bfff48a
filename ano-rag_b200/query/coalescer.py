"""Request coalescer: concurrent single-query calls become one batched device call.

The reference calls the scoring path one query at a time -- ``self.vector_retriever.search([entity])`` inside the
per-entity loop of ``QueryProcessor._second_hop_fallback`` (query/query_processor.py:1478-1482), one call per
sub-question from a four-thread pool (:2761-2764), ``EnhancedRecallOptimizer`` (vector_store/
enhanced_recall_optimizer.py:421,521).  One query costs the device 0.54 ms of latency (the corpus is streamed once
whatever the batch holds), 128 queries cost 1.35 ms: callers that overlap in time should share a launch.

``RequestCoalescer`` wraps any ``batch_fn(list_of_requests) -> list_of_results``.  ``submit`` returns a future;
a worker thread takes whatever is queued (up to ``max_batch``), waits at most ``max_wait_ms`` for stragglers once
the first request is in hand, runs ONE batch call and distributes the results.  Errors of a batch call reach
every future of that batch.  ``coalesced_search`` adapts ``VectorRetriever.search`` (list of query strings in,
list of result lists out) and ``coalesced_pipeline`` adapts ``HybridPipeline.search_host``.
"""
from __future__ import annotations

import queue
import threading
import time
from concurrent.futures import Future
from typing import Any, Callable, List, Sequence

import numpy as np


class RequestCoalescer:
    def __init__(self, batch_fn: Callable[[List[Any]], Sequence[Any]], max_batch: int = 128, max_wait_ms: float = 0.2):
        self.batch_fn = batch_fn
        self.max_batch = int(max_batch)
        self.max_wait = float(max_wait_ms) * 1e-3
        self.batches = 0                      # batch calls issued
        self.requests = 0                     # requests served
        self._q: "queue.Queue" = queue.Queue()
        self._closed = False
        self._worker = threading.Thread(target=self._run, name="arb-coalescer", daemon=True)
        self._worker.start()

    def submit(self, request: Any) -> Future:
        if self._closed:
            raise RuntimeError("coalescer is closed")
        fut: Future = Future()
        self._q.put((request, fut))
        return fut

    def __call__(self, request: Any) -> Any:
        return self.submit(request).result()

    def close(self) -> None:
        if not self._closed:
            self._closed = True
            self._q.put(None)
            self._worker.join()

    def _run(self) -> None:
        while True:
            item = self._q.get()
            if item is None:
                return
            batch = [item]
            deadline = time.perf_counter() + self.max_wait
            stop = False
            while len(batch) < self.max_batch:
                try:
                    left = deadline - time.perf_counter()
                    nxt = self._q.get(timeout=left) if left > 0 else self._q.get_nowait()
                except queue.Empty:
                    break
                if nxt is None:
                    stop = True
                    break
                batch.append(nxt)
            try:
                results = self.batch_fn([r for r, _ in batch])
                if len(results) != len(batch):
                    raise RuntimeError(f"batch_fn returned {len(results)} results for {len(batch)} requests")
                for (_, fut), res in zip(batch, results):
                    fut.set_result(res)
            except BaseException as e:                                   # every caller of the batch sees the failure
                for _, fut in batch:
                    if not fut.done():
                        fut.set_exception(e)
            self.batches += 1
            self.requests += len(batch)
            if stop:
                return


def coalesced_search(retriever, top_k=None, similarity_threshold=None, include_metadata=True, **kw) -> RequestCoalescer:
    """``c = coalesced_search(vector_retriever); c("bridge entity")`` == ``vector_retriever.search([q])[0]``, with calls
    that overlap in time sharing one ``search`` of all their queries (vector_store/retriever.py:186-272 takes a list)."""
    return RequestCoalescer(lambda qs: retriever.search(list(qs), top_k=top_k, similarity_threshold=similarity_threshold,
                                                        include_metadata=include_metadata), **kw)


def coalesced_pipeline(pipe, **kw) -> RequestCoalescer:
    """Requests are ``(embedding [d], term ids [<= mq] or None)``; results ``(ids, final, count)`` rows of
    ``HybridPipeline.search_host`` (copies).  Term rows are padded with -1 to the longest of the batch."""
    def run(reqs):
        emb = np.stack([np.asarray(e, dtype=np.float32).reshape(-1) for e, _ in reqs])
        terms = None
        if any(t is not None for _, t in reqs):
            width = max(1, max(len(t) for _, t in reqs if t is not None))
            terms = np.full((len(reqs), width), -1, dtype=np.int32)
            for i, (_, t) in enumerate(reqs):
                if t is not None:
                    terms[i, :len(t)] = np.asarray(t, dtype=np.int32)
        ids, fin, cnt = pipe.search_host(emb, terms)[:3]
        return [(ids[i].copy(), fin[i].copy(), int(cnt[i])) for i in range(len(reqs))]
    return RequestCoalescer(run, **kw)
