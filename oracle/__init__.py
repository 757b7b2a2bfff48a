"""CPU oracle for AnoRAG's query-time candidate-scoring path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ano-rag_b200/`` may import this
package; the only legal callers are ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` (as the checker
or the timed CPU baseline, never as the product path).

What it restates (all citations relative to the reference tree):

* dense     -- ``vector_store/vector_index.py:206-282`` (pre-processing and result
               shaping) + the documented contract of ``faiss.IndexFlatIP.search``
               (faiss >= 1.7.4, requirements.txt:9-10; NOT vendored, NOT installed).
* bm25      -- ``utils/bm25_search.py:16-63,237-340`` verbatim semantics +
               ``rank_bm25.BM25Okapi`` (unpinned, absent from requirements.txt;
               optional import at ``utils/bm25_search.py:7-13``), restated from the
               published rank_bm25 0.2.x algorithm.
* fusion    -- ``retrieval/hybrid_search.py:13-103``.

Parity pinning status
---------------------
* ``SimpleBM25`` / ``bm25_scores`` / ``HybridSearcher.fuse``: PINNED.  The reference
  files load by path in the build container; ``tests/golden/make_golden.py`` ran
  them and committed their outputs under ``tests/golden/``; the oracle is checked
  bit-for-bit against those fixtures (and against the live reference files when
  ``/root/reference`` exists).
* ``BM25Okapi`` and ``IndexFlatIP``: PARITY UNPINNED.  rank_bm25 and faiss are not
  installed and not installable offline, and the reference holds no numeric test
  for them; the restatements follow the published algorithms and the reference's
  call sites only.
"""

from . import dense, bm25, fusion  # noqa: F401
