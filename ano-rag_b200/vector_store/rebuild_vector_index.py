#!/usr/bin/env python
"""Rebuild of the vector index that feeds the scoring path -- drop-in for the reference's
``vector_store/rebuild_vector_index.py`` (``rebuild_vector_index`` :127-190, ``should_rebuild_index`` :88-119,
``_incremental_rebuild`` :202-234, ``_full_rebuild`` :236-245, ``_test_retrieval`` :262-272).

What the reference's script does around the index -- turning raw documents into atomic notes with an LLM (:30-86) and the
embedding-version bookkeeping of ``EmbeddingStrategy`` -- is outside the scoring path (out of scope): notes
arrive here already made (a ``.json`` list or a ``.jsonl`` file with one note per line) and the version object is anything
with ``current_version_id`` / ``versions`` (or None: then only the note-count and forced triggers apply).  The index
itself is built by the device path: ``VectorRetriever.build_index`` -> ``VectorIndex.add_vectors`` -> the row-preparation
kernel, saved in faiss's own on-disk format (``faiss_io``).

    python -m ano_rag_b200.vector_store.rebuild_vector_index --data_path notes.jsonl [--force] [--incremental]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import sys
import time
from typing import Any, Dict, List, Optional

from .._log import logger
from ..config import config
from .retriever import VectorRetriever


def load_atomic_notes(path: str) -> List[Dict[str, Any]]:
    """Atomic notes from ``atomic_notes.json`` (a list, the file ``VectorRetriever._save_index_data`` writes,
    retriever.py:724-749) or from a ``.jsonl`` file with one note per line."""
    with open(path, "r", encoding="utf-8") as f:
        if path.endswith(".jsonl"):
            return [json.loads(line) for line in f if line.strip()]
        data = json.load(f)
    if isinstance(data, dict):
        data = data.get("atomic_notes", data.get("notes", []))
    return list(data)


def _rebuild_config() -> Dict[str, Any]:
    return (config.get("embedding_strategy", {}) or {}).get("index_rebuilding", {}) or {}


def _calculate_config_hash() -> str:                                                      # :121-125
    return hashlib.md5(json.dumps(config.get("embedding_strategy", {}) or {}, sort_keys=True).encode()).hexdigest()


def should_rebuild_index(atomic_notes: List[Dict[str, Any]], embedding_strategy: Any = None) -> bool:
    """The reference's three triggers (:88-119): note count, age of the current index version, config hash."""
    triggers = _rebuild_config().get("rebuild_triggers", {}) or {}
    note_threshold = triggers.get("note_count_threshold", 10000)
    if len(atomic_notes) >= note_threshold:
        logger.info(f"note count {len(atomic_notes)} reached the threshold {note_threshold}: rebuild")
        return True
    version = None
    if embedding_strategy is not None and getattr(embedding_strategy, "current_version_id", None):
        version = getattr(embedding_strategy, "versions", {}).get(embedding_strategy.current_version_id)
    if version is not None:
        age_hours = (time.time() - version.created_at) / 3600
        limit = triggers.get("time_threshold_hours", 168)
        if age_hours >= limit:
            logger.info(f"index age {age_hours:.1f} h reached the threshold {limit}: rebuild")
            return True
        if triggers.get("config_change", True) and version.metadata.get("config_hash") != _calculate_config_hash():
            logger.info("embedding configuration changed: rebuild")
            return True
    return False


def _incremental_rebuild(retriever: VectorRetriever, atomic_notes: List[Dict[str, Any]], batch_size: int,
                         checkpoint_interval: int) -> bool:                               # :202-234
    total, processed = len(atomic_notes), 0
    for i in range(0, total, batch_size):
        batch = atomic_notes[i:i + batch_size]
        try:
            if not retriever.add_notes(batch, rebuild_index=False):
                logger.error(f"batch {i // batch_size + 1} failed")
                return False
            processed += len(batch)
            logger.info(f"processed {processed}/{total} notes")
            if processed % checkpoint_interval == 0:
                logger.info(f"checkpoint at {processed} notes")
        except Exception as e:
            logger.error(f"batch {i // batch_size + 1} raised: {e}")
            return False
    return True


def _full_rebuild(retriever: VectorRetriever, atomic_notes: List[Dict[str, Any]], save_index: bool) -> bool:   # :236-245
    return retriever.build_index(atomic_notes=atomic_notes, force_rebuild=True, save_index=save_index)


def _test_retrieval(retriever: VectorRetriever) -> List[int]:
    """Smoke queries of :262-272; returns the result counts the reference only logs."""
    counts = []
    for query in ["测试查询", "example query", "示例问题"]:
        try:
            results = retriever.retrieve(query, top_k=3)
            logger.info(f"test query '{query}' returned {len(results)} results")
            counts.append(len(results))
        except Exception as e:
            logger.warning(f"test query '{query}' failed: {e}")
            break
    return counts


def rebuild_vector_index(atomic_notes: List[Dict[str, Any]], save_index: bool = True, force_rebuild: bool = False,
                         incremental: bool = False, retriever: Optional[VectorRetriever] = None,
                         embedding_strategy: Any = None) -> bool:
    """Same contract as the reference's function (:127-190): True on success or when no rebuild is needed, False on any
    failure, never raises.  ``retriever`` / ``embedding_strategy`` default to what the reference constructs itself."""
    rc = _rebuild_config()
    enable_incremental = rc.get("enable_incremental", True) and incremental
    batch_size = rc.get("batch_size", 1000)
    checkpoint_interval = rc.get("checkpoint_interval", 5000)
    if not force_rebuild and not should_rebuild_index(atomic_notes, embedding_strategy):
        logger.info("index does not need a rebuild")
        return True
    try:
        retriever = retriever if retriever is not None else VectorRetriever()
        if enable_incremental and not force_rebuild:
            ok = _incremental_rebuild(retriever, atomic_notes, batch_size, checkpoint_interval)
        else:
            ok = _full_rebuild(retriever, atomic_notes, save_index)
        if ok:
            logger.info(f"config hash {_calculate_config_hash()}, {len(atomic_notes)} notes")
            try:
                logger.info(f"index stats: {retriever.vector_index.get_index_stats()}")
            except Exception as e:
                logger.warning(f"index stats unavailable: {e}")
            _test_retrieval(retriever)
        else:
            logger.error("vector index rebuild failed")
        return bool(ok)
    except Exception as e:
        logger.error(f"rebuild raised: {e}")
        return False


def main(argv: Optional[List[str]] = None) -> int:
    ap = argparse.ArgumentParser(description="rebuild the vector index on the B200 path")
    ap.add_argument("--data_path", required=True, help="atomic notes: a .json list or a .jsonl file (one note per line)")
    ap.add_argument("--no_save", action="store_true")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--incremental", action="store_true")
    ap.add_argument("--dry_run", action="store_true")
    ap.add_argument("--batch_size", type=int)
    args = ap.parse_args(argv)
    if args.incremental and args.force:                                                   # :293-295
        logger.warning("--incremental and --force together: full rebuild")
        args.incremental = False
    try:
        notes = load_atomic_notes(args.data_path)
    except Exception as e:
        logger.error(f"could not load atomic notes: {e}")
        return 1
    if not notes:
        logger.error("no atomic notes loaded")
        return 1
    if args.dry_run:
        logger.info("index needs a rebuild" if should_rebuild_index(notes) else "index does not need a rebuild")
        return 0
    if args.batch_size:
        config.update({"embedding_strategy": {"index_rebuilding": {"batch_size": args.batch_size}}})
    t0 = time.time()
    ok = rebuild_vector_index(notes, save_index=not args.no_save, force_rebuild=args.force, incremental=args.incremental)
    dt = time.time() - t0
    if ok:
        logger.info(f"rebuild finished in {dt:.2f} s ({len(notes) / dt if dt > 0 else 0:.1f} notes/s)")
        return 0
    logger.error(f"rebuild failed after {dt:.2f} s")
    return 1


if __name__ == "__main__":
    sys.exit(main())
