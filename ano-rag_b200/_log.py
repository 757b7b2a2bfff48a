"""loguru when present (the reference logs through it), stdlib logging otherwise."""
try:
    from loguru import logger  # noqa: F401
except Exception:  # pragma: no cover
    import logging

    logger = logging.getLogger("ano_rag_b200")
