"""`LLM` is LLMEngine, as in ssd/llm.py:4."""
from .engine.llm_engine import LLMEngine


class LLM(LLMEngine):
    pass
