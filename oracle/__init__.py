"""ORACLE — test infrastructure only.

CPU restatements of the algorithms the reference delegates to llm-compressor and llama.cpp
(SURVEY.md §A-§D).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; nothing under
``quantool_b200/`` does (tests/test_no_oracle_in_product.py enforces it).
"""
