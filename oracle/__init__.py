"""TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's algorithms (numpy / plain torch fp32), used as the checker by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` arm.
Nothing under ``artspeech_b200/`` imports this package; the product path has no CPU fallback.
"""
