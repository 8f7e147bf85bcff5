"""CPU oracle for the ALD hot path -- TEST INFRASTRUCTURE ONLY (see sbc_oracle.c header).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this."""
