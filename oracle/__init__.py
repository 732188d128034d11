"""TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement of the reference's PPO hot
path (or machinery to run the unmodified reference here).  It is the checker:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
(``cleanrl_b200/``) never does.
"""
