"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference algorithms on the SmartStart hot path
(Gaussian-KDE + UCB start-state selection, NND_MB random-shooting MPC).
Nothing under this package may be imported by the product package
``smartstartcontinuous_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs use it, and only as the
checker / the timed CPU arm -- never as a fallback for the CUDA path.

Parity status: the reference's own tests pin only the geometry helpers
(tests/utilities/test_numerical.py) and the replay buffer bookkeeping
(tests/RLAgents/test_replayBuffer.py).  KDE densities, UCB choice, forward-sim
states and MPC scores are pinned by golden vectors produced by *running the
reference's own Python code* (oracle/ref_harness.py + oracle/make_golden.py,
executed in the build container where /root/reference is mounted) and committed
under tests/golden/.
"""
