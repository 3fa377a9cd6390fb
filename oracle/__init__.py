"""ORACLE / TEST INFRASTRUCTURE.

CPU restatement of the reference's hot path (MuJoCo 2.1.0 step subset + MyoSuite/reference env
logic).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product package ``myochallenge_b200``
never does.
"""
