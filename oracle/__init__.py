"""CPU oracle for the brute-force KNN hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pathway_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker or
as the timed CPU baseline, never as the shipped compute path.
"""
