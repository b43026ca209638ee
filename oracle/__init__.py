"""CPU oracle for the scedar distance-matrix / kNN hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``scedar_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` use it, and there only as the
checker (or as the timed CPU arm), never as the product path.

Parity status: PINNED.  ``oracle.sdm_oracle`` is checked (tests/test_oracle_*)
against the reference's own known-answer tests and
against golden vectors produced by importing the live reference in the build
container (tests/golden/make_golden.py).
"""
