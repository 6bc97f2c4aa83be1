"""CPU oracle for the YOGO hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``yogo_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` use it, and there only as the checker
or as the thing timed on host cores.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real
reference from ``/root/reference`` (with the stubs listed in SURVEY.md 8c),
runs it on seeded inputs and commits its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every oracle function against those
vectors and against the reference's own known-answer tests
(tests/test_utils_tensor_formatting.py, tests/test_count_predictions.py).
"""
