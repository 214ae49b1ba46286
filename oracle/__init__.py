"""CPU oracle for the DCPT hot path (TEST INFRASTRUCTURE ONLY).

Everything under ``oracle/`` is a CPU restatement of the reference's algorithm
(MILab-PKU/dcpt, ``basicsr/archs``).  It exists to *check* the sm_100a CUDA
path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it; the product
package ``dcpt_b200`` never does (tests/test_no_oracle_in_product.py enforces
that).
"""
