"""CPU oracle for the rasterizer hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
package; ``splatam_b200`` (the product) never does.
"""
